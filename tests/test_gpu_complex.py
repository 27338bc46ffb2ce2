"""-m gpu: the complex path (cublasZgemm / gemm(..., complx); reference src/gemm.cu:412-521) through the
C-ABI against the CPU oracle and, bit for bit, against the unmodified reference."""
import numpy as np
import pytest
import torch

import oracle_lib
import ozimmu_b200 as oz
from gpu_util import Reference, bits, to_dev, ulp_distance

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("op_a,op_b", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("m,n,k,num_split,kind,alpha,beta", [
    (64, 48, 100, 9, "normal01", 1.0 + 0.0j, 0.0 + 0.0j),
    (130, 70, 257, 13, "exp_rand-1", -1.5 + 0.5j, 0.75 - 0.25j),
    (33, 200, 16, 3, "urand01", 0.0 + 2.0j, 0.0 + 1.0j),
    (129, 131, 300, 18, "mixed", 1.0 - 1.0j, 1.0 + 0.0j),
])
def test_zgemm_matches_oracle(handle, op_a, op_b, m, n, k, num_split, kind, alpha, beta):
    lda = m if op_a == 0 else k
    ldb = k if op_b == 0 else n
    a = oracle_lib.gen_complex(kind, m * k, m + k)
    b = oracle_lib.gen_complex(kind, k * n, n + k)
    c = oracle_lib.gen_complex("normal01", m * n, 5)
    want = oracle_lib.oracle_gemm_complex(op_a, op_b, m, n, k, alpha, a, lda, b, ldb, beta, c, m, num_split)
    da, db, dc = to_dev(a), to_dev(b), to_dev(c)
    rc = oz.gemm(handle, op_a, op_b, m, n, k, alpha, da, lda, db, ldb, beta, dc, m, oz.fp64_int8(num_split), oz.complx)
    assert rc == 0
    torch.cuda.synchronize()
    got = dc.cpu().numpy()
    assert np.array_equal(got.view(np.int64), want.view(np.int64)), \
        f"max ulp distance {ulp_distance(got.view(np.float64), want.view(np.float64))}"


@pytest.mark.parametrize("planes_first", ["0", "1"])
@pytest.mark.parametrize("shape", [(64, 128), (0, 128), (0, 256)])
def test_zgemm_forced_tiles_match_oracle(handle, shape, planes_first, monkeypatch):
    """the complex epilogue (four plane products per tile) under every accumulator layout of the fused kernel: 64 rows
    per CTA (UMMA M = 128, columns folded over the TMEM lanes), 128 rows all in registers, 128 rows with spill columns"""
    m, n, k, num_split = 300, 270, 200, 9
    a = oracle_lib.gen_complex("exp_rand-1", m * k, 21)
    b = oracle_lib.gen_complex("exp_rand-1", k * n, 22)
    c = oracle_lib.gen_complex("normal01", m * n, 23)
    alpha, beta = 0.5 - 1.25j, -0.75 + 0.0j
    want = oracle_lib.oracle_gemm_complex(0, 1, m, n, k, alpha, a, m, b, n, beta, c, m, num_split)
    da, db, dc = to_dev(a), to_dev(b), to_dev(c)
    monkeypatch.setenv("OZIMMU_B200_ZGEMM_PLANES_FIRST", planes_first)   # fused four-group launch / planes + combine
    oz.lib().ozk_set_cluster_shape(*shape)
    try:
        assert oz.gemm(handle, 0, 1, m, n, k, alpha, da, m, db, n, beta, dc, m, oz.fp64_int8(num_split), oz.complx) == 0
        torch.cuda.synchronize()
    finally:
        oz.lib().ozk_set_cluster_shape(0, 0)
    assert np.array_equal(dc.cpu().numpy().view(np.int64), want.view(np.int64))


@pytest.mark.parametrize("op_a,op_b", [(0, 0), (1, 1)])
@pytest.mark.parametrize("m,n,k,num_split", [(1024, 1024, 1024, 9), (1024, 1023, 1025, 12), (2048, 512, 768, 16)])
def test_zgemm_bit_exact_vs_reference(handle, op_a, op_b, m, n, k, num_split):
    if oracle_lib.reference() is None:
        pytest.skip("oracle/_ref/libozref.so not built")
    ref = Reference()
    try:
        g = torch.Generator(device="cuda").manual_seed(num_split)
        lda = m if op_a == 0 else k
        ldb = k if op_b == 0 else n
        a = torch.randn(m * k, dtype=torch.complex128, device="cuda", generator=g)
        b = torch.randn(k * n, dtype=torch.complex128, device="cuda", generator=g)
        c0 = torch.randn(m * n, dtype=torch.complex128, device="cuda", generator=g)
        for alpha, beta in ((1.0 + 0j, 0j), (0.5 - 1.5j, -0.25 + 2.0j)):
            c_ref, c_new = c0.clone(), c0.clone()
            ref.gemm_complex(op_a, op_b, m, n, k, alpha, a, lda, b, ldb, beta, c_ref, m, num_split - 1)
            assert oz.gemm(handle, op_a, op_b, m, n, k, alpha, a, lda, b, ldb, beta, c_new, m, oz.fp64_int8(num_split),
                           oz.complx) == 0
            torch.cuda.synchronize()
            r_ref, r_new = torch.view_as_real(c_ref), torch.view_as_real(c_new)
            if beta.imag == 0:
                assert torch.equal(r_ref.view(torch.int64), r_new.view(torch.int64))
            else:
                # Im(beta) != 0: the reference's C = beta*C reads the updated real part when it forms the imaginary
                # part; the product does not.  Real parts are untouched by that and stay bit-identical;
                # the imaginary parts must be the accurate ones.
                assert torch.equal(r_ref[:, 0].contiguous().view(torch.int64), r_new[:, 0].contiguous().view(torch.int64))
                A = a.view(k, lda)[:, :m].T if op_a == 0 else a.view(m, lda)[:, :k]
                B = b.view(n, ldb)[:, :k].T if op_b == 0 else b.view(k, ldb)[:, :n]
                want = alpha * (A @ B) + beta * c0.view(n, m).T
                scale = want.abs().max().item()
                err_new = (c_new.view(n, m).T - want).abs().max().item() / scale
                err_ref = (c_ref.view(n, m).T - want).abs().max().item() / scale
                assert err_new < 1e-12 and err_ref > 1e-3, (err_new, err_ref)
        # accuracy against cuBLAS ZGEMM (column-major: C^T = B^T A^T in torch's row-major view)
        if op_a == 0 and op_b == 0 and num_split >= 12:
            c_new = torch.zeros_like(c0)
            assert oz.gemm(handle, 0, 0, m, n, k, 1.0 + 0j, a, m, b, k, 0j, c_new, m, oz.fp64_int8(num_split), oz.complx) == 0
            want = (b.view(n, k) @ a.view(k, m)).reshape(-1)
            assert (torch.linalg.vector_norm(c_new - want) / torch.linalg.vector_norm(want)).item() < 1e-14
    finally:
        ref.close()


def test_complex_auto_mode_vs_oracle(handle):
    m, n, k = 96, 80, 160
    a = oracle_lib.gen_complex("exp_rand-2", m * k, 1)
    b = oracle_lib.gen_complex("exp_rand-2", k * n, 2)
    import ctypes as C
    L = oracle_lib.oracle()
    for thr in (0.0, 1.0, 4.0):
        cnt_want = np.zeros(16, dtype=np.uint64)
        s = L.oz_auto_select_complex(0, 0, m, n, k, a.ctypes.data, m, b.ctypes.data, k, thr, cnt_want.ctypes.data)
        cnt = []
        mode = oz.auto_mode_select(handle, 0, 0, m, n, k, to_dev(a), m, to_dev(b), k, oz.complx, thr, cnt)
        assert cnt == [int(v) for v in cnt_want]
        assert mode == (oz.fp64_int8(s) if s else oz.compute_mode_t.dgemm)


def test_zgemm_is_one_product_launch(handle):
    """the four plane products of a complex GEMM run inside ONE launch (the reference: 4 x P(s) cuBLAS GEMMs plus
    4 x (P(s) + 1) elementwise kernels, src/gemm.cu:476-518)"""
    m, n, k = 512, 384, 256
    g = torch.Generator(device="cuda").manual_seed(3)
    a = torch.randn(m * k, dtype=torch.complex128, device="cuda", generator=g)
    b = torch.randn(k * n, dtype=torch.complex128, device="cuda", generator=g)
    c = torch.zeros(m * n, dtype=torch.complex128, device="cuda")
    assert oz.gemm(handle, 1, 1, m, n, k, 1.0 + 0j, a, k, b, n, 0j, c, m, oz.fp64_int8(9), oz.complx) == 0   # warm
    before = oz.launch_count()
    assert oz.gemm(handle, 1, 1, m, n, k, 1.0 + 0j, a, k, b, n, 0j, c, m, oz.fp64_int8(9), oz.complx) == 0
    torch.cuda.synchronize()
    # op_t A / op_t B: A's rows are contiguous (1 split kernel per plane), B's strided (2 per plane) -> 6 + 1 product
    assert oz.launch_count() - before <= 7
    want = (a.view(m, k) @ b.view(k, n)).T.reshape(-1)
    assert (torch.linalg.vector_norm(c - want) / torch.linalg.vector_norm(want)).item() < 1e-14


@pytest.mark.parametrize("beta", [0j, 0.5 - 2.0j])
def test_zgemm_k_zero_scales_c(handle, beta):
    """k == 0: C = beta * C for complex data as well (the beta pre-scale of the complex path)"""
    m, n = 70, 33
    c = oracle_lib.gen_complex("normal01", m * n, 9)
    dc = to_dev(c)
    dummy = torch.zeros(4, dtype=torch.complex128, device="cuda")
    assert oz.gemm(handle, 0, 0, m, n, 0, 1.0 + 1.0j, dummy, m, dummy, 1, beta, dc, m, oz.fp64_int8(9), oz.complx) == 0
    torch.cuda.synchronize()
    got = dc.cpu().numpy()
    want = beta * c
    assert np.allclose(got, want, rtol=1e-15, atol=0)


@pytest.mark.parametrize("kind", ["real", "complex"])
def test_device_pointer_mode_scalars(handle, kind):
    """cuBLAS device pointer mode: alpha / beta live in device memory and are read by the kernels in stream order
    (the reference dereferences them on the host, src/gemm.cu:405); bits equal to the host-scalar call, for the
    fp64_int8 modes (device read) and for auto mode (one blocking fetch)"""
    m, n, k = 300, 260, 200
    if kind == "real":
        a, b, c = (to_dev(oracle_lib.gen_matrix("exp_rand-1", cnt, sd)) for cnt, sd in ((m * k, 1), (k * n, 2), (m * n, 3)))
        alpha, beta, ek = 1.5, -0.25, oz.real
        d_alpha = torch.tensor([alpha], dtype=torch.float64, device="cuda")
        d_beta = torch.tensor([beta], dtype=torch.float64, device="cuda")
    else:
        a, b, c = (to_dev(oracle_lib.gen_complex("exp_rand-1", cnt, sd)) for cnt, sd in ((m * k, 1), (k * n, 2), (m * n, 3)))
        alpha, beta, ek = 0.5 - 1.5j, -0.25 + 2.0j, oz.complx
        d_alpha = torch.tensor([alpha], dtype=torch.complex128, device="cuda")
        d_beta = torch.tensor([beta], dtype=torch.complex128, device="cuda")
    for mode in (oz.fp64_int8(9), oz.compute_mode_t.fp64_int8_auto):
        want, got = c.clone(), c.clone()
        assert oz.gemm(handle, 0, 0, m, n, k, alpha, a, m, b, k, beta, want, m, mode, ek) == 0
        oz.set_scalar_pointer_mode(handle, True)
        try:
            assert oz.gemm(handle, 0, 0, m, n, k, d_alpha, a, m, b, k, d_beta, got, m, mode, ek) == 0
        finally:
            oz.set_scalar_pointer_mode(handle, False)
        torch.cuda.synchronize()
        assert torch.equal(torch.view_as_real(got).view(torch.int64) if kind == "complex" else got.view(torch.int64),
                           torch.view_as_real(want).view(torch.int64) if kind == "complex" else want.view(torch.int64))
