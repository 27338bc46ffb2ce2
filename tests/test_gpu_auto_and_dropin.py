"""-m gpu: auto mode (mantissa-loss totals + selection) against oracle and reference, the dgemm
passthrough, and the LD_PRELOAD drop-in (cublasDgemm interception under an unmodified torch)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import oracle_lib
import ozimmu_b200 as oz
from gpu_util import Reference, bits, to_dev

pytestmark = pytest.mark.gpu
ROOT = oracle_lib.ROOT


@pytest.mark.parametrize("op_a,op_b", [(0, 0), (1, 0), (0, 1), (1, 1)])
@pytest.mark.parametrize("phi", [0.0, 1.0, 4.0])
def test_auto_counters_and_selection_vs_oracle(handle, op_a, op_b, phi):
    m, n, k = 96, 80, 160
    a = oracle_lib.gen_matrix(f"exp_rand-{phi}", m * k, 1)
    b = oracle_lib.gen_matrix(f"exp_rand-{phi}", k * n, 2)
    lda = m if op_a == 0 else k
    ldb = k if op_b == 0 else n
    da, db = to_dev(a), to_dev(b)
    for thr in (0.0, 0.5, 1.5, 8.0):
        want_s, want_cnt = oracle_lib.oracle_auto_select(op_a, op_b, m, n, k, a, lda, b, ldb, thr)
        cnt = []
        mode = oz.auto_mode_select(handle, op_a, op_b, m, n, k, da, lda, db, ldb, oz.real, thr, cnt)
        assert cnt == [int(v) for v in want_cnt]
        assert mode == (oz.fp64_int8(want_s) if want_s else oz.compute_mode_t.dgemm)


def test_auto_vs_reference_counters(handle):
    """Reference counters are only defined for fp64_int8_3..10 (8 counters) and for
    inputs without exact zeros with k % 32 == 0 (App. B.2)."""
    if oracle_lib.reference() is None:
        pytest.skip("oracle/_ref/libozref.so not built")
    ref = Reference()
    try:
        m = n = k = 1024
        for phi in (0.0, 1.0, 2.0):
            a = to_dev(oracle_lib.gen_matrix(f"exp_rand-{phi}", m * k, 3))
            b = to_dev(oracle_lib.gen_matrix(f"exp_rand-{phi}", k * n, 4))
            for thr in (0.0, 1.0, 4.0):
                ref_mode, ref_cnt = ref.auto_mode_select(0, 0, m, n, k, a, m, b, k, thr)
                cnt = []
                mode = oz.auto_mode_select(handle, 0, 0, m, n, k, a, m, b, k, oz.real, thr, cnt)
                assert cnt[:8] == ref_cnt
                if oz.compute_mode_t.fp64_int8_3 <= ref_mode <= oz.compute_mode_t.fp64_int8_10:
                    assert int(mode) == ref_mode
    finally:
        ref.close()


def _exp_rand_dev(phi: float, count: int, seed: int) -> torch.Tensor:
    """exp_rand-phi on the device (reference test/main_test.cu:56-70: (u - 0.5) * exp(phi * normal))"""
    g = torch.Generator(device="cuda").manual_seed(seed)
    u = torch.rand(count, dtype=torch.float64, device="cuda", generator=g) - 0.5
    u = torch.where(u == 0, torch.full_like(u, 0.25), u)      # no exact zeros
    return u * torch.exp(phi * torch.randn(count, dtype=torch.float64, device="cuda", generator=g))


@pytest.mark.parametrize("phi", [0.0, 0.5, 1.0, 2.0, 4.0, 8.0])
def test_config5_auto_4096_vs_reference(handle, phi):
    """BASELINE config 5 at its own size: fp64_int8_auto on ill-conditioned 4096^3 inputs, thresholds of SURVEY 8(d).
    Counters for fp64_int8_3..10 and the selected mode (where the reference's 8 counters can express it) equal the
    unmodified reference's; the GEMM in the selected mode equals the reference's GEMM in that mode bit for bit."""
    if oracle_lib.reference() is None:
        pytest.skip("oracle/_ref/libozref.so not built")
    ref = Reference()
    try:
        m = n = k = 4096
        a, b = _exp_rand_dev(phi, m * k, 11), _exp_rand_dev(phi, k * n, 12)
        checked = set()
        for thr in (0.0, 0.5, 1.0, 1.5, 2.0, 4.0, 8.0):
            ref_mode, ref_cnt = ref.auto_mode_select(0, 0, m, n, k, a, m, b, k, thr)
            cnt = []
            mode = oz.auto_mode_select(handle, 0, 0, m, n, k, a, m, b, k, oz.real, thr, cnt)
            assert cnt[:8] == ref_cnt, (phi, thr)
            assert all(x >= y for x, y in zip(cnt, cnt[1:]))            # more slices never lose more
            in_range = oz.compute_mode_t.fp64_int8_3 <= ref_mode <= oz.compute_mode_t.fp64_int8_10
            if in_range:
                assert int(mode) == ref_mode, (phi, thr)
            else:
                # the reference has no counters beyond fp64_int8_10 (App. B.1): ours must need more than 10 slices too
                assert mode == oz.compute_mode_t.dgemm or oz.num_split_of(mode) > 10, (phi, thr, mode)
            if in_range and int(mode) not in checked and len(checked) < 2:
                checked.add(int(mode))
                oz.set_auto_mantissa_loss_threashold(handle, thr)
                c_new = torch.zeros(m * n, dtype=torch.float64, device="cuda")
                c_ref = torch.zeros_like(c_new)
                assert oz.gemm(handle, 0, 0, m, n, k, 1.0, a, m, b, k, 0.0, c_new, m, oz.compute_mode_t.fp64_int8_auto) == 0
                ref.gemm(0, 0, m, n, k, 1.0, a, m, b, k, 0.0, c_ref, m, ref_mode)
                torch.cuda.synchronize()
                assert torch.equal(c_new.view(torch.int64), c_ref.view(torch.int64)), (phi, thr)
        oz.set_auto_mantissa_loss_threashold(handle, 0.0)
    finally:
        ref.close()


def test_sgemm_mode_vs_reference_sgemm(handle):
    """compute mode `sgemm` against the REFERENCE's own sgemm mode (src/cublas_helper.cu:84-134: FP32 copies,
    cublasSgemm, widen): same conversions around the same library GEMM -- equal up to cuBLAS' FP32 summation order
    (the two private cuBLAS handles may pick different kernels), every output an FP32 value in both."""
    if oracle_lib.reference() is None:
        pytest.skip("oracle/_ref/libozref.so not built")
    ref = Reference()
    try:
        m, n, k = 1024, 768, 1536
        g = torch.Generator(device="cuda").manual_seed(5)
        a = torch.randn(m * k, dtype=torch.float64, device="cuda", generator=g)
        b = torch.randn(k * n, dtype=torch.float64, device="cuda", generator=g)
        c0 = torch.randn(m * n, dtype=torch.float64, device="cuda", generator=g)
        for op_a, op_b, alpha, beta in ((0, 0, 1.0, 0.0), (1, 0, -0.5, 1.5), (0, 1, 2.0, -1.0)):
            lda, ldb = (m if op_a == 0 else k), (k if op_b == 0 else n)
            c_ref, c_new = c0.clone(), c0.clone()
            ref.dgemm_f32(op_a, op_b, m, n, k, alpha, a, lda, b, ldb, beta, c_ref, m)
            assert oz.gemm(handle, op_a, op_b, m, n, k, alpha, a, lda, b, ldb, beta, c_new, m, oz.compute_mode_t.sgemm) == 0
            torch.cuda.synchronize()
            assert torch.equal(c_new, c_new.float().double()) and torch.equal(c_ref, c_ref.float().double())
            rel = (torch.linalg.vector_norm(c_new - c_ref) / torch.linalg.vector_norm(c_ref)).item()
            assert rel < 5e-7, (op_a, op_b, rel)
    finally:
        ref.close()


def test_auto_mode_gemm_runs_selected_mode(handle):
    m, n, k = 300, 200, 500
    a = to_dev(oracle_lib.gen_matrix("exp_rand-1", m * k, 5))
    b = to_dev(oracle_lib.gen_matrix("exp_rand-1", k * n, 6))
    oz.set_auto_mantissa_loss_threashold(handle, 1.0)
    assert oz.get_auto_mantissa_loss_threashold(handle) == 1.0
    mode = oz.auto_mode_select(handle, 0, 0, m, n, k, a, m, b, k, oz.real, 1.0)
    c_auto = torch.zeros(m * n, dtype=torch.float64, device="cuda")
    c_fix = torch.zeros_like(c_auto)
    assert oz.gemm(handle, 0, 0, m, n, k, 1.0, a, m, b, k, 0.0, c_auto, m, oz.compute_mode_t.fp64_int8_auto) == 0
    assert oz.gemm(handle, 0, 0, m, n, k, 1.0, a, m, b, k, 0.0, c_fix, m, mode) == 0
    torch.cuda.synchronize()
    assert torch.equal(c_auto.view(torch.int64), c_fix.view(torch.int64))


def test_dgemm_passthrough(handle):
    m, n, k = 200, 150, 100
    a = torch.randn(k, m, dtype=torch.float64, device="cuda")  # column-major m x k
    b = torch.randn(n, k, dtype=torch.float64, device="cuda")  # column-major k x n
    c = torch.zeros(n, m, dtype=torch.float64, device="cuda")
    assert oz.gemm(handle, 0, 0, m, n, k, 1.0, a, m, b, k, 0.0, c, m, oz.compute_mode_t.dgemm) == 0
    torch.cuda.synchronize()
    assert torch.allclose(c, b @ a, rtol=1e-12, atol=1e-12)


DROPIN = r"""
import os, torch
torch.manual_seed(0)
n = 1536
a = torch.rand(n, n, dtype=torch.float64, device="cuda")
b = torch.rand(n, n, dtype=torch.float64, device="cuda")
c = a @ b
torch.cuda.synchronize()
za = torch.randn(n, n, dtype=torch.complex128, device="cuda")
zb = torch.randn(n, n, dtype=torch.complex128, device="cuda")
zc = za @ zb
torch.cuda.synchronize()
torch.save({"a": a.cpu(), "b": b.cpu(), "c": c.cpu(), "za": za.cpu(), "zb": zb.cpu(), "zc": zc.cpu()},
           os.environ["OZ_DROPIN_OUT"])
"""


def test_ld_preload_dropin(tmp_path, handle):
    """An unmodified PyTorch program under LD_PRELOAD=libozimmu.so OZIMMU_COMPUTE_MODE=fp64_int8_9:
    its float64 matmul (cublasDgemm / cublasGemmEx) must be served by the Ozaki path, i.e. be
    bit-identical to a direct ozimmu_gemm call on the same operands, and must log the interception."""
    out = tmp_path / "dropin.pt"
    env = dict(os.environ, LD_PRELOAD=str(oz.LIB_PATH), OZIMMU_COMPUTE_MODE="fp64_int8_9", OZIMMU_INFO="1",
               OZIMMU_ENABLE_CULIP_PROFILING="1", OZ_DROPIN_OUT=str(out))
    p = subprocess.run([sys.executable, "-c", DROPIN], env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    assert "[CULiP Result][Dfp64_int8_9-" in p.stdout, p.stdout[-2000:]
    d = torch.load(out)
    n = d["a"].shape[0]
    # torch row-major C = A @ B  <=>  column-major C^T = B^T A^T: cuBLAS is called with (B, A)
    a, b = d["a"].cuda(), d["b"].cuda()
    c = torch.zeros(n, n, dtype=torch.float64, device="cuda")
    assert oz.gemm(handle, 0, 0, n, n, n, 1.0, b, n, a, n, 0.0, c, n, oz.fp64_int8(9)) == 0
    torch.cuda.synchronize()
    assert np.array_equal(bits(c), bits(d["c"]))
    # and it is an accurate DGEMM
    ref = d["a"] @ d["b"]
    assert (torch.linalg.norm(d["c"] - ref) / torch.linalg.norm(ref)).item() < 1e-15
    # complex128 matmul -> cublasZgemm / cublasGemmEx(C_64F) -> the complex Ozaki path
    assert "[CULiP Result][Zfp64_int8_9-" in p.stdout, p.stdout[-2000:]
    za, zb = d["za"].cuda(), d["zb"].cuda()
    zc = torch.zeros(n, n, dtype=torch.complex128, device="cuda")
    assert oz.gemm(handle, 0, 0, n, n, n, 1.0 + 0j, zb, n, za, n, 0j, zc, n, oz.fp64_int8(9), oz.complx) == 0
    torch.cuda.synchronize()
    assert torch.equal(torch.view_as_real(zc).cpu().view(torch.int64), torch.view_as_real(d["zc"]).view(torch.int64))
    zref = d["za"] @ d["zb"]
    assert (torch.linalg.norm(d["zc"] - zref) / torch.linalg.norm(zref)).item() < 1e-14


def test_cuda_graph_capture(handle):
    """The device entry has no hidden synchronisation (the reference calls cudaDeviceSynchronize twice per
    GEMM, src/split.cu:261): once the workspace is sized it can be captured into a CUDA graph and replayed."""
    m, n, k = 512, 640, 768
    a = to_dev(oracle_lib.gen_matrix("exp_rand-1", m * k, 1))
    b = to_dev(oracle_lib.gen_matrix("exp_rand-1", k * n, 2))
    c_eager = torch.zeros(m * n, dtype=torch.float64, device="cuda")
    c_graph = torch.zeros_like(c_eager)
    s = torch.cuda.Stream()
    oz.set_cuda_stream(handle, s)
    with torch.cuda.stream(s):
        assert oz.gemm(handle, 0, 0, m, n, k, 1.0, a, m, b, k, 0.0, c_eager, m, oz.fp64_int8(9)) == 0  # sizes the workspace
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            assert oz.gemm(handle, 0, 0, m, n, k, 1.0, a, m, b, k, 0.0, c_graph, m, oz.fp64_int8(9)) == 0
    for _ in range(3):
        c_graph.zero_()
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(c_graph.view(torch.int64), c_eager.view(torch.int64))
    # an eager call on ANOTHER stream after the capture: the handle must not wait on an event that was only ever
    # recorded inside the graph (cudaErrorInvalidValue before round 2)
    oz.set_cuda_stream(handle, None)
    c_after = torch.zeros_like(c_eager)
    assert oz.gemm(handle, 0, 0, m, n, k, 1.0, a, m, b, k, 0.0, c_after, m, oz.fp64_int8(9)) == 0
    torch.cuda.synchronize()
    assert torch.equal(c_after.view(torch.int64), c_eager.view(torch.int64))


@pytest.mark.parametrize("op_a,op_b,beta", [(0, 0, 0.0), (1, 0, -0.5), (0, 1, 2.0), (1, 1, 0.0)])
def test_sgemm_mode_real(handle, op_a, op_b, beta):
    """compute mode `sgemm` (reference src/cublas_helper.cu:84-134): FP64 in / out, FP32 GEMM inside -- equal to an
    FP32 matmul of the rounded operands up to FP32 summation order, padded leading dimensions untouched."""
    m, n, k, pad = 300, 200, 500, 3
    lda, ldb, ldc = (m if op_a == 0 else k) + pad, (k if op_b == 0 else n) + pad, m + pad
    A = torch.randn(k if op_a == 0 else m, lda, dtype=torch.float64, device="cuda")     # [col][row]
    B = torch.randn(n if op_b == 0 else k, ldb, dtype=torch.float64, device="cuda")
    C0 = torch.randn(n, ldc, dtype=torch.float64, device="cuda")
    C = C0.clone()
    assert oz.gemm(handle, op_a, op_b, m, n, k, 1.5, A, lda, B, ldb, beta, C, ldc, oz.compute_mode_t.sgemm) == 0
    torch.cuda.synchronize()
    opA = (A[:, :m].T if op_a == 0 else A[:, :k]).float()      # m x k
    opB = (B[:, :k].T if op_b == 0 else B[:, :n]).float()      # k x n
    ref = 1.5 * (opA.double() @ opB.double()) + beta * C0[:, :m].T.float().double()
    got = C[:, :m].T
    assert (torch.linalg.norm(got - ref) / torch.linalg.norm(ref)).item() < 1e-5
    assert torch.equal(got, got.float().double())               # every output is an FP32 value
    assert torch.equal(C[:, m:], C0[:, m:])                     # ld padding untouched


def test_sgemm_mode_complex(handle):
    m, n, k = 128, 96, 200
    A = torch.randn(k, m, dtype=torch.complex128, device="cuda")   # column-major m x k
    B = torch.randn(n, k, dtype=torch.complex128, device="cuda")   # column-major k x n
    C = torch.zeros(n, m, dtype=torch.complex128, device="cuda")
    assert oz.gemm(handle, 0, 0, m, n, k, 1.0 + 0.5j, A, m, B, k, 0j, C, m, oz.compute_mode_t.sgemm, oz.complx) == 0
    torch.cuda.synchronize()
    ref = (1.0 + 0.5j) * (B @ A)
    assert (torch.linalg.norm(C - ref) / torch.linalg.norm(ref)).item() < 1e-5
