"""-m gpu: strided-batched DGEMMs (reference src/cublas.cu:315-492 loops one Ozaki GEMM per entry).  The grouped
launch must give every entry the bits of a separate ozimmu_gemm call, for every tile width, ragged sizes,
padded strides, stride-0 (shared) operands, and under LD_PRELOAD through torch.bmm."""
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


@pytest.fixture(scope="module")
def handle():
    h = oz.create()
    yield h
    oz.destroy(h)


def _case(op_a, op_b, m, n, k, batch, pad, seed, share_b=False):
    lda = (m if op_a == 0 else k) + pad
    ldb = (k if op_b == 0 else n) + pad
    ldc = m + pad
    sa = lda * (k if op_a == 0 else m) + 3 * pad
    sb = 0 if share_b else ldb * (n if op_b == 0 else k) + pad
    sc = ldc * n + 5 * pad
    a = oracle_lib.gen_matrix("exp_rand-1", sa * batch, seed)
    b = oracle_lib.gen_matrix("exp_rand-1", max(sb, ldb * (n if op_b == 0 else k)) * (1 if share_b else batch), seed + 1)
    c = oracle_lib.gen_matrix("normal01", sc * batch, seed + 2)
    return a, lda, sa, b, ldb, sb, c, ldc, sc


@pytest.mark.parametrize("op_a,op_b", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("m,n,k,batch,pad,beta,num_split", [
    (256, 256, 256, 5, 0, 0.0, 9),
    (300, 200, 520, 7, 2, -0.5, 8),      # ragged tiles, padded leading dimensions and strides
    (130, 700, 129, 3, 1, 2.0, 13),
])
def test_batched_equals_entry_by_entry(handle, op_a, op_b, m, n, k, batch, pad, beta, num_split):
    a, lda, sa, b, ldb, sb, c, ldc, sc = _case(op_a, op_b, m, n, k, batch, pad, 7)
    da, db = to_dev(a), to_dev(b)
    want, got = to_dev(c), to_dev(c)
    mode = oz.fp64_int8(num_split)
    for e in range(batch):
        assert oz.gemm(handle, op_a, op_b, m, n, k, 1.5, da[e * sa:], lda, db[e * sb:], ldb, beta, want[e * sc:], ldc,
                       mode) == 0
    assert oz.gemm_strided_batched(handle, op_a, op_b, m, n, k, 1.5, da, lda, sa, db, ldb, sb, beta, got, ldc, sc,
                                   batch, mode) == 0
    torch.cuda.synchronize()
    assert torch.equal(got.view(torch.int64), want.view(torch.int64))   # incl. the untouched gaps between entries


@pytest.mark.parametrize("shape", [(0, 256), (0, 128), (64, 128)])
def test_batched_tile_widths_and_shared_operand(handle, shape):
    """both tile widths of the grouped kernel; B shared by all entries (stride 0); more tiles than SM pairs"""
    m, n, k, batch = 512, 768, 384, 40
    a, lda, sa, b, ldb, sb, c, ldc, sc = _case(0, 0, m, n, k, batch, 0, 11, share_b=True)
    da, db = to_dev(a), to_dev(b)
    want, got = to_dev(c), to_dev(c)
    for e in range(batch):
        assert oz.gemm(handle, 0, 0, m, n, k, 1.0, da[e * sa:], lda, db, ldb, 0.0, want[e * sc:], ldc, oz.fp64_int8(9)) == 0
    oz.lib().ozk_set_cluster_shape(*shape)
    try:
        assert oz.gemm_strided_batched(handle, 0, 0, m, n, k, 1.0, da, lda, sa, db, ldb, 0, 0.0, got, ldc, sc, batch,
                                       oz.fp64_int8(9)) == 0
        torch.cuda.synchronize()
    finally:
        oz.lib().ozk_set_cluster_shape(0, 0)
    assert torch.equal(got.view(torch.int64), want.view(torch.int64))


def test_batched_chunks_auto_dgemm_and_errors(handle, monkeypatch):
    m, n, k, batch = 256, 256, 300, 6
    a, lda, sa, b, ldb, sb, c, ldc, sc = _case(0, 1, m, n, k, batch, 0, 13)
    da, db = to_dev(a), to_dev(b)
    want = to_dev(c)
    for e in range(batch):
        assert oz.gemm(handle, 0, 1, m, n, k, 1.0, da[e * sa:], lda, db[e * sb:], ldb, 1.0, want[e * sc:], ldc,
                       oz.fp64_int8(9)) == 0
    # a workspace cap below two entries: the batch is processed in chunks of one
    monkeypatch.setenv("OZIMMU_B200_BATCH_WORKSPACE_MB", "1")
    got = to_dev(c)
    assert oz.gemm_strided_batched(handle, 0, 1, m, n, k, 1.0, da, lda, sa, db, ldb, sb, 1.0, got, ldc, sc, batch,
                                   oz.fp64_int8(9)) == 0
    torch.cuda.synchronize()
    assert torch.equal(got.view(torch.int64), want.view(torch.int64))
    monkeypatch.delenv("OZIMMU_B200_BATCH_WORKSPACE_MB")
    # auto mode: one counter pass for the whole batch, the split count still chosen per entry
    oz.set_auto_mantissa_loss_threashold(handle, 1.0)
    got_auto, want_auto = to_dev(c), to_dev(c)
    for e in range(batch):
        assert oz.gemm(handle, 0, 1, m, n, k, 1.0, da[e * sa:], lda, db[e * sb:], ldb, 0.0, want_auto[e * sc:], ldc,
                       oz.compute_mode_t.fp64_int8_auto) == 0
    assert oz.gemm_strided_batched(handle, 0, 1, m, n, k, 1.0, da, lda, sa, db, ldb, sb, 0.0, got_auto, ldc, sc, batch,
                                   oz.compute_mode_t.fp64_int8_auto) == 0
    torch.cuda.synchronize()
    assert torch.equal(got_auto.view(torch.int64), want_auto.view(torch.int64))
    # dgemm passthrough
    got_d = to_dev(c)
    assert oz.gemm_strided_batched(handle, 0, 1, m, n, k, 1.0, da, lda, sa, db, ldb, sb, 0.0, got_d, ldc, sc, batch,
                                   oz.compute_mode_t.dgemm) == 0
    torch.cuda.synchronize()
    ref = torch.stack([da[e * sa:e * sa + m * k].view(k, m).T @ db[e * sb:e * sb + n * k].view(k, n)
                       for e in range(batch)])
    out = torch.stack([got_d[e * sc:e * sc + m * n].view(n, m).T for e in range(batch)])
    assert torch.allclose(out, ref, rtol=1e-12, atol=1e-12)
    # overlapping entries of C are rejected (they would be written concurrently); k == 0 scales C
    assert oz.gemm_strided_batched(handle, 0, 1, m, n, k, 1.0, da, lda, sa, db, ldb, sb, 0.0, got, ldc, ldc * n - 1,
                                   batch, oz.fp64_int8(9)) == 1
    z = to_dev(c)
    assert oz.gemm_strided_batched(handle, 0, 1, m, n, 0, 1.0, da, lda, sa, db, ldb, sb, 2.0, z, ldc, sc, batch,
                                   oz.fp64_int8(9)) == 0
    torch.cuda.synchronize()
    assert torch.equal(z.view(torch.int64), (to_dev(c) * 2.0).view(torch.int64))


def test_batched_matches_oracle_and_reference(handle):
    """the grouped launch against the CPU oracle entry by entry (small), and against the unmodified reference looping
    over the entries as its interposer does (src/cublas.cu:380-406) -- not only against this library's own gemm()"""
    m, n, k, batch, pad, s = 96, 80, 200, 4, 1, 9
    for op_a, op_b in ((0, 0), (1, 1)):
        a, lda, sa, b, ldb, sb, c, ldc, sc = _case(op_a, op_b, m, n, k, batch, pad, 21)
        got = to_dev(c)
        assert oz.gemm_strided_batched(handle, op_a, op_b, m, n, k, -0.75, to_dev(a), lda, sa, to_dev(b), ldb, sb, 0.5, got,
                                       ldc, sc, batch, oz.fp64_int8(s)) == 0
        torch.cuda.synchronize()
        got = got.cpu().numpy()
        for e in range(batch):
            want = oracle_lib.oracle_gemm(op_a, op_b, m, n, k, -0.75, a[e * sa:], lda, b[e * sb:], ldb, 0.5,
                                          c[e * sc:e * sc + ldc * n], ldc, s)
            assert np.array_equal(bits(got[e * sc:e * sc + ldc * n]), bits(want)), (op_a, op_b, e)
    if oracle_lib.reference() is None:
        pytest.skip("oracle/_ref/libozref.so not built")
    ref = Reference()
    try:
        m, n, k, batch = 1024, 1024, 1028, 5     # m % 4 == 0: the reference's int8 cublasGemmEx needs it on this cuBLAS
        a, lda, sa, b, ldb, sb, c, ldc, sc = _case(0, 0, m, n, k, batch, 0, 23)
        da, db = to_dev(a), to_dev(b)
        want, got = to_dev(c), to_dev(c)
        for e in range(batch):
            ref.gemm(0, 0, m, n, k, 1.0, da[e * sa:], lda, db[e * sb:], ldb, -1.0, want[e * sc:], ldc, 9 - 1)
        assert oz.gemm_strided_batched(handle, 0, 0, m, n, k, 1.0, da, lda, sa, db, ldb, sb, -1.0, got, ldc, sc, batch,
                                       oz.fp64_int8(9)) == 0
        torch.cuda.synchronize()
        assert torch.equal(got.view(torch.int64), want.view(torch.int64))
    finally:
        ref.close()


@pytest.mark.parametrize("op_a,op_b", [(0, 0), (1, 0), (1, 1)])
def test_complex_batched_one_launch(handle, op_a, op_b):
    """a complex strided batch: one grouped launch, every tile running its four plane products; equal to the oracle
    (small) and to per-entry complex gemm() calls; exactly one product launch for the whole batch"""
    m, n, k, batch, s = 72, 40, 130, 3, 9
    lda, ldb = (m if op_a == 0 else k), (k if op_b == 0 else n)
    sa, sb, sc = m * k + 5, k * n + 3, m * n + 7
    a = oracle_lib.gen_complex("exp_rand-1", sa * batch, 31)
    b = oracle_lib.gen_complex("exp_rand-1", sb * batch, 32)
    c = oracle_lib.gen_complex("normal01", sc * batch, 33)
    alpha, beta = 0.5 - 1.5j, -0.25 + 2.0j
    da, db, got, loop = to_dev(a), to_dev(b), to_dev(c), to_dev(c)
    before = oz.launch_count()
    assert oz.gemm_strided_batched(handle, op_a, op_b, m, n, k, alpha, da, lda, sa, db, ldb, sb, beta, got, m, sc, batch,
                                   oz.fp64_int8(s), oz.complx) == 0
    launches = oz.launch_count() - before
    for e in range(batch):
        assert oz.gemm(handle, op_a, op_b, m, n, k, alpha, da[e * sa:], lda, db[e * sb:], ldb, beta, loop[e * sc:], m,
                       oz.fp64_int8(s), oz.complx) == 0
    torch.cuda.synchronize()
    assert torch.equal(torch.view_as_real(got).view(torch.int64), torch.view_as_real(loop).view(torch.int64))
    g = got.cpu().numpy()
    for e in range(batch):
        want = oracle_lib.oracle_gemm_complex(op_a, op_b, m, n, k, alpha, a[e * sa:], lda, b[e * sb:], ldb, beta,
                                              c[e * sc:e * sc + m * n], m, s)
        assert np.array_equal(g[e * sc:e * sc + m * n].view(np.int64), want.view(np.int64)), e
    # splits: <= 2 planes x 2 operands x 2 kernels; products: ONE launch for 3 entries x 4 plane products
    assert launches <= 9, launches


def test_batched_auto_mode_one_pass(handle):
    """fp64_int8_auto over a batch whose entries need DIFFERENT split counts (entries scaled to different dynamic
    ranges): per-entry choice from one counter pass == gemm(auto) entry by entry, bit for bit"""
    m, n, k, batch = 256, 192, 320, 6
    sa, sb, sc = m * k, k * n, m * n
    a = np.concatenate([oracle_lib.gen_matrix(f"exp_rand-{phi}", sa, 40 + i)
                        for i, phi in enumerate((0, 0, 2, 2, 4, 0))])
    b = np.concatenate([oracle_lib.gen_matrix(f"exp_rand-{phi}", sb, 50 + i)
                        for i, phi in enumerate((0, 0, 2, 2, 4, 0))])
    da, db = to_dev(a), to_dev(b)
    got = torch.zeros(sc * batch, dtype=torch.float64, device="cuda")
    want = torch.zeros_like(got)
    oz.set_auto_mantissa_loss_threashold(handle, 1.0)
    modes = []
    for e in range(batch):
        modes.append(oz.auto_mode_select(handle, 0, 0, m, n, k, da[e * sa:], m, db[e * sb:], k, oz.real, 1.0))
        assert oz.gemm(handle, 0, 0, m, n, k, 1.0, da[e * sa:], m, db[e * sb:], k, 0.0, want[e * sc:], m,
                       oz.compute_mode_t.fp64_int8_auto) == 0
    assert len(set(modes)) >= 2, modes     # the batch really mixes split counts
    assert oz.gemm_strided_batched(handle, 0, 0, m, n, k, 1.0, da, m, sa, db, k, sb, 0.0, got, m, sc, batch,
                                   oz.compute_mode_t.fp64_int8_auto) == 0
    torch.cuda.synchronize()
    assert torch.equal(got.view(torch.int64), want.view(torch.int64))
    oz.set_auto_mantissa_loss_threashold(handle, 0.0)


DROPIN_BMM = r"""
import os, torch
torch.manual_seed(0)
a = torch.rand(6, 1024, 1152, dtype=torch.float64, device="cuda")
b = torch.rand(6, 1152, 1280, dtype=torch.float64, device="cuda")
c = torch.bmm(a, b)
torch.cuda.synchronize()
za = torch.randn(3, 1024, 1024, dtype=torch.complex128, device="cuda")
zb = torch.randn(3, 1024, 1024, dtype=torch.complex128, device="cuda")
zc = torch.bmm(za, zb)
torch.cuda.synchronize()
torch.save({"a": a.cpu(), "b": b.cpu(), "c": c.cpu(), "za": za.cpu(), "zb": zb.cpu(), "zc": zc.cpu()},
           os.environ["OZ_DROPIN_OUT"])
"""


def test_ld_preload_bmm(tmp_path, handle):
    """torch.bmm on float64 -> cublasDgemmStridedBatched / cublasGemmStridedBatchedEx -> one grouped launch;
    bit-identical to per-entry ozimmu_gemm calls on the same operands."""
    out = tmp_path / "bmm.pt"
    env = dict(os.environ, LD_PRELOAD=str(oz.LIB_PATH), OZIMMU_COMPUTE_MODE="fp64_int8_9",
               OZIMMU_ENABLE_CULIP_PROFILING="1", OZ_DROPIN_OUT=str(out))
    p = subprocess.run([sys.executable, "-c", DROPIN_BMM], env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    assert "[CULiP Result][Dfp64_int8_9-batched6-" in p.stdout, p.stdout[-2000:]
    d = torch.load(out)
    a, b = d["a"].cuda(), d["b"].cuda()
    bt, m, k = a.shape
    n = b.shape[2]
    # row-major C_e = A_e @ B_e  <=>  column-major C_e^T (n x m) = B_e^T (n x k) * A_e^T (k x m)
    c = torch.zeros(bt, m, n, dtype=torch.float64, device="cuda")
    for e in range(bt):
        assert oz.gemm(handle, 0, 0, n, m, k, 1.0, b[e], n, a[e], k, 0.0, c[e], n, oz.fp64_int8(9)) == 0
    torch.cuda.synchronize()
    assert np.array_equal(bits(c), bits(d["c"]))
    ref = d["a"] @ d["b"]
    assert (torch.linalg.norm(d["c"] - ref) / torch.linalg.norm(ref)).item() < 1e-15
    # complex128 bmm -> cublasZgemmStridedBatched / GemmStridedBatchedEx(C_64F): one grouped complex launch (the reference:
    # one complex Ozaki GEMM per entry, src/cublas.cu:380-406,494-512)
    assert "[CULiP Result][Zfp64_int8_9-batched3-" in p.stdout, p.stdout[-2000:]
    za, zb = d["za"].cuda(), d["zb"].cuda()
    zc = torch.zeros_like(za)
    nz = za.shape[1]
    for e in range(za.shape[0]):
        assert oz.gemm(handle, 0, 0, nz, nz, nz, 1.0 + 0j, zb[e], nz, za[e], nz, 0j, zc[e], nz, oz.fp64_int8(9),
                       oz.complx) == 0
    torch.cuda.synchronize()
    assert torch.equal(torch.view_as_real(zc).cpu().view(torch.int64), torch.view_as_real(d["zc"]).view(torch.int64))
