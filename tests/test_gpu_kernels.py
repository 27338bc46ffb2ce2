"""-m gpu: kernel-level parity of the sm_100a kernels (through the ozk_* C-ABI) against the CPU
oracle: split slices / max exponents bit-exact, tcgen05 int8 products exact, mantissa-loss totals exact."""
import numpy as np
import pytest
import torch

import oracle_lib
import ozimmu_b200 as oz
from gpu_util import bits, block, split_product, stream_ptr, to_dev

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("col_major", [False, True])
@pytest.mark.parametrize("rows,length,ld_extra", [(1, 1, 0), (7, 5, 3), (64, 128, 0), (130, 257, 5), (33, 1000, 1),
                                                  (257, 4099, 0)])
@pytest.mark.parametrize("kind,num_split", [("urand01", 9), ("exp_rand-2", 13), ("mixed", 18), ("normal01", 3)])
def test_split_matches_oracle(col_major, rows, length, ld_extra, kind, num_split):
    # storage: col_major -> element (r, c) at c*ld + r (ld >= rows); else r*ld + c (ld >= length)
    ld = (rows if col_major else length) + ld_extra
    count = ld * (length if col_major else rows)
    x = oracle_lib.gen_matrix(kind, count, seed=rows * 7919 + length)
    L = oz.lib().ozk_bits_per_int8(length)
    want, want_mx = oracle_lib.oracle_split(x, ld, rows, length, col_major, num_split, L)
    got, got_mx = split_product(to_dev(x), ld, rows, length, col_major, num_split, L)
    assert np.array_equal(bits(got_mx), bits(want_mx))
    assert np.array_equal(got[:, :, :length], want[:, :, :length])
    assert not got[:, :, length:].any(), "padding bytes must be zero"


def test_split_special_rows():
    # all-zero row, all-subnormal row, row with Inf, row with huge dynamic range
    rows, length = 5, 48
    x = np.zeros((rows, length))
    x[1, :] = 5e-324 * np.arange(1, length + 1)
    x[2, :] = np.linspace(-3, 3, length)
    x[2, 7] = np.inf
    x[3, :] = np.ldexp(1.0, np.arange(length) * 20 - 480)
    x[4, :] = -np.ldexp(1.5, -np.arange(length))
    for col_major in (False, True):
        store = np.ascontiguousarray(x.T if col_major else x).ravel()
        ld = rows if col_major else length
        want, want_mx = oracle_lib.oracle_split(store, ld, rows, length, col_major, 9, 7)
        got, got_mx = split_product(to_dev(store), ld, rows, length, col_major, 9, 7)
        assert np.array_equal(bits(got_mx), bits(want_mx))
        assert np.array_equal(got[:, :, :length], want[:, :, :length])


# (0, BN): force the CTA-pair kernel's tile width; (64, 128): 128 x 128 tiles (64 rows per CTA, UMMA M = 128);
# (0, 0): per-problem choice
@pytest.mark.parametrize("shape", [(0, 256), (0, 240), (0, 224), (0, 208), (0, 192), (0, 128), (64, 128), (0, 0)])
@pytest.mark.parametrize("m,n,k", [(128, 128, 128), (256, 384, 512), (100, 60, 70), (513, 259, 1031), (1, 1, 1),
                                   (1025, 1023, 1024)])
def test_int8_pair_product_exact(shape, m, n, k):
    """tcgen05 kind::i8 product of one slice pair == exact integer GEMM (what cublasGemmEx int8
    returns at reference src/gemm.cu:315-329)."""
    L = oz.lib()
    rng = np.random.default_rng(m * 31 + n * 17 + k)
    pitch = int(L.ozk_slice_pitch(k))
    s = 3
    a = rng.integers(-127, 128, size=(s, m, k), dtype=np.int8)
    b = rng.integers(-127, 128, size=(s, n, k), dtype=np.int8)
    da, db = to_dev(block(a, pitch)), to_dev(block(b, pitch))
    out = torch.full((n, m), 12345, dtype=torch.int32, device="cuda")  # column-major m x n, ld = m
    L.ozk_set_cluster_shape(*shape)
    try:
        for (ai, bi) in [(1, 1), (2, 3), (3, 1)]:
            rc = L.ozk_gemm_i8_pair(m, n, k, da.data_ptr(), db.data_ptr(), pitch, s, ai, bi, out.data_ptr(), stream_ptr())
            assert rc == 0
            torch.cuda.synchronize()
            want = (a[ai - 1].astype(np.int64) @ b[bi - 1].astype(np.int64).T).T  # (n, m)
            assert np.array_equal(out.cpu().numpy().astype(np.int64), want), f"pair {(ai, bi)} cluster {shape}"
    finally:
        L.ozk_set_cluster_shape(0, 0)


@pytest.mark.parametrize("col_major", [False, True])
@pytest.mark.parametrize("rows,length", [(3, 5), (64, 256), (129, 1000), (300, 4100)])
@pytest.mark.parametrize("kind", ["urand01", "exp_rand-4", "mixed"])
def test_mantissa_loss_matches_oracle(col_major, rows, length, kind):
    ld = (rows if col_major else length) + 2
    count = ld * (length if col_major else rows)
    x = oracle_lib.gen_matrix(kind, count, seed=rows + length)
    L = oz.lib().ozk_bits_per_int8(length)
    want = np.zeros(16, dtype=np.uint64)
    oracle_lib.oracle().oz_mantissa_loss(want.ctypes.data, rows, length, x.ctypes.data, ld, int(col_major), L)
    cnt = torch.zeros(16, dtype=torch.int64, device="cuda")
    scratch = torch.zeros(rows, dtype=torch.int32, device="cuda")
    rc = oz.lib().ozk_mantissa_loss(cnt.data_ptr(), scratch.data_ptr(), rows, length, to_dev(x).data_ptr(), ld,
                                    int(col_major), L, stream_ptr())
    assert rc == 0
    torch.cuda.synchronize()
    assert np.array_equal(cnt.cpu().numpy().astype(np.uint64), want)
