"""CPU: properties of the oracle itself (BASELINE config 1: 256^3 correctness vs host OpenBLAS dgemm)
and of the host-side helpers that need no GPU."""
import numpy as np
import pytest

import oracle_lib


def colmajor(x):
    """2-D array -> flat column-major storage"""
    return np.asfortranarray(x).ravel(order="F")


@pytest.mark.parametrize("op_a,op_b", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("num_split,tol", [(9, 1e-15), (13, 1e-15), (18, 1e-15), (3, 1e-3)])
def test_config1_256_vs_openblas(op_a, op_b, num_split, tol):
    """256 x 256 x 256, uniform(0,1]: relative residual vs numpy (OpenBLAS) dgemm below the reference's
    CI gate (test/main_test.cu:744: < 1e-15 for fp64_int8_8..16)."""
    m = n = k = 256
    rng = np.random.default_rng(0)
    A = 1.0 - rng.random((m, k))
    B = 1.0 - rng.random((k, n))
    a = colmajor(A if op_a == 0 else A.T)
    b = colmajor(B if op_b == 0 else B.T)
    lda = m if op_a == 0 else k
    ldb = k if op_b == 0 else n
    c = oracle_lib.oracle_gemm(op_a, op_b, m, n, k, 1.0, a, lda, b, ldb, 0.0, np.zeros(m * n), m, num_split)
    C = c.reshape(n, m).T
    ref = A @ B
    resid = np.linalg.norm(C - ref) / np.linalg.norm(ref)
    assert resid < tol


def test_alpha_beta_and_ld():
    m, n, k, ld = 37, 29, 61, 5
    rng = np.random.default_rng(1)
    A = rng.standard_normal((m, k))
    B = rng.standard_normal((k, n))
    C0 = rng.standard_normal((m, n))
    a = np.zeros((k, m + ld)); a[:, :m] = A.T
    b = np.zeros((n, k + ld)); b[:, :k] = B.T
    c = np.zeros((n, m + ld)); c[:, :m] = C0.T
    out = oracle_lib.oracle_gemm(0, 0, m, n, k, -1.25, a.ravel(), m + ld, b.ravel(), k + ld, 0.5, c.ravel(), m + ld, 13)
    got = out.reshape(n, m + ld)[:, :m].T
    want = -1.25 * (A @ B) + 0.5 * C0
    assert np.allclose(got, want, rtol=1e-12, atol=1e-12)
    assert not out.reshape(n, m + ld)[:, m:].any(), "ld padding of C must stay untouched"


def test_split_reconstructs_value():
    """Invariant of SURVEY App. A.3: a = 2*S * sum_t slice[t] * 2^(-L(t+1)) up to the truncation."""
    rng = np.random.default_rng(2)
    rows, length, s = 8, 200, 9
    x = (rng.random((rows, length)) - 0.5) * np.exp(2.0 * rng.standard_normal((rows, length)))
    sl, mx = oracle_lib.oracle_split(x.ravel(), length, rows, length, False, s, 7)
    acc = np.zeros((rows, length))
    for t in range(s):
        acc += sl[t, :, :length].astype(np.float64) * 2.0 ** (-7 * (t + 1))
    recon = 2.0 * mx[:, None] * acc
    bound = 2.0 * mx[:, None] * 2.0 ** (-7 * s)
    assert (np.abs(recon - x) <= bound).all()
    assert np.abs(sl[0]).max() <= 63 and np.abs(sl[1:]).max() <= 127


@pytest.mark.parametrize("k,bits", [(1, 7), (256, 7), (8192, 7), (131072, 7), (131073, 6), (524288, 6), (524289, 5),
                                    (1 << 20, 5)])
def test_bits_per_int8(k, bits):
    """reference src/split.cu:520-536; the product library's host helper must agree with the oracle"""
    import ozimmu_b200 as oz
    assert oracle_lib.oracle().oz_bits_per_int8(k) == bits
    assert oz.get_bits_per_int8(k) == bits


def test_pair_order():
    """reference src/config.cu:85-92: diagonal-major pair list"""
    import ctypes as C
    L = oracle_lib.oracle()
    a = (C.c_int * 200)()
    b = (C.c_int * 200)()
    n = L.oz_pair_list(4, a, b)
    assert [(a[i], b[i]) for i in range(n)] == [(1, 1), (1, 2), (2, 1), (1, 3), (2, 2), (3, 1), (1, 4), (2, 3), (3, 2),
                                                  (4, 1)]
    for s in range(3, 19):
        assert L.oz_pair_list(s, None, None) == s * (s + 1) // 2


def test_auto_select_monotone():
    rng = np.random.default_rng(3)
    m = n = k = 64
    a = (rng.random(m * k) - 0.5) * np.exp(2 * rng.standard_normal(m * k))
    b = (rng.random(k * n) - 0.5) * np.exp(2 * rng.standard_normal(k * n))
    prev = 99
    for thr in (0.0, 0.5, 1.0, 2.0, 4.0, 8.0, 100.0):
        s, cnt = oracle_lib.oracle_auto_select(0, 0, m, n, k, a, m, b, k, thr)
        s_eff = 19 if s == 0 else s
        assert s_eff <= prev
        prev = s_eff
        assert all(cnt[i] >= cnt[i + 1] for i in range(15))
    assert s == 3
