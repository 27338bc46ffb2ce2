"""CPU: pins oracle/oz_oracle.c to the UNMODIFIED reference.  The fixtures under tests/golden/ were
produced by tests/golden/make_golden.py on a B200 from oracle/_ref/libozref.so (the reference's own
sources compiled as they lie): C = alpha*op(A)*op(B)+beta*C through mtk::ozimmu::gemm, the int8 slices
and row scales through the reference's split_int8<double>, the mantissa-loss counters and the
selected mode through mtk::ozimmu::auto_mode_select.  Everything is compared bit for bit."""
import glob
from pathlib import Path

import numpy as np
import pytest

import oracle_lib

GOLDEN = Path(__file__).resolve().parent / "golden"
GEMM_FILES = sorted(glob.glob(str(GOLDEN / "gemm_*.npz")))
AUTO_FILES = sorted(glob.glob(str(GOLDEN / "auto_*.npz")))
ZGEMM_FILES = sorted(glob.glob(str(GOLDEN / "zgemm_*.npz")))


def test_fixtures_present():
    assert len(GEMM_FILES) >= 6 and len(AUTO_FILES) >= 4


@pytest.mark.parametrize("path", GEMM_FILES, ids=lambda p: Path(p).stem)
def test_gemm_matches_reference_bits(path):
    g = np.load(path)
    op_a, op_b, m, n, k, s = (int(g[x]) for x in ("op_a", "op_b", "m", "n", "k", "num_split"))
    got = oracle_lib.oracle_gemm(op_a, op_b, m, n, k, float(g["alpha"]), g["a"], int(g["lda"]), g["b"], int(g["ldb"]),
                                 float(g["beta"]), g["c_in"], int(g["ldc"]), s)
    want = g["c_out"]
    assert np.array_equal(got.view(np.int64), want.view(np.int64))


@pytest.mark.parametrize("path", GEMM_FILES, ids=lambda p: Path(p).stem)
def test_split_matches_reference_bits(path):
    g = np.load(path)
    op_a, op_b, m, n, k, s, bits = (int(g[x]) for x in ("op_a", "op_b", "m", "n", "k", "num_split", "bits"))
    assert oracle_lib.oracle().oz_bits_per_int8(k) == bits
    a_sl, amax = oracle_lib.oracle_split(g["a"], int(g["lda"]), m, k, op_a == 0, s, bits)
    b_sl, bmax = oracle_lib.oracle_split(g["b"], int(g["ldb"]), n, k, op_b != 0, s, bits)
    assert np.array_equal(a_sl, g["a_slices"]) and np.array_equal(b_sl, g["b_slices"])
    assert np.array_equal(amax.view(np.int64), g["amax"].view(np.int64))
    assert np.array_equal(bmax.view(np.int64), g["bmax"].view(np.int64))


@pytest.mark.parametrize("path", AUTO_FILES, ids=lambda p: Path(p).stem)
def test_auto_mode_matches_reference(path):
    g = np.load(path)
    op_a, op_b, m, n, k = (int(g[x]) for x in ("op_a", "op_b", "m", "n", "k"))
    for thr, ref_mode in zip(g["thresholds"], g["modes"]):
        s, cnt = oracle_lib.oracle_auto_select(op_a, op_b, m, n, k, g["a"], int(g["lda"]), g["b"], int(g["ldb"]),
                                               float(thr))
        # the reference owns only the 8 counters of fp64_int8_3..10
        assert np.array_equal(cnt[:8], g["counters8"])
        # compute_mode_t: fp64_int8_S == S - 1, dgemm == 1.  The reference's choice is defined when it
        # lands in fp64_int8_3..10; beyond that it reads uninitialised counters.
        if 2 <= int(ref_mode) <= 9:
            assert s - 1 == int(ref_mode)
        elif s != 0:
            assert s >= 11


@pytest.mark.parametrize("path", ZGEMM_FILES, ids=lambda p: Path(p).stem)
def test_zgemm_matches_reference_bits(path):
    """complex path (reference src/gemm.cu:412-521), incl. the reference's beta pre-scale as compiled"""
    g = np.load(path)
    op_a, op_b, m, n, k, s = (int(g[x]) for x in ("op_a", "op_b", "m", "n", "k", "num_split"))
    got = oracle_lib.oracle_gemm_complex(op_a, op_b, m, n, k, complex(g["alpha"]), g["a"], int(g["lda"]), g["b"],
                                         int(g["ldb"]), complex(g["beta"]), g["c_in"], int(g["ldc"]), s,
                                         reference_beta_quirk=True)
    want = g["c_out"]
    assert np.array_equal(got.view(np.int64), want.view(np.int64))
    # the corrected beta pre-scale (what the product computes) differs from the reference only when Im(beta) != 0,
    # and there it is the one that agrees with exact complex arithmetic
    fixed = oracle_lib.oracle_gemm_complex(op_a, op_b, m, n, k, complex(g["alpha"]), g["a"], int(g["lda"]), g["b"],
                                           int(g["ldb"]), complex(g["beta"]), g["c_in"], int(g["ldc"]), s)
    beta = complex(g["beta"])
    if beta.imag == 0:
        assert np.array_equal(fixed.view(np.int64), want.view(np.int64))
    elif s >= 9:
        lda, ldb, ldc = int(g["lda"]), int(g["ldb"]), int(g["ldc"])
        A = g["a"].reshape(-1, lda).T[: (m if op_a == 0 else k)]
        B = g["b"].reshape(-1, ldb).T[: (k if op_b == 0 else n)]
        A = A if op_a == 0 else A.T
        B = B if op_b == 0 else B.T
        exact = complex(g["alpha"]) * (A @ B) + beta * g["c_in"].reshape(-1, ldc).T[:m]
        err_fixed = np.abs(fixed.reshape(-1, ldc).T[:m] - exact).max()
        err_ref = np.abs(want.reshape(-1, ldc).T[:m] - exact).max()
        assert err_fixed < 1e-10 * np.abs(exact).max() and err_fixed < err_ref


def test_zgemm_fixtures_present():
    assert len(ZGEMM_FILES) >= 4
