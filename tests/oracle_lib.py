"""ctypes access to the parity checkers under oracle/ (test infrastructure only).

  oracle()    -> oracle/liboz_oracle.so   plain-C CPU restatement (oracle/oz_oracle.c)
  reference() -> oracle/_ref/libozref.so  the UNMODIFIED reference ozIMMU + ozref_* doorway
                                          (oracle/ref_shim.cu); needs a GPU to run.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_SO = ROOT / "oracle" / "liboz_oracle.so"
REF_SO = ROOT / "oracle" / "_ref" / "libozref.so"

_oracle = None
_ref = None

sz, i32, u32, vp, dbl = C.c_size_t, C.c_int, C.c_uint, C.c_void_p, C.c_double


def build_oracle() -> None:
    if not ORACLE_SO.exists() or ORACLE_SO.stat().st_mtime < (ROOT / "oracle" / "oz_oracle.c").stat().st_mtime:
        subprocess.run(["make", "-C", str(ROOT / "oracle"), "liboz_oracle.so"], check=True, capture_output=True)


def oracle() -> C.CDLL:
    global _oracle
    if _oracle is None:
        build_oracle()
        L = C.CDLL(str(ORACLE_SO))
        L.oz_bits_per_int8.restype, L.oz_bits_per_int8.argtypes = C.c_uint32, [C.c_uint32]
        L.oz_pair_list.restype, L.oz_pair_list.argtypes = i32, [i32, vp, vp]
        L.oz_slice_ld.restype, L.oz_slice_ld.argtypes = C.c_uint32, [C.c_uint32]
        L.oz_split.restype, L.oz_split.argtypes = None, [vp, C.c_uint32, vp, sz, sz, vp, sz, i32, u32, u32]
        L.oz_int8_gemm.restype, L.oz_int8_gemm.argtypes = None, [vp, sz, sz, sz, vp, vp]
        L.oz_gemm.restype, L.oz_gemm.argtypes = i32, [i32, i32, sz, sz, sz, dbl, vp, sz, vp, sz, dbl, vp, sz, u32,
                                                      vp, vp, vp, vp]
        L.oz_mantissa_loss.restype, L.oz_mantissa_loss.argtypes = None, [vp, sz, sz, vp, sz, i32, u32]
        L.oz_gemm_complex_q.restype, L.oz_gemm_complex_q.argtypes = i32, [i32, i32, sz, sz, sz, vp, vp, sz, vp, sz, vp, vp,
                                                                          sz, u32, i32]
        L.oz_gemm_complex.restype, L.oz_gemm_complex.argtypes = i32, [i32, i32, sz, sz, sz, vp, vp, sz, vp, sz, vp, vp,
                                                                      sz, u32]
        L.oz_auto_select_complex.restype = i32
        L.oz_auto_select_complex.argtypes = [i32, i32, sz, sz, sz, vp, sz, vp, sz, dbl, vp]
        L.oz_auto_select.restype, L.oz_auto_select.argtypes = i32, [i32, i32, sz, sz, sz, vp, sz, vp, sz, dbl, vp]
        _oracle = L
    return _oracle


def reference():
    global _ref
    if _ref is None:
        if not REF_SO.exists():
            return None
        L = C.CDLL(str(REF_SO), mode=os.RTLD_LOCAL | os.RTLD_NOW)
        L.ozref_create.restype, L.ozref_create.argtypes = i32, [C.POINTER(vp)]
        L.ozref_destroy.restype, L.ozref_destroy.argtypes = i32, [vp]
        L.ozref_set_stream.restype, L.ozref_set_stream.argtypes = i32, [vp, vp]
        L.ozref_gemm.restype, L.ozref_gemm.argtypes = i32, [vp, i32, i32, sz, sz, sz, vp, vp, sz, vp, sz, vp, vp, sz,
                                                            i32, i32]
        L.ozref_split_int8.restype, L.ozref_split_int8.argtypes = i32, [vp, C.c_uint32, vp, sz, sz, vp, sz, i32, i32,
                                                                        u32, u32, vp]
        L.ozref_auto_mode_select.restype = i32
        L.ozref_auto_mode_select.argtypes = [vp, i32, i32, sz, sz, sz, vp, sz, vp, sz, i32, dbl, vp]
        L.ozref_bits_per_int8.restype, L.ozref_bits_per_int8.argtypes = u32, [u32]
        L.ozref_reallocate.restype, L.ozref_reallocate.argtypes = sz, [vp, i32, i32, sz, sz, sz, i32, i32]
        L.ozref_profiling.restype, L.ozref_profiling.argtypes = None, [vp, i32]
        L.ozref_print_profile.restype, L.ozref_print_profile.argtypes = None, [vp, C.c_char_p]
        _ref = L
    return _ref


# ---- numpy front-ends of the CPU oracle ---------------------------------------------------------
def _p(a: np.ndarray) -> int:
    return a.ctypes.data


def oracle_split(x: np.ndarray, ld: int, rows: int, length: int, col_major: bool, num_split: int, bits: int):
    """-> (slices[num_split, rows, ld4] int8, max_exp[rows] f64); x is the flat FP64 storage."""
    L = oracle()
    ld4 = int(L.oz_slice_ld(length))
    out = np.zeros((num_split, rows, ld4), dtype=np.int8)
    mx = np.zeros(rows, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    L.oz_split(_p(out), ld4, _p(mx), rows, length, _p(x), ld, int(col_major), num_split, bits)
    return out, mx


def oracle_gemm(op_a: int, op_b: int, m: int, n: int, k: int, alpha: float, a: np.ndarray, lda: int, b: np.ndarray,
                ldb: int, beta: float, c: np.ndarray, ldc: int, num_split: int) -> np.ndarray:
    """C (flat column-major storage, ldc) updated by the CPU restatement; returns a new array."""
    L = oracle()
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    out = np.array(c, dtype=np.float64, copy=True)
    rc = L.oz_gemm(op_a, op_b, m, n, k, alpha, _p(a), lda, _p(b), ldb, beta, _p(out), ldc, num_split, None, None,
                   None, None)
    assert rc == 0
    return out


def oracle_gemm_complex(op_a: int, op_b: int, m: int, n: int, k: int, alpha: complex, a: np.ndarray, lda: int,
                        b: np.ndarray, ldb: int, beta: complex, c: np.ndarray, ldc: int, num_split: int,
                        reference_beta_quirk: bool = False) -> np.ndarray:
    """a, b, c: complex128 flat column-major storage (ld in complex elements); returns the new C.
    reference_beta_quirk=True reproduces the reference's aliasing bug in C = beta*C (Im(beta) != 0 only; SURVEY
    App. B.6) -- that is what its golden vectors contain; the product ships the corrected arithmetic (False)."""
    L = oracle()
    a = np.ascontiguousarray(a, dtype=np.complex128)
    b = np.ascontiguousarray(b, dtype=np.complex128)
    out = np.array(c, dtype=np.complex128, copy=True)
    al = np.array([alpha.real, alpha.imag], dtype=np.float64)
    be = np.array([beta.real, beta.imag], dtype=np.float64)
    rc = L.oz_gemm_complex_q(op_a, op_b, m, n, k, _p(al), _p(a), lda, _p(b), ldb, _p(be), _p(out), ldc, num_split,
                             int(reference_beta_quirk))
    assert rc == 0
    return out


def gen_complex(kind: str, count: int, seed: int) -> np.ndarray:
    return gen_matrix(kind, count, seed) + 1j * gen_matrix(kind, count, seed + 7919)


def oracle_auto_select(op_a: int, op_b: int, m: int, n: int, k: int, a: np.ndarray, lda: int, b: np.ndarray, ldb: int,
                       threshold: float):
    """-> (num_split or 0 for dgemm, counters[16] uint64)"""
    L = oracle()
    cnt = np.zeros(16, dtype=np.uint64)
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    s = L.oz_auto_select(op_a, op_b, m, n, k, _p(a), lda, _p(b), ldb, threshold, _p(cnt))
    return int(s), cnt


# ---- deterministic inputs (reference test/main_test.cu:56-80,195-212 distributions) -----------------
def gen_matrix(kind: str, count: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    if kind == "urand01":
        return 1.0 - rng.random(count)  # (0, 1]
    if kind == "normal01":
        return rng.standard_normal(count)
    if kind.startswith("exp_rand-"):
        phi = float(kind.split("-", 1)[1])
        return (rng.random(count) - 0.5) * np.exp(phi * rng.standard_normal(count))
    if kind == "mixed":  # zeros, negatives, tiny and huge magnitudes, a few subnormals
        x = (rng.random(count) - 0.5) * np.exp(6.0 * rng.standard_normal(count))
        x[rng.random(count) < 0.05] = 0.0
        x[rng.random(count) < 0.01] = 5e-324 * 12345
        return x
    raise ValueError(kind)
