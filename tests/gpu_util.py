"""Helpers for the -m gpu parity tests: run the product (C-ABI of libozimmu.so) and the reference
(oracle/_ref/libozref.so) on torch CUDA tensors."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

import ozimmu_b200 as oz
import oracle_lib


def to_dev(x: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def stream_ptr() -> int:
    return int(torch.cuda.current_stream().cuda_stream)


def bits(x) -> np.ndarray:
    """FP64 array -> its int64 bit patterns (NaN-safe exact comparison)."""
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(x, dtype=np.float64).view(np.int64)


def ulp_distance(x, y) -> int:
    """max distance in units in the last place between two FP64 arrays (finite values)."""
    a, b = bits(x).copy(), bits(y).copy()
    a = np.where(a < 0, np.int64(-(2**63)) - a, a)
    b = np.where(b < 0, np.int64(-(2**63)) - b, b)
    d = np.abs(a.astype(np.float64) - b.astype(np.float64))
    return int(d.max()) if d.size else 0


def blocked_index(rows: int, pitch: int) -> np.ndarray:
    """byte offset inside one slice plane of element (r, c) in the blocked, pre-swizzled slice layout
    (include/ozimmu_b200.h): [row tile][k block][128 rows][128 B], 16-byte chunk c16 of row r at c16 ^ (r & 7)"""
    r = np.arange(rows, dtype=np.int64)[:, None]
    c = np.arange(pitch, dtype=np.int64)[None, :]
    kb = pitch // 128
    g = c >> 4
    return ((r >> 7) * kb + (g >> 3)) * 16384 + (r & 127) * 128 + (((g & 7) ^ (r & 7)) << 4) + (c & 15)


def plane_bytes(rows: int, pitch: int) -> int:
    return ((rows + 255) // 256) * 256 * pitch


def unblock(buf: np.ndarray, num_split: int, rows: int, pitch: int) -> np.ndarray:
    """flat blocked int8 buffer -> [num_split, rows, pitch]; also checks that all padding rows are zero"""
    plane = plane_bytes(rows, pitch)
    b = np.ascontiguousarray(buf).reshape(-1).view(np.int8)[:num_split * plane].reshape(num_split, plane)
    idx = blocked_index(rows, pitch)
    out = b[:, idx.reshape(-1)].reshape(num_split, rows, pitch)
    mask = np.ones(plane, dtype=bool)
    mask[idx.reshape(-1)] = False
    assert not b[:, mask].any(), "row padding of the slice planes must be zero"
    return out


def block(slices: np.ndarray, pitch: int) -> np.ndarray:
    """[num_split, rows, k] int8 -> flat blocked buffer (zero padded)"""
    s, rows, k = slices.shape
    plane = plane_bytes(rows, pitch)
    padded = np.zeros((s, rows, pitch), dtype=np.int8)
    padded[:, :, :k] = slices
    out = np.zeros((s, plane), dtype=np.int8)
    out[:, blocked_index(rows, pitch).reshape(-1)] = padded.reshape(s, -1)
    return out.reshape(-1)


def split_product(x: torch.Tensor, ld: int, rows: int, length: int, col_major: bool, num_split: int, bits_: int):
    """ozk_split_int8 -> (slices[num_split, rows, pitch] int8 numpy array in plain layout, max_exp[rows] tensor)"""
    L = oz.lib()
    pitch = int(L.ozk_slice_pitch(length))
    nbytes = int(L.ozk_slices_bytes(rows, length, num_split))
    assert nbytes == num_split * plane_bytes(rows, pitch)
    out = torch.full((nbytes,), 77, dtype=torch.int8, device="cuda")
    mx = torch.full((rows,), -1.0, dtype=torch.float64, device="cuda")
    scratch = torch.zeros(max(rows, 1), dtype=torch.int32, device="cuda")
    rc = L.ozk_split_int8(out.data_ptr(), pitch, mx.data_ptr(), scratch.data_ptr(), rows, length, x.data_ptr(), ld,
                          int(col_major), num_split, bits_, stream_ptr())
    assert rc == 0, f"ozk_split_int8 -> {rc}"
    torch.cuda.synchronize()
    return unblock(out.cpu().numpy(), num_split, rows, pitch), mx


class Reference:
    """The unmodified reference through oracle/_ref/libozref.so."""

    def __init__(self):
        self.L = oracle_lib.reference()
        assert self.L is not None
        h = C.c_void_p()
        assert self.L.ozref_create(C.byref(h)) == 0
        self.h = h

    def close(self):
        if self.h:
            self.L.ozref_destroy(self.h)
            self.h = None

    def gemm(self, op_a, op_b, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, mode):
        al, be = C.c_double(alpha), C.c_double(beta)
        self.L.ozref_set_stream(self.h, stream_ptr())
        rc = self.L.ozref_gemm(self.h, int(op_a), int(op_b), m, n, k, C.addressof(al), a.data_ptr(), lda, b.data_ptr(),
                               ldb, C.addressof(be), c.data_ptr(), ldc, int(mode), 0)
        assert rc == 0, f"ozref_gemm -> {rc}"

    def gemm_complex(self, op_a, op_b, m, n, k, alpha: complex, a, lda, b, ldb, beta: complex, c, ldc, mode):
        al = (C.c_double * 2)(alpha.real, alpha.imag)
        be = (C.c_double * 2)(beta.real, beta.imag)
        self.L.ozref_set_stream(self.h, stream_ptr())
        rc = self.L.ozref_gemm(self.h, int(op_a), int(op_b), m, n, k, C.addressof(al), a.data_ptr(), lda, b.data_ptr(),
                               ldb, C.addressof(be), c.data_ptr(), ldc, int(mode), 1)
        assert rc == 0, f"ozref_gemm (complex) -> {rc}"

    def dgemm_f32(self, op_a, op_b, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc):
        """the reference's compute mode `sgemm` (src/cublas_helper.cu:84-134, reached from its interposers only)"""
        self.L.ozref_set_stream(self.h, stream_ptr())
        self.L.ozref_dgemm_f32.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, C.c_double,
                                           C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_double, C.c_void_p,
                                           C.c_size_t]
        rc = self.L.ozref_dgemm_f32(self.h, int(op_a), int(op_b), m, n, k, float(alpha), a.data_ptr(), lda, b.data_ptr(),
                                    ldb, float(beta), c.data_ptr(), ldc)
        assert rc == 0, f"ozref_dgemm_f32 -> {rc}"

    def split(self, x: torch.Tensor, ld: int, m: int, n: int, op: int, matrix: int, num_split: int, bits_: int):
        """reference split_int8<double>: rows = m (matrix A) / n (matrix B after the swap)."""
        length = n if matrix == 0 else m
        rows = m if matrix == 0 else n
        ld4 = (length + 3) // 4 * 4
        out = torch.full((num_split, rows, ld4), 77, dtype=torch.int8, device="cuda")
        mx = torch.full((rows,), -1.0, dtype=torch.float64, device="cuda")
        rc = self.L.ozref_split_int8(out.data_ptr(), ld4, mx.data_ptr(), m, n, x.data_ptr(), ld, op, matrix, num_split,
                                     bits_, stream_ptr())
        assert rc == 0
        torch.cuda.synchronize()
        return out, mx

    def auto_mode_select(self, op_a, op_b, m, n, k, a, lda, b, ldb, threshold):
        cnt = (C.c_ulonglong * 8)()
        self.L.ozref_set_stream(self.h, stream_ptr())
        mode = self.L.ozref_auto_mode_select(self.h, int(op_a), int(op_b), m, n, k, a.data_ptr(), lda, b.data_ptr(),
                                             ldb, 0, float(threshold), C.addressof(cnt))
        return mode, [int(v) for v in cnt]
