"""Host-side mirror of the reference's public interface (reference include/ozimmu/ozimmu.hpp:12-100)
over the C-ABI of libozimmu.so.  Same names, argument order and meaning, same return codes:
`gemm` returns 0 on success and 1 for an invalid argument; CUDA failures raise RuntimeError
(the reference throws std::runtime_error).

Matrices are BLAS column-major.  `a_ptr`/`b_ptr`/`c_ptr` may be raw device addresses (int) or
torch CUDA tensors of dtype float64 (their storage is used as-is: a row-major torch tensor of
shape (k, m) IS a column-major m x k matrix with lda = its row stride).
"""
from __future__ import annotations

import ctypes as C
import enum
from typing import Optional, Sequence, Tuple, Union

from . import _lib


class operation_t(enum.IntEnum):  # reference include/ozimmu/ozimmu.hpp:12
    op_n = 0
    op_t = 1


class compute_mode_t(enum.IntEnum):  # reference include/ozimmu/ozimmu.hpp:14-37
    sgemm = 0
    dgemm = 1
    fp64_int8_3 = 2
    fp64_int8_4 = 3
    fp64_int8_5 = 4
    fp64_int8_6 = 5
    fp64_int8_7 = 6
    fp64_int8_8 = 7
    fp64_int8_9 = 8
    fp64_int8_10 = 9
    fp64_int8_11 = 10
    fp64_int8_12 = 11
    fp64_int8_13 = 12
    fp64_int8_14 = 13
    fp64_int8_15 = 14
    fp64_int8_16 = 15
    fp64_int8_17 = 16
    fp64_int8_18 = 17
    fp64_int8_auto = 18


class malloc_mode_t(enum.IntEnum):  # :41
    malloc_sync = 0
    malloc_async = 1


class element_kind_t(enum.IntEnum):  # :43-46
    real = 0
    complx = 1


op_n, op_t = operation_t.op_n, operation_t.op_t
real, complx = element_kind_t.real, element_kind_t.complx


def fp64_int8(num_split: int) -> compute_mode_t:
    """compute mode of a split count 3..18"""
    if not 3 <= num_split <= 18:
        raise ValueError("split count must be in 3..18")
    return compute_mode_t(num_split - 1)


def num_split_of(mode: compute_mode_t) -> int:
    mode = compute_mode_t(mode)
    if not compute_mode_t.fp64_int8_3 <= mode <= compute_mode_t.fp64_int8_18:
        raise ValueError(f"{mode.name} has no split count")
    return int(mode) + 1


def _ptr(x) -> int:
    if x is None:
        return 0
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return int(x.data_ptr())
    if hasattr(x, "ctypes"):
        return int(x.ctypes.data)
    raise TypeError(f"cannot take the address of {type(x)!r}")


def _stream_ptr(stream) -> int:
    if stream is None:
        return 0
    if isinstance(stream, int):
        return stream
    return int(stream.cuda_stream)  # torch.cuda.Stream


def _check(rc: int, what: str) -> int:
    if rc < 0:
        raise RuntimeError(f"{what} failed inside libozimmu.so (see the [ozIMMU ERROR] line above)")
    return rc


class handle_t:
    """Opaque handle (reference include/ozimmu/ozimmu.hpp:9-11, src/handle.hpp:6-31)."""

    def __init__(self, raw: int):
        self.raw = raw

    def __bool__(self):
        return bool(self.raw)


def create(malloc_mode: malloc_mode_t = malloc_mode_t.malloc_sync) -> handle_t:  # :47
    raw = C.c_void_p()
    _check(_lib.lib().ozimmu_create(C.byref(raw), int(malloc_mode)), "create")
    return handle_t(raw.value)


def destroy(handle: handle_t) -> int:  # :48
    rc = _check(_lib.lib().ozimmu_destroy(handle.raw), "destroy")
    handle.raw = 0
    return rc


def set_cuda_stream(handle: handle_t, cuda_stream) -> None:  # :49-50
    _lib.lib().ozimmu_set_cuda_stream(handle.raw, _stream_ptr(cuda_stream))


def enable_profiling(handle: handle_t) -> None:  # :52
    _lib.lib().ozimmu_enable_profiling(handle.raw)


def disable_profiling(handle: handle_t) -> None:  # :53
    _lib.lib().ozimmu_disable_profiling(handle.raw)


def print_profiler_result(handle: handle_t, tag: str, csv: bool = False) -> None:  # :54-55
    _lib.lib().ozimmu_print_profiler_result(handle.raw, tag.encode(), int(csv))


def clear_profiler_result(handle: handle_t) -> None:  # :56
    _lib.lib().ozimmu_clear_profiler_result(handle.raw)


def set_auto_mantissa_loss_threashold(handle: handle_t, threshold: float) -> None:  # :58-59 (sic)
    _lib.lib().ozimmu_set_auto_mantissa_loss_threshold(handle.raw, float(threshold))


def get_auto_mantissa_loss_threashold(handle: handle_t) -> float:  # :60
    return float(_lib.lib().ozimmu_get_auto_mantissa_loss_threshold(handle.raw))


gemm_params_t = Tuple[operation_t, operation_t, int, int, int, element_kind_t, compute_mode_t]


def reallocate_working_memory(handle: handle_t, arg: Union[int, Sequence[gemm_params_t]]) -> int:  # :68-74
    """Grow-only workspace; returns the new size in bytes if it grew, else 0."""
    L = _lib.lib()
    if isinstance(arg, int):
        return int(L.ozimmu_reallocate_working_memory_bytes(handle.raw, arg))
    grown = 0
    for (oa, ob, m, n, k, kind, mode) in arg:
        grown = max(grown, int(L.ozimmu_reallocate_working_memory(handle.raw, int(oa), int(ob), m, n, k, int(kind),
                                                                  int(mode))))
    return grown


def gemm(handle: handle_t, op_A: operation_t, op_B: operation_t, m: int, n: int, k: int, alpha: float, a_ptr, lda: int,
         b_ptr, ldb: int, beta: float, c_ptr, ldc: int, compute_mode: compute_mode_t,
         element_kind: element_kind_t = element_kind_t.real) -> int:  # :76-83
    """C = alpha * op(A) * op(B) + beta * C on the handle's stream (asynchronous)."""
    al, be, pal, pbe = _scalars(alpha, beta, element_kind)
    rc = _lib.lib().ozimmu_gemm(handle.raw, int(op_A), int(op_B), m, n, k, pal, _ptr(a_ptr), lda,
                                _ptr(b_ptr), ldb, pbe, _ptr(c_ptr), ldc, int(compute_mode),
                                int(element_kind))
    return _check(rc, "gemm")


def _scalars(alpha, beta, element_kind):
    """alpha / beta as the C-ABI takes them: host double (real) / double[2] (complex), or -- after
    set_scalar_pointer_mode(handle, True) -- device tensors / addresses holding them"""
    if hasattr(alpha, "data_ptr") or hasattr(beta, "data_ptr"):
        return alpha, beta, _ptr(alpha), _ptr(beta)
    if element_kind == element_kind_t.real:
        al, be = C.c_double(alpha), C.c_double(beta)
    else:
        alpha, beta = complex(alpha), complex(beta)
        al, be = (C.c_double * 2)(alpha.real, alpha.imag), (C.c_double * 2)(beta.real, beta.imag)
    return al, be, C.addressof(al), C.addressof(be)


def set_scalar_pointer_mode(handle: handle_t, on_device: bool) -> None:
    """alpha / beta of the following gemm / gemm_strided_batched calls are device pointers (cuBLAS device pointer mode)"""
    _lib.lib().ozimmu_set_scalar_pointer_mode(handle.raw, int(bool(on_device)))


def gemm_strided_batched(handle: handle_t, op_A: operation_t, op_B: operation_t, m: int, n: int, k: int, alpha,
                         a_ptr, lda: int, stride_a: int, b_ptr, ldb: int, stride_b: int, beta, c_ptr, ldc: int,
                         stride_c: int, batch_count: int, compute_mode: compute_mode_t,
                         element_kind: element_kind_t = element_kind_t.real) -> int:
    """Strided batch of DGEMMs / ZGEMMs (X_e = X + e*stride_x elements) through one grouped launch; what the
    reference's cublas{D,Z}gemmStridedBatched / cublasGemmStridedBatchedEx interposers loop over
    (reference src/cublas.cu:315-512).  Asynchronous on the handle's stream."""
    al, be, pal, pbe = _scalars(alpha, beta, element_kind)
    rc = _lib.lib().ozimmu_gemm_strided_batched_ex(handle.raw, int(op_A), int(op_B), m, n, k, pal, _ptr(a_ptr),
                                                   lda, stride_a, _ptr(b_ptr), ldb, stride_b, pbe,
                                                   _ptr(c_ptr), ldc, stride_c, batch_count, int(compute_mode),
                                                   int(element_kind))
    return _check(rc, "gemm_strided_batched")


def gemm_streamed_b(handle: handle_t, op_A: operation_t, op_B: operation_t, m: int, n: int, k: int, alpha: float, a_ptr,
                    lda: int, b_ptr, ldb: int, beta: float, c_ptr, ldc: int, compute_mode: compute_mode_t,
                    col_edges: Sequence[int], ready_events: Sequence[int]) -> int:
    """gemm() when B becomes valid column panel by column panel: panel p = columns [col_edges[p], col_edges[p+1])
    may be read once the CUDA event ready_events[p] (raw cudaEvent_t handles, e.g. torch.cuda.Event.cuda_event) has
    fired.  Inner edges must be multiples of 256.  Asynchronous on the handle's stream."""
    al, be = C.c_double(alpha), C.c_double(beta)
    npan = len(ready_events)
    assert len(col_edges) == npan + 1
    edges = (C.c_size_t * (npan + 1))(*[int(x) for x in col_edges])
    evs = (C.c_void_p * npan)(*[int(e) for e in ready_events])
    rc = _lib.lib().ozimmu_gemm_streamed_b(handle.raw, int(op_A), int(op_B), m, n, k, C.addressof(al), _ptr(a_ptr), lda,
                                           _ptr(b_ptr), ldb, C.addressof(be), _ptr(c_ptr), ldc, int(compute_mode), npan,
                                           C.addressof(edges), C.addressof(evs))
    return _check(rc, "gemm_streamed_b")


def gemm_host(handle: handle_t, op_A: operation_t, op_B: operation_t, m: int, n: int, k: int, alpha: float, a_host,
              lda: int, b_host, ldb: int, beta: float, c_host, ldc: int, compute_mode: compute_mode_t) -> int:
    """Same product with HOST operands (numpy arrays / pinned CPU tensors); returns when C is complete."""
    al, be = C.c_double(alpha), C.c_double(beta)
    rc = _lib.lib().ozimmu_gemm_host(handle.raw, int(op_A), int(op_B), m, n, k, C.addressof(al), _ptr(a_host), lda,
                                     _ptr(b_host), ldb, C.addressof(be), _ptr(c_host), ldc, int(compute_mode))
    return _check(rc, "gemm_host")


def auto_mode_select(handle: handle_t, op_A: operation_t, op_B: operation_t, m: int, n: int, k: int, a_ptr, lda: int,
                     b_ptr, ldb: int, element_kind: element_kind_t, mantissa_loss_threshold: float,
                     counters_out: Optional[list] = None) -> compute_mode_t:  # :85-94
    cnt = (C.c_ulonglong * 16)()
    rc = _lib.lib().ozimmu_auto_mode_select(handle.raw, int(op_A), int(op_B), m, n, k, _ptr(a_ptr), lda, _ptr(b_ptr),
                                            ldb, int(element_kind), float(mantissa_loss_threshold), C.addressof(cnt))
    _check(rc, "auto_mode_select")
    if counters_out is not None:
        counters_out[:] = [int(v) for v in cnt]
    return compute_mode_t(rc)


def get_compute_mode_name_str(mode: compute_mode_t) -> str:  # :96
    s = _lib.lib().ozimmu_get_compute_mode_name_str(int(mode))
    if s is None:
        raise RuntimeError(f"unknown compute mode {mode}")
    return s.decode()


def get_bits_per_int8(k: int) -> int:  # :102
    return int(_lib.lib().ozimmu_get_bits_per_int8(k))


def launch_count() -> int:
    """Kernels launched by libozimmu.so since load."""
    return int(_lib.lib().ozimmu_launch_count())
