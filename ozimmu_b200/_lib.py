"""ctypes binding of ozimmu_b200/lib/libozimmu.so (the C-ABI declared in include/ozimmu_b200.h).

The library is the product: sm_100a kernels + C++ host + cuBLAS interposers.  There is no
Python or CPU fallback -- if the shared object is missing, importing a symbol raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "lib" / "libozimmu.so"

_lib = None

c_size_t, c_int, c_uint, c_void_p, c_double = C.c_size_t, C.c_int, C.c_uint, C.c_void_p, C.c_double

# name -> (restype, argtypes); mirrors include/ozimmu_b200.h one to one
PROTOTYPES = {
    "ozk_bits_per_int8": (C.c_uint32, [C.c_uint32]),
    "ozk_slice_pitch": (c_size_t, [c_size_t]),
    "ozk_slices_bytes": (c_size_t, [c_size_t, c_size_t, c_uint]),
    "ozk_split_int8": (c_int, [c_void_p, c_size_t, c_void_p, c_void_p, c_size_t, c_size_t, c_void_p, c_size_t,
                               c_int, c_uint, c_uint, c_void_p]),
    "ozk_split_int8_strided": (c_int, [c_void_p, c_size_t, c_void_p, c_void_p, c_size_t, c_size_t, c_void_p, c_size_t,
                                       c_int, c_uint, c_uint, c_uint, c_void_p]),
    "ozk_split_int8_block": (c_int, [c_void_p, c_size_t, c_size_t, c_size_t, c_void_p, c_void_p, c_size_t, c_size_t,
                                     c_void_p, c_size_t, c_int, c_uint, c_uint, c_uint, c_void_p]),
    "ozk_gemm_i8_fused_block": (c_int, [c_size_t, c_size_t, c_size_t, c_void_p, c_size_t, c_size_t, c_void_p, c_size_t,
                                        c_size_t, c_size_t, c_void_p, c_void_p, c_uint, c_uint, c_double, c_double,
                                        c_void_p, c_size_t, c_uint, c_void_p]),
    "ozk_gemm_i8_fused_batched": (c_int, [c_size_t, c_size_t, c_size_t, c_size_t, c_void_p, c_size_t, c_void_p, c_size_t,
                                          c_size_t, c_void_p, c_size_t, c_void_p, c_size_t, c_uint, c_uint, c_double,
                                          c_double, c_void_p, c_size_t, c_size_t, c_void_p]),
    "ozimmu_gemm_strided_batched": (c_int, [c_void_p, c_int, c_int, c_size_t, c_size_t, c_size_t, c_void_p, c_void_p,
                                            c_size_t, C.c_longlong, c_void_p, c_size_t, C.c_longlong, c_void_p, c_void_p,
                                            c_size_t, C.c_longlong, c_size_t, c_int]),
    "ozimmu_gemm_streamed_b": (c_int, [c_void_p, c_int, c_int, c_size_t, c_size_t, c_size_t, c_void_p, c_void_p, c_size_t,
                                       c_void_p, c_size_t, c_void_p, c_void_p, c_size_t, c_int, c_size_t, c_void_p,
                                       c_void_p]),
    "ozk_split_int8_batched": (c_int, [c_void_p, c_size_t, c_size_t, c_void_p, c_size_t, c_void_p, c_size_t, c_size_t,
                                       c_size_t, c_void_p, c_size_t, c_size_t, c_int, c_uint, c_uint, c_size_t,
                                       c_void_p]),
    "ozk_mantissa_loss_strided": (c_int, [c_void_p, c_void_p, c_size_t, c_size_t, c_void_p, c_size_t, c_int, c_uint,
                                          c_uint, c_void_p]),
    "ozk_gemm_i8_fused_ex": (c_int, [c_void_p, c_void_p]),
    "ozk_zgemm_combine": (c_int, [c_size_t, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                  c_void_p]),
    "ozk_scale_c_ex": (c_int, [c_size_t, c_size_t, c_void_p, c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
    "ozk_split_int8_batched_strided": (c_int, [c_void_p, c_size_t, c_size_t, c_void_p, c_size_t, c_void_p, c_size_t,
                                               c_size_t, c_size_t, c_void_p, c_size_t, c_size_t, c_int, c_uint, c_uint,
                                               c_uint, c_size_t, c_void_p]),
    "ozk_mantissa_loss_batched": (c_int, [c_void_p, c_void_p, c_size_t, c_size_t, c_size_t, c_void_p, c_size_t, c_size_t,
                                          c_int, c_uint, c_uint, c_size_t, c_void_p]),
    "ozimmu_gemm_strided_batched_ex": (c_int, [c_void_p, c_int, c_int, c_size_t, c_size_t, c_size_t, c_void_p, c_void_p,
                                               c_size_t, C.c_longlong, c_void_p, c_size_t, C.c_longlong, c_void_p,
                                               c_void_p, c_size_t, C.c_longlong, c_size_t, c_int, c_int]),
    "ozimmu_set_scalar_pointer_mode": (c_int, [c_void_p, c_int]),
    "ozk_gemm_i8_fused": (c_int, [c_size_t, c_size_t, c_size_t, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p,
                                  c_uint, c_uint, c_double, c_double, c_void_p, c_size_t, c_void_p]),
    "ozk_scale_c": (c_int, [c_size_t, c_size_t, c_double, c_void_p, c_size_t, c_void_p]),
    "ozk_set_cluster_shape": (c_int, [c_int, c_int]),
    "ozk_fused_tile_choice": (c_int, [c_size_t, c_size_t, c_size_t, c_size_t, c_int, c_int, c_void_p, c_void_p]),
    "ozk_gemm_i8_pair": (c_int, [c_size_t, c_size_t, c_size_t, c_void_p, c_void_p, c_size_t, c_uint, c_uint, c_uint,
                                 c_void_p, c_void_p]),
    "ozk_mantissa_loss": (c_int, [c_void_p, c_void_p, c_size_t, c_size_t, c_void_p, c_size_t, c_int, c_uint,
                                  c_void_p]),
    "ozimmu_create": (c_int, [C.POINTER(c_void_p), c_int]),
    "ozimmu_destroy": (c_int, [c_void_p]),
    "ozimmu_set_cuda_stream": (c_int, [c_void_p, c_void_p]),
    "ozimmu_enable_profiling": (c_int, [c_void_p]),
    "ozimmu_disable_profiling": (c_int, [c_void_p]),
    "ozimmu_print_profiler_result": (c_int, [c_void_p, C.c_char_p, c_int]),
    "ozimmu_clear_profiler_result": (c_int, [c_void_p]),
    "ozimmu_set_auto_mantissa_loss_threshold": (c_int, [c_void_p, c_double]),
    "ozimmu_get_auto_mantissa_loss_threshold": (c_double, [c_void_p]),
    "ozimmu_reallocate_working_memory": (c_size_t, [c_void_p, c_int, c_int, c_size_t, c_size_t, c_size_t, c_int,
                                                    c_int]),
    "ozimmu_reallocate_working_memory_bytes": (c_size_t, [c_void_p, c_size_t]),
    "ozimmu_gemm": (c_int, [c_void_p, c_int, c_int, c_size_t, c_size_t, c_size_t, c_void_p, c_void_p, c_size_t,
                            c_void_p, c_size_t, c_void_p, c_void_p, c_size_t, c_int, c_int]),
    "ozimmu_auto_mode_select": (c_int, [c_void_p, c_int, c_int, c_size_t, c_size_t, c_size_t, c_void_p, c_size_t,
                                        c_void_p, c_size_t, c_int, c_double, c_void_p]),
    "ozimmu_get_compute_mode_name_str": (C.c_char_p, [c_int]),
    "ozimmu_get_bits_per_int8": (C.c_uint32, [C.c_uint32]),
    "ozimmu_gemm_host": (c_int, [c_void_p, c_int, c_int, c_size_t, c_size_t, c_size_t, c_void_p, c_void_p, c_size_t,
                                 c_void_p, c_size_t, c_void_p, c_void_p, c_size_t, c_int]),
    "ozimmu_host_block_edges": (c_size_t, [c_size_t, c_size_t, c_int, c_void_p, c_size_t]),
    "ozimmu_comm_unique_id": (c_int, [c_void_p]),
    "ozimmu_comm_create": (c_int, [C.POINTER(c_void_p), c_int, c_int, c_void_p]),
    "ozimmu_comm_adopt": (c_int, [C.POINTER(c_void_p), c_void_p]),
    "ozimmu_comm_destroy": (c_int, [c_void_p]),
    "ozimmu_comm_rank": (c_int, [c_void_p]),
    "ozimmu_comm_size": (c_int, [c_void_p]),
    "ozimmu_row_block": (None, [c_size_t, c_int, c_int, C.POINTER(c_size_t), C.POINTER(c_size_t)]),
    "ozimmu_gemm_sharded": (c_int, [c_void_p, c_void_p, c_int, c_int, c_size_t, c_size_t, c_size_t, c_void_p, c_void_p,
                                    c_size_t, c_void_p, c_size_t, c_void_p, c_void_p, c_size_t, c_int, c_int, c_uint]),
    "ozimmu_gemm_sharded_host": (c_int, [c_void_p, c_void_p, c_int, c_int, c_size_t, c_size_t, c_size_t, c_void_p,
                                         c_void_p, c_size_t, c_void_p, c_size_t, c_void_p, c_void_p, c_size_t, c_int,
                                         c_int]),
    "ozimmu_sharded_panel_edges": (c_size_t, [c_size_t, c_size_t, c_void_p, c_size_t]),
    "ozimmu_launch_count": (C.c_ulonglong, []),
}

# the cuBLAS entry points the library interposes under LD_PRELOAD (reference src/cublas.cu:103-513)
INTERPOSED = [
    "cublasCreate_v2", "cublasDestroy_v2", "cublasGemmEx", "cublasDgemm_v2", "cublasZgemm_v2",
    "cublasGemmStridedBatchedEx", "cublasDgemmStridedBatched", "cublasZgemmStridedBatched",
]


def build(verbose: bool = False) -> Path:
    """Compile lib/libozimmu.so in-tree (nvcc, sm_100a).  Needs no GPU."""
    out = subprocess.run(["make", "-C", str(PKG_DIR), "-j", str(os.cpu_count() or 4)], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout[-4000:])
        print(out.stderr[-4000:])
    if out.returncode != 0:
        raise RuntimeError("building libozimmu.so failed")
    return LIB_PATH


def lib() -> C.CDLL:
    """The loaded shared library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `make -C {PKG_DIR}` (or __graft_entry__.build()); "
                "ozimmu_b200 has no CPU or PyTorch fallback")
        handle = C.CDLL(str(LIB_PATH), mode=os.RTLD_LOCAL | os.RTLD_NOW)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib
