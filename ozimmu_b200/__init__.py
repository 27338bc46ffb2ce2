"""ozimmu_b200 -- B200-native Ozaki-scheme DGEMM (drop-in for enp1s0/ozIMMU's hot path).

The product is the shared library `lib/libozimmu.so` (hand-written sm_100a kernels, C++ host,
C-ABI in include/ozimmu_b200.h, cuBLAS interposers for LD_PRELOAD).  This package is the thin
Python mirror of the reference's public interface over that C-ABI.
"""
from ._lib import LIB_PATH, build, lib  # noqa: F401
from .api import *  # noqa: F401,F403
from .sharded import Comm, column_panels, comm_create, row_block, sharded_gemm, sharded_gemm_host  # noqa: F401
from .api import (auto_mode_select, compute_mode_t, create, destroy, element_kind_t, fp64_int8, gemm, gemm_host,  # noqa: F401
                  gemm_strided_batched, gemm_streamed_b,
                  get_bits_per_int8, get_compute_mode_name_str, handle_t, launch_count, malloc_mode_t, num_split_of,
                  operation_t, reallocate_working_memory, set_cuda_stream, set_scalar_pointer_mode)
