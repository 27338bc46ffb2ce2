// oz_common.cuh -- shared device/host helpers for the sm_100a kernels.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

namespace oz {

constexpr uint64_t kExpMask = 0x7FF0000000000000ull;   // binary64 exponent field
constexpr uint64_t kMantMask = 0x000FFFFFFFFFFFFFull;  // binary64 fraction field

// Global launch counter (bench.py reports it as gpu_launches).
extern unsigned long long g_launch_count;
inline void count_launch(unsigned n = 1) { g_launch_count += n; }

// k rounded up to 16: TMA needs 16-byte global strides.
inline size_t slice_pitch(size_t k) { return (k + 15) / 16 * 16; }

__host__ __device__ inline uint32_t ceil_div_u32(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

#define OZ_CUDA_TRY(expr)                                                                    \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) return static_cast<int>(_e);                                      \
  } while (0)

}  // namespace oz
