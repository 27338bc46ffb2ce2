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

// ---- int8 slice layout ("blocked SW128") -------------------------------------------------------
// One operand plane is [slice][row tile][k block][128 rows][128 bytes]: every 128-row x 128-byte tile is
// 16 KB CONTIGUOUS in HBM and already carries the shared-memory 128-byte swizzle (16-byte chunk c of row r
// sits at chunk c ^ (r & 7)), so the GEMM kernel stages a tile with ONE linear bulk copy
// (cp.async.bulk, 54-76 B/clk/SM measured; ~115 clocks per copy whatever its size, profiles/r2_ubench_bulk_per_copy.txt)
// instead of a tiled TMA box of 128 separate rows (33-48 B/clk/SM; profiles/r1_ubench_sm_ingest.txt, r1_ubench_bulk_pair.txt).  Rows are padded to a multiple of 256 (one CTA
// pair's 2 x 128 rows), k to a multiple of 128; padding is zero.
constexpr size_t kTileRows = 128, kTileK = 128, kTileBytes = kTileRows * kTileK;
inline size_t slice_pitch(size_t k) { return (k + kTileK - 1) / kTileK * kTileK; }              // bytes of K per row
__host__ __device__ inline size_t slice_row_tiles(size_t rows) { return (rows + 255) / 256 * 2; }  // 128-row tiles
inline size_t slices_bytes(size_t rows, size_t k, unsigned num_split) {
  return static_cast<size_t>(num_split) * slice_row_tiles(rows) * kTileRows * slice_pitch(k);
}
// byte offset of the 16-byte chunk holding k positions [16*g, 16*g + 16) of row r inside one slice plane
__host__ __device__ inline size_t slice_chunk_offset(size_t r, size_t g, size_t k_blocks) {
  return ((r >> 7) * k_blocks + (g >> 3)) * kTileBytes + (r & 127) * kTileK + (((g & 7) ^ (r & 7)) << 4);
}

// Strided batch of operands split by one launch (grid.y for the row kernels, grid.z for the column kernels):
// entry e reads in + e*in_stride (doubles) and writes out + e*out_stride (bytes), max_exp + e*max_stride,
// scratch + e*scr_stride.  count = 1 with zero strides for a plain call.
struct SplitBatch {
  uint32_t count;
  size_t in_stride, out_stride, max_stride, scr_stride;
};

__host__ __device__ inline uint32_t ceil_div_u32(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

#define OZ_CUDA_TRY(expr)                                                                    \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) return static_cast<int>(_e);                                      \
  } while (0)

}  // namespace oz
