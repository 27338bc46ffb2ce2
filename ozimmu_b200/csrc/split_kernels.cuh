// split_kernels.cuh -- per-row max-exponent scan, FP64 -> s x int8 mantissa split, mantissa-loss totals.
//
// Replaces reference src/split.cu:14-67 (get_exp_max_element), :155-185 (cut_int8_core),
// :193-283 (split_int8_kernel + host wrappers) and :302-380 (mantissa-loss kernels).
// Arithmetic follows SURVEY App. A.2/A.3/A.6 bit for bit; the data movement is new:
//
//  * "rows contiguous" inputs (op_t A, op_n B): one CTA per row, the row is read from HBM
//    once into shared memory (bank-swizzled), the max exponent is a warp-shuffle reduction,
//    and every thread then emits 16 consecutive K elements as one 16-byte store per slice.
//  * "rows strided" inputs (op_n A, op_t B; column-major source): threads are mapped to
//    rows so that global loads stay coalesced along the contiguous dimension; the row max
//    is an atomicMax over the integer exponent field, and the split kernel has every thread
//    gather 16 K-consecutive elements of its row (each a coalesced column segment across the
//    warp) so it can also emit one 16-byte store per slice -- no shared-memory transpose.
//  * no device synchronisation (the reference calls cudaDeviceSynchronize per split,
//    src/split.cu:261).
//
// Output layout: the blocked, pre-swizzled [slice][row tile][k block][128][128] layout of oz_common.cuh
// (slice_chunk_offset); pitch = k rounded up to 128, rows padded to 256, padding zero.
#pragma once
#include <mutex>
#include "oz_common.cuh"
#include "ozimmu_b200.h"

namespace oz {
namespace {

constexpr int kSplitThreads = 256;
constexpr int kMaxCachedLen = 24 * 1024;  // doubles per row kept in SMEM (192 KB)

__device__ __forceinline__ uint32_t exp_field(double x) {
  return static_cast<uint32_t>((static_cast<uint64_t>(__double_as_longlong(x)) >> 52) & 0x7FFu);
}

// reference src/split.cu:191,202-204: max_exp = 2 * asdouble(max exponent field)
__device__ __forceinline__ double max_exp_from_field(uint32_t e) {
  return __dmul_rn(__longlong_as_double(static_cast<long long>(static_cast<uint64_t>(e) << 52)), 2.0);
}

// reference src/split.cu:155-185 for 16 K-consecutive elements of one row.
// w[t][q] receives bytes 4q..4q+3 of slice t.
template <int S>
__device__ __forceinline__ void cut16(const double (&v)[16], const uint64_t max_exp_bits,
                                      const unsigned L, uint32_t (&w)[S][4]) {
#pragma unroll
  for (int t = 0; t < S; t++) {
    w[t][0] = w[t][1] = w[t][2] = w[t][3] = 0u;
  }
#pragma unroll
  for (int j = 0; j < 16; j++) {
    const double a = v[j];
    const uint64_t bits = static_cast<uint64_t>(__double_as_longlong(a));
    const uint64_t ea = bits & kExpMask;
    const bool positive = a > 0;
    // 53-bit significand, MSB at bit 127 of a 128-bit word (hi:lo); no implicit bit for subnormals
    uint64_t hi = ((bits & kMantMask) | (ea ? (1ull << 52) : 0ull)) << 11;
    uint64_t lo = 0;
    const uint64_t off = (max_exp_bits - ea) >> 52;
    if (off >= 128) {
      hi = 0;
    } else if (off >= 64) {
      lo = hi >> (off - 64);
      hi = 0;
    } else if (off > 0) {
      lo = hi << (64 - off);
      hi >>= off;
    }
#pragma unroll
    for (int t = 0; t < S; t++) {
      const int32_t top = static_cast<int32_t>(hi >> (64 - L));
      const int32_t sv = positive ? top : -top;
      w[t][j >> 2] |= (static_cast<uint32_t>(sv) & 0xFFu) << (8 * (j & 3));
      hi = (hi << L) | (lo >> (64 - L));
      lo <<= L;
    }
  }
}

// The same cut for L == 7 (every k <= 2^17, i.e. every BASELINE configuration) with the 128-bit
// aligned significand held in four 32-bit words: one variable shift (word select + funnel shifts),
// then every slice is a COMPILE-TIME bit-field (2 instructions), signed by one IMAD and dropped into
// its byte lane by one PRMT -- about half the integer instructions of the generic version, which is what
// bounds the split kernels (they are issue-bound before they are HBM-bound).
template <int S>
__device__ __forceinline__ void cut16_l7(const double (&v)[16], const uint64_t max_exp_bits, uint32_t (&w)[S][4]) {
  const uint32_t emaxp1 = static_cast<uint32_t>(max_exp_bits >> 52);  // exponent field of 2*max (sign is 0)
#pragma unroll
  for (int t = 0; t < S; t++) {
    w[t][0] = w[t][1] = w[t][2] = w[t][3] = 0u;
  }
#pragma unroll
  for (int j = 0; j < 16; j++) {
    const uint32_t hi32 = static_cast<uint32_t>(__double2hiint(v[j]));
    const uint32_t lo32 = static_cast<uint32_t>(__double2loint(v[j]));
    const uint32_t ea = (hi32 >> 20) & 0x7FFu;
    const int sgn = (v[j] > 0) ? 1 : -1;
    const uint32_t mh = (hi32 & 0xFFFFFu) | (ea ? 0x100000u : 0u);
    // significand MSB at bit 127: words (x3, x2, 0, 0), then >> off, off = (emax+1) - e(a) (mod 2^12 as in
    // the reference's 64-bit subtraction followed by >> 52)
    const uint32_t x3 = (mh << 11) | (lo32 >> 21), x2 = lo32 << 11;
    const uint32_t off = (emaxp1 - ea) & 0xFFFu;
    const uint32_t bs = off & 31u, ws = off >> 5;
    const uint32_t u3 = x3 >> bs, u2 = __funnelshift_r(x2, x3, bs), u1 = __funnelshift_r(0u, x2, bs);
    uint32_t W[4];  // W[3] = bits 127..96
    W[3] = (ws == 0) ? u3 : 0u;
    W[2] = (ws == 0) ? u2 : (ws == 1) ? u3 : 0u;
    W[1] = (ws == 0) ? u1 : (ws == 1) ? u2 : (ws == 2) ? u3 : 0u;
    W[0] = (ws == 1) ? u1 : (ws == 2) ? u2 : (ws == 3) ? u3 : 0u;
#pragma unroll
    for (int t = 0; t < S; t++) {
      constexpr int kDummy = 0;
      (void)kDummy;
      const int p = 121 - 7 * t;          // bit position of the field's LSB (compile time after unrolling)
      const int word = p >> 5, sh = p & 31;
      uint32_t f;
      if (sh <= 25) f = W[word] >> sh;
      else f = __funnelshift_r(W[word], W[word + 1], sh);
      const int sv = static_cast<int>(f & 0x7Fu) * sgn;
      w[t][j >> 2] = __byte_perm(w[t][j >> 2], static_cast<uint32_t>(sv), 0x3210u ^ ((0x4u ^ (j & 3)) << (4 * (j & 3))));
    }
  }
}

template <int S>
__device__ __forceinline__ void cut16_any(const double (&v)[16], const uint64_t max_exp_bits, const unsigned L,
                                          uint32_t (&w)[S][4]) {
  if (L == 7) cut16_l7<S>(v, max_exp_bits, w);
  else cut16<S>(v, max_exp_bits, L, w);
}

__device__ __forceinline__ uint32_t block_max_u32(uint32_t v, uint32_t *s_red) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  v = (threadIdx.x < kSplitThreads / 32) ? s_red[threadIdx.x] : 0u;
  if (threadIdx.x < 32) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (threadIdx.x == 0) s_red[0] = v;
  }
  __syncthreads();
  return s_red[0];
}

// SMEM position of row element i: 16-element groups, in-group index XOR-swizzled by the
// group id so that "thread t reads element j of group t" is bank-conflict free.
__device__ __forceinline__ uint32_t swz(uint32_t i) { return (i & ~15u) | ((i ^ (i >> 4)) & 15u); }

// ---------------------------------------------------------------------------------------------
// Rows contiguous in memory: one CTA per row.
// ---------------------------------------------------------------------------------------------
template <int S, bool CACHED>
__global__ void __launch_bounds__(kSplitThreads)
split_rows_kernel(int8_t *__restrict__ out_, const size_t pitch, double *__restrict__ max_exp_,
                  const size_t rows, const uint32_t len, const double *__restrict__ in_,
                  const size_t ld, const unsigned L, const uint32_t es, const size_t slice_stride,
                  const SplitBatch bt) {
  int8_t *__restrict__ out = out_ + blockIdx.y * bt.out_stride;
  double *__restrict__ max_exp = max_exp_ + blockIdx.y * bt.max_stride;
  const double *__restrict__ in = in_ + blockIdx.y * bt.in_stride;
  // slice_stride: bytes between consecutive slices of the destination plane (the plane may hold more rows than
  // this launch cuts: row-block calls of the host-operand pipeline).
  // es: distance between consecutive elements in doubles (1 = real matrix, 2 = one plane of an
  // interleaved complex matrix; `in` then points at the plane's first double, ld counts complex elements)
  extern __shared__ double s_row[];
  __shared__ uint32_t s_red[kSplitThreads / 32];
  const size_t row = blockIdx.x;
  const double *__restrict__ src = in + row * ld * es;

  uint32_t e = 0;
  for (uint32_t i = threadIdx.x; i < len; i += kSplitThreads) {
    const double x = __ldg(src + static_cast<size_t>(i) * es);
    if (CACHED) s_row[swz(i)] = x;
    e = max(e, exp_field(x));
  }
  e = block_max_u32(e, s_red);  // also orders the s_row writes before the reads below
  const double mx = max_exp_from_field(e);
  const uint64_t mx_bits = static_cast<uint64_t>(__double_as_longlong(mx));
  if (threadIdx.x == 0) max_exp[row] = mx;

  const size_t kblocks = pitch / kTileK;
  const uint32_t ngroups = static_cast<uint32_t>(pitch / 16);
  for (uint32_t g = threadIdx.x; g < ngroups; g += kSplitThreads) {
    double v[16];
#pragma unroll
    for (int j = 0; j < 16; j++) {
      const uint32_t i = g * 16 + j;
      if (CACHED) {
        v[j] = (i < len) ? s_row[(g * 16) | ((j ^ g) & 15u)] : 0.0;
      } else {
        v[j] = (i < len) ? __ldg(src + static_cast<size_t>(i) * es) : 0.0;
      }
    }
    uint32_t w[S][4];
    cut16_any<S>(v, mx_bits, L, w);
#pragma unroll
    for (int t = 0; t < S; t++) {
      uint4 q = make_uint4(w[t][0], w[t][1], w[t][2], w[t][3]);
      // elements >= len are exact zeros => their slice bytes are already 0
      *reinterpret_cast<uint4 *>(out + t * slice_stride + slice_chunk_offset(row, g, kblocks)) = q;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Rows contiguous, row length <= 32 * THREADS: the row lives in registers.  Every thread loads the
// (up to two) 16-element groups it will cut as 16-byte vector loads (one HBM pass, no SMEM staging, so
// occupancy is bounded by registers only and all of a row's loads are in flight at once), the block
// reduces the exponent maximum, and the groups are cut straight from registers.
// ---------------------------------------------------------------------------------------------
template <int S, int THREADS, int GROUPS>
__global__ void __launch_bounds__(THREADS)
split_rows_reg_kernel(int8_t *__restrict__ out_, const size_t pitch, double *__restrict__ max_exp_,
                      const size_t rows, const uint32_t len, const double *__restrict__ in_,
                      const size_t ld, const unsigned L, const size_t slice_stride, const SplitBatch bt) {
  int8_t *__restrict__ out = out_ + blockIdx.y * bt.out_stride;
  double *__restrict__ max_exp = max_exp_ + blockIdx.y * bt.max_stride;
  const double *__restrict__ in = in_ + blockIdx.y * bt.in_stride;
  __shared__ uint32_t s_red[THREADS / 32];
  __shared__ uint32_t s_max;
  const size_t row = blockIdx.x;
  const double *__restrict__ src = in + row * ld;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(src) & 15u) == 0);
  double v[GROUPS][16];
  uint32_t e = 0;
#pragma unroll
  for (int gi = 0; gi < GROUPS; gi++) {
    const uint32_t g = threadIdx.x + gi * THREADS;
    const uint32_t i0 = g * 16;
    if (i0 + 16 <= len && vec_ok) {
#pragma unroll
      for (int q = 0; q < 8; q++) {
        const double2 d = __ldg(reinterpret_cast<const double2 *>(src + i0) + q);
        v[gi][2 * q] = d.x;
        v[gi][2 * q + 1] = d.y;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; j++) v[gi][j] = (i0 + j < len) ? __ldg(src + i0 + j) : 0.0;
    }
#pragma unroll
    for (int j = 0; j < 16; j++) e = max(e, exp_field(v[gi][j]));
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) e = max(e, __shfl_xor_sync(0xffffffffu, e, o));
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = e;
  __syncthreads();
  if (threadIdx.x < 32) {
    uint32_t x = (threadIdx.x < THREADS / 32) ? s_red[threadIdx.x] : 0u;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) x = max(x, __shfl_xor_sync(0xffffffffu, x, o));
    if (threadIdx.x == 0) s_max = x;
  }
  __syncthreads();
  const double mx = max_exp_from_field(s_max);
  const uint64_t mx_bits = static_cast<uint64_t>(__double_as_longlong(mx));
  if (threadIdx.x == 0) max_exp[row] = mx;

  const size_t kblocks = pitch / kTileK;
  const uint32_t ngroups = static_cast<uint32_t>(pitch / 16);
#pragma unroll
  for (int gi = 0; gi < GROUPS; gi++) {
    const uint32_t g = threadIdx.x + gi * THREADS;
    if (g < ngroups) {
      uint32_t w[S][4];
      cut16_any<S>(v[gi], mx_bits, L, w);
      int8_t *__restrict__ dst = out + slice_chunk_offset(row, g, kblocks);
#pragma unroll
      for (int t = 0; t < S; t++)
        *reinterpret_cast<uint4 *>(dst + t * slice_stride) = make_uint4(w[t][0], w[t][1], w[t][2], w[t][3]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Rows strided (column-major source): thread <-> row so loads coalesce along rows.
// ---------------------------------------------------------------------------------------------
constexpr int kColChunk = 64;  // columns per CTA in the row-max pass

__global__ void __launch_bounds__(256)
rowmax_cols_kernel(uint32_t *__restrict__ emax_, const size_t rows, const uint32_t len,
                   const double *__restrict__ in_, const size_t ld, const uint32_t es, const SplitBatch bt) {
  uint32_t *__restrict__ emax = emax_ + blockIdx.z * bt.scr_stride;
  const double *__restrict__ in = in_ + blockIdx.z * bt.in_stride;
  const size_t r = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x;
  if (r >= rows) return;
  const uint32_t c0 = blockIdx.y * kColChunk;
  const uint32_t c1 = min(len, c0 + kColChunk);
  uint32_t e = 0;
#pragma unroll 8
  for (uint32_t c = c0; c < c1; c++) e = max(e, exp_field(__ldg(in + (static_cast<size_t>(c) * ld + r) * es)));
  atomicMax(emax + r, e);
}

// CTA = 32 rows x 128 K-positions: warp w cuts K-group w (16 positions) of 32 consecutive rows, so every
// gathered load is one coalesced 256-byte row segment of a column.  The 16-byte slice pieces are
// transposed through shared memory (XOR-swizzled, conflict-free) and leave as full 128-byte lines per
// row and slice -- writing them straight from the cutting threads would scatter 16-byte partial-sector
// stores over 32 rows per instruction.
constexpr int kColsRows = 32, kColsK = 128;

// ONE launch does both passes, band by band: the matrix is cut into bands of kBandRows rows (16 MB of FP64 at
// k = 8192), and the CTAs of a band are, in blockIdx order, first its row-max CTAs (256 rows x 64 columns each, as
// rowmax_cols_kernel) and then its cut CTAs.  A cut CTA waits until the band's row-max CTAs have all arrived (a counter
// in global memory; they have lower block indices, so they were dispatched earlier and the wait cannot deadlock -- the
// same dependency direction as a decoupled look-back scan), then re-reads its 32 x 128 elements -- from L2, where the
// band still sits: the matrix crosses HBM once instead of twice (two separate launches over 512 MiB: the second pass
// finds nothing of the first in the 126 MB L2).
constexpr uint32_t kBandRows = 256;

struct ColsBands {
  uint32_t bands, p1_per_band, p2_per_band;   // CTAs per band: row-max pass, cut pass
  uint32_t rchunks, rtiles;                   // 256-row chunks (pass 1) and 32-row tiles (pass 2) per full band
  uint32_t col_chunk;                         // columns per row-max CTA (kColChunk; fewer for small matrices)
  uint32_t *band_done;                        // [batch][bands], zeroed before the launch
};

template <int S>
__global__ void __launch_bounds__(256)
split_cols_kernel(int8_t *__restrict__ out_, const size_t pitch, double *__restrict__ max_exp_,
                  uint32_t *__restrict__ emax_, const size_t rows, const uint32_t len,
                  const double *__restrict__ in_, const size_t ld, const unsigned L, const uint32_t es,
                  const size_t slice_stride, const SplitBatch bt, const ColsBands cb) {
  int8_t *__restrict__ out = out_ + blockIdx.z * bt.out_stride;
  double *__restrict__ max_exp = max_exp_ + blockIdx.z * bt.max_stride;
  uint32_t *__restrict__ emax = emax_ + blockIdx.z * bt.scr_stride;
  const double *__restrict__ in = in_ + blockIdx.z * bt.in_stride;
  uint32_t *done = cb.band_done + blockIdx.z * cb.bands;
  extern __shared__ uint4 s_out[];  // [S][32 rows][8 chunks of 16 B], chunk index ^ (row & 7)
  // block order: max(0), max(1), cut(0), max(2), cut(1), ..., max(B-1), cut(B-2), cut(B-1) -- the row-max pass of band
  // b + 1 is issued before the cut pass of band b, so it runs while that one still occupies the SMs (a cut CTA only ever
  // waits for CTAs with lower block indices)
  const uint32_t per_band = cb.p1_per_band + cb.p2_per_band;
  uint32_t band, j;   // j < p1: row-max CTA j of the band, else cut CTA j - p1
  if (blockIdx.x < cb.p1_per_band) {
    band = 0, j = blockIdx.x;
  } else {
    const uint32_t idx = blockIdx.x - cb.p1_per_band, g = idx / per_band, r = idx - g * per_band;
    if (g + 1 < cb.bands) {
      if (r < cb.p1_per_band) band = g + 1, j = r;
      else band = g, j = r;
    } else {
      band = g, j = cb.p1_per_band + r;
    }
  }
  const size_t row_lo = static_cast<size_t>(band) * kBandRows;
  if (j < cb.p1_per_band) {
    // ---- pass 1: exponent maximum of 256 rows over 64 columns (thread <-> row: coalesced column segments) ----
    const size_t r = row_lo + static_cast<size_t>(j % cb.rchunks) * 256 + threadIdx.x;
    const uint32_t c0 = (j / cb.rchunks) * cb.col_chunk;
    if (r < rows && r < row_lo + kBandRows && c0 < len) {
      const uint32_t c1 = min(len, c0 + cb.col_chunk);
      uint32_t e = 0;
#pragma unroll 8
      for (uint32_t c = c0; c < c1; c++) e = max(e, exp_field(__ldg(in + (static_cast<size_t>(c) * ld + r) * es)));
      atomicMax(emax + r, e);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(done + band, 1u);
    return;
  }
  // ---- pass 2: cut 32 rows x 128 K-positions once the band's row maxima are final ----
  if (threadIdx.x == 0) {
    while (*reinterpret_cast<volatile uint32_t *>(done + band) < cb.p1_per_band) __nanosleep(64);
    __threadfence();
  }
  __syncthreads();
  const uint32_t jj = j - cb.p1_per_band;
  const uint32_t rl = threadIdx.x & 31, cg = threadIdx.x >> 5;
  const size_t tile_row = row_lo + static_cast<size_t>(jj % cb.rtiles) * kColsRows;
  const size_t r = tile_row + rl;
  const uint32_t kbase = (jj / cb.rtiles) * kColsK;
  const uint32_t cbase = kbase + cg * 16;
  if (r < rows && cbase < pitch) {
    const double mx = max_exp_from_field(__ldcg(emax + r));
    const uint64_t mx_bits = static_cast<uint64_t>(__double_as_longlong(mx));
    if (kbase == 0 && cg == 0) max_exp[r] = mx;
    double v[16];
#pragma unroll
    for (int jx = 0; jx < 16; jx++) {
      const uint32_t c = cbase + jx;
      v[jx] = (c < len) ? __ldcs(in + (static_cast<size_t>(c) * ld + r) * es) : 0.0;   // last use: evict first
    }
    uint32_t w[S][4];
    cut16_any<S>(v, mx_bits, L, w);
#pragma unroll
    for (int t = 0; t < S; t++)
      s_out[(t * kColsRows + rl) * 8 + (cg ^ (rl & 7))] = make_uint4(w[t][0], w[t][1], w[t][2], w[t][3]);
  }
  __syncthreads();
  // write-out: 8 consecutive threads cover one row's 128-byte line of one slice
  const uint32_t orow = threadIdx.x >> 3, chunk = threadIdx.x & 7;
  const size_t gr = tile_row + orow;
  const uint32_t gk = kbase + chunk * 16;
  if (gr < rows && gk < pitch) {
    int8_t *__restrict__ dst = out + slice_chunk_offset(gr, gk >> 4, pitch / kTileK);
#pragma unroll
    for (int t = 0; t < S; t++)
      __stcs(reinterpret_cast<uint4 *>(dst + t * slice_stride), s_out[(t * kColsRows + orow) * 8 + (chunk ^ (orow & 7))]);
  }
}

// ---------------------------------------------------------------------------------------------
// Mantissa-loss totals (auto mode).  reference src/split.cu:317-380, intended semantics
// (SURVEY App. A.6): zeros contribute nothing, 16 counters for num_split = 3..18.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void loss_accumulate(uint32_t (&cnt)[16], const double x,
                                                const uint64_t mx_exp_bits, const bool mx_zero,
                                                const unsigned L) {
  if (x == 0 || mx_zero) return;
  const uint32_t req = static_cast<uint32_t>(
      ((mx_exp_bits - (static_cast<uint64_t>(__double_as_longlong(x)) & kExpMask)) >> 52) + 53);
#pragma unroll
  for (int s = 0; s < 16; s++) {
    const uint32_t space = (s + 3) * L;
    cnt[s] += (space < req) ? (req - space) : 0u;
  }
}

__device__ __forceinline__ void loss_block_reduce(uint32_t (&cnt)[16], unsigned long long *out) {
  __shared__ unsigned long long s_cnt[16];
  if (threadIdx.x < 16) s_cnt[threadIdx.x] = 0ull;
  __syncthreads();
#pragma unroll
  for (int s = 0; s < 16; s++) {
    uint32_t v = cnt[s];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_cnt[s], static_cast<unsigned long long>(v));
  }
  __syncthreads();
  if (threadIdx.x < 16 && s_cnt[threadIdx.x]) atomicAdd(out + threadIdx.x, s_cnt[threadIdx.x]);
}

__global__ void __launch_bounds__(kSplitThreads)
loss_rows_kernel(unsigned long long *__restrict__ counters_, const uint32_t len,
                 const double *__restrict__ in, const size_t ld, const unsigned L, const uint32_t es,
                 const size_t in_stride) {
  // blockIdx.y = entry of a strided batch: its own input and its own 16 counters
  __shared__ uint32_t s_red[kSplitThreads / 32];
  unsigned long long *__restrict__ counters = counters_ + 16 * blockIdx.y;
  const double *__restrict__ src = in + blockIdx.y * in_stride + static_cast<size_t>(blockIdx.x) * ld * es;
  uint32_t e = 0;
  for (uint32_t i = threadIdx.x; i < len; i += kSplitThreads)
    e = max(e, exp_field(__ldg(src + static_cast<size_t>(i) * es)));
  e = block_max_u32(e, s_red);
  const double mx = max_exp_from_field(e);
  const uint64_t mx_exp_bits = static_cast<uint64_t>(__double_as_longlong(mx)) & kExpMask;
  uint32_t cnt[16];
#pragma unroll
  for (int s = 0; s < 16; s++) cnt[s] = 0;
  for (uint32_t i = threadIdx.x; i < len; i += kSplitThreads)
    loss_accumulate(cnt, __ldg(src + static_cast<size_t>(i) * es), mx_exp_bits, mx == 0, L);
  loss_block_reduce(cnt, counters);
}

__global__ void __launch_bounds__(256)
loss_cols_kernel(unsigned long long *__restrict__ counters_, const uint32_t *__restrict__ emax_,
                 const size_t rows, const uint32_t len, const double *__restrict__ in_,
                 const size_t ld, const unsigned L, const uint32_t es, const SplitBatch bt) {
  // blockIdx.z = entry of a strided batch
  unsigned long long *__restrict__ counters = counters_ + 16 * blockIdx.z;
  const uint32_t *__restrict__ emax = emax_ + blockIdx.z * bt.scr_stride;
  const double *__restrict__ in = in_ + blockIdx.z * bt.in_stride;
  const size_t r = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x;
  uint32_t cnt[16];
#pragma unroll
  for (int s = 0; s < 16; s++) cnt[s] = 0;
  if (r < rows) {
    const double mx = max_exp_from_field(emax[r]);
    const uint64_t mx_exp_bits = static_cast<uint64_t>(__double_as_longlong(mx)) & kExpMask;
    const uint32_t c0 = blockIdx.y * kColChunk;
    const uint32_t c1 = min(len, c0 + kColChunk);
#pragma unroll 4
    for (uint32_t c = c0; c < c1; c++)
      loss_accumulate(cnt, __ldg(in + (static_cast<size_t>(c) * ld + r) * es), mx_exp_bits, mx == 0, L);
  }
  loss_block_reduce(cnt, counters);
}

// Band counters of the strided split: a ring of small zeroed device buffers per device (a launch zeroes its buffer in
// stream order; kBandRing launches may be in flight per device before a buffer is reused, and reuse is ordered by the
// memset only on the same stream -- concurrent splits on more streams than that would share counters, so the ring is
// sized well above the handful of streams the host pipelines use).
inline uint32_t *band_counters(size_t count, cudaStream_t stream) {
  constexpr int kMaxDev = 64, kBandRing = 16;
  constexpr size_t kBandCap = 1 << 16;   // counters per buffer (bands x batch entries of one launch)
  static std::mutex mu;
  static uint32_t *pool[kMaxDev] = {};   // kBandRing buffers, ONE allocation per device on first use (a later call
  static int next[kMaxDev] = {};         // may be inside a CUDA graph capture, where cudaMalloc is not allowed)
  int dev = 0;
  if (count > kBandCap || cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) return nullptr;
  uint32_t *buf = nullptr;
  {
    std::lock_guard<std::mutex> lock(mu);
    if (pool[dev] == nullptr && cudaMalloc(&pool[dev], kBandRing * kBandCap * sizeof(uint32_t)) != cudaSuccess) {
      pool[dev] = nullptr;
      cudaGetLastError();
      return nullptr;
    }
    buf = pool[dev] + static_cast<size_t>(next[dev]) * kBandCap;
    next[dev] = (next[dev] + 1) % kBandRing;
  }
  if (cudaMemsetAsync(buf, 0, count * sizeof(uint32_t), stream) != cudaSuccess) return nullptr;
  return buf;
}

template <int S>
int launch_split(int8_t *out, size_t pitch, size_t plane_rows, double *max_exp, uint32_t *scratch, size_t rows,
                 size_t len, const double *in, size_t ld, int col_major, unsigned L, uint32_t es,
                 cudaStream_t stream, const SplitBatch bt = SplitBatch{1, 0, 0, 0, 0}) {
  if (bt.count == 0 || bt.count > 65535) return static_cast<int>(cudaErrorInvalidValue);
  // `out` points at this call's first row tile inside a plane of plane_rows rows (plane_rows == rows for a
  // whole-matrix call)
  const size_t slice_stride = slice_row_tiles(plane_rows) * kTileRows * pitch;
  if (col_major) {
    if (bt.count == 1)
      OZ_CUDA_TRY(cudaMemsetAsync(scratch, 0, rows * sizeof(uint32_t), stream));
    else
      OZ_CUDA_TRY(cudaMemset2DAsync(scratch, bt.scr_stride * sizeof(uint32_t), 0, rows * sizeof(uint32_t), bt.count, stream));
    // [rows] exponent maxima, followed by the band counters (the caller's scratch holds rows entries; the counters
    // live in a small per-device pool)
    ColsBands cb{};
    cb.bands = static_cast<uint32_t>((rows + kBandRows - 1) / kBandRows);
    cb.rchunks = kBandRows / 256;
    cb.rtiles = kBandRows / kColsRows;
    // a small matrix (the whole grid fits on the GPU at once) is latency-bound: four times as many row-max CTAs,
    // each with a quarter of the dependent load batches
    cb.col_chunk = (rows * len <= (2048u * 2048u)) ? kColChunk / 4 : kColChunk;
    cb.p1_per_band = cb.rchunks * ceil_div_u32(static_cast<uint32_t>(len), cb.col_chunk);
    cb.p2_per_band = cb.rtiles * static_cast<uint32_t>((pitch + kColsK - 1) / kColsK);
    cb.band_done = band_counters(static_cast<size_t>(cb.bands) * bt.count, stream);
    if (cb.band_done == nullptr) return static_cast<int>(cudaErrorMemoryAllocation);
    dim3 g2(cb.bands * (cb.p1_per_band + cb.p2_per_band), 1, bt.count);
    const size_t smem_cols = static_cast<size_t>(S) * kColsRows * kColsK;
    if (smem_cols > 48 * 1024)  // per device and cheap: no caching
      OZ_CUDA_TRY(cudaFuncSetAttribute(split_cols_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem_cols)));
    split_cols_kernel<S><<<g2, 256, smem_cols, stream>>>(out, pitch, max_exp, scratch, rows,
                                                         static_cast<uint32_t>(len), in, ld, L, es, slice_stride, bt, cb);
    count_launch(1);
  } else if (es == 1 && len <= 16384) {
    // register-resident rows: (threads, 16-element groups per thread) sized to the row
    const dim3 nrows(static_cast<unsigned>(rows), bt.count);
    const uint32_t len32 = static_cast<uint32_t>(len);
    if (len <= 2048)
      split_rows_reg_kernel<S, 128, 1><<<nrows, 128, 0, stream>>>(out, pitch, max_exp, rows, len32, in, ld, L, slice_stride, bt);
    else if (len <= 4096)
      split_rows_reg_kernel<S, 256, 1><<<nrows, 256, 0, stream>>>(out, pitch, max_exp, rows, len32, in, ld, L, slice_stride, bt);
    else if (len <= 8192)
      split_rows_reg_kernel<S, 256, 2><<<nrows, 256, 0, stream>>>(out, pitch, max_exp, rows, len32, in, ld, L, slice_stride, bt);
    else
      split_rows_reg_kernel<S, 512, 2><<<nrows, 512, 0, stream>>>(out, pitch, max_exp, rows, len32, in, ld, L, slice_stride, bt);
    count_launch(1);
  } else {
    if (len <= static_cast<size_t>(kMaxCachedLen)) {
      const size_t smem = ((len + 15) / 16 * 16) * sizeof(double);
      if (smem > 48 * 1024) {
        OZ_CUDA_TRY(cudaFuncSetAttribute(split_rows_kernel<S, true>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem)));
      }
      split_rows_kernel<S, true><<<dim3(static_cast<unsigned>(rows), bt.count), kSplitThreads, smem, stream>>>(
          out, pitch, max_exp, rows, static_cast<uint32_t>(len), in, ld, L, es, slice_stride, bt);
    } else {
      split_rows_kernel<S, false><<<dim3(static_cast<unsigned>(rows), bt.count), kSplitThreads, 0, stream>>>(
          out, pitch, max_exp, rows, static_cast<uint32_t>(len), in, ld, L, es, slice_stride, bt);
    }
    count_launch(1);
  }
  return static_cast<int>(cudaGetLastError());
}

}  // namespace
}  // namespace oz

