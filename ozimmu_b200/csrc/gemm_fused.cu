// gemm_fused.cu -- the hot path: s(s+1)/2 exact int8 x int8 -> int32 slice products on tcgen05 tensor cores,
// each folded into a per-element FP64 accumulator in the reference's exact op order, then the exponent rescale
// and alpha/beta -- ONE persistent kernel.
//
// Replaces, fused: reference src/gemm.cu:266-334 (matmul_core -> cublasGemmEx int8, 45 library GEMMs at s=9),
// :77-102 (accumulate_in_f64, one 20 B/element HBM pass per product), :104-122 (init_accumulator_buffer) and
// :124-158 (axby); for complex data also :160-239 (axy_complex, init_c_complex).  Arithmetic per output element
// follows SURVEY App. A.4/A.5 exactly (same FMA sequence => bit-identical C).
//
// Design (B200 / sm_100a), see DESIGN.md 3.2 for the measurements behind each choice:
//  * CTA pair = one 256 x BN tile of C (tcgen05.mma.cta_group::2.kind::i8, UMMA M=256, N=BN, K=32), BN = 256 for large
//    problems, 240 ... 128 where the tile count quantises better or k is small (dispatch_fused: a table of measured
//    per-round times per tile shape and k).  Each CTA owns 128 x BN outputs for the whole pair list, so the FP64
//    accumulators never leave the SM: 96 columns per epilogue thread in registers, the rest (BN=256: 32) in SMEM.
//    Small problems: 128 x 128 tiles, 64 rows per CTA (UMMA M=128, template parameter BMC_ = 64): twice as many tiles,
//    8 KB of A per k-step; the pair MMA then folds a CTA's 64 x 128 block of D over all 128 TMEM lanes (PairCfg).
//  * Operands: int8 slices in the blocked, pre-swizzled layout of oz_common.cuh -- every 128-row x 128-byte tile
//    is 16 KB contiguous, so a stage is filled by two linear bulk copies (cp.async.bulk: 54-76 B/clk/SM) rather
//    than tiled TMA boxes of 128 separate rows (33-48 B/clk/SM).  What bounds the kernel is the SS-mode MMA itself
//    (~45 clocks per instruction on top of its tensor time: 80 % of the pipe) and, sustained, power.
//  * Roles per CTA: warp 0 producer (bulk copies, completing on the CTA's own `full` barrier), warp 1 MMA issuer
//    (leader CTA only), warp 2 TMEM allocator, warp 3 relay (non-leader CTA: forwards "my stage landed" to the
//    leader, because a bulk copy can only signal a barrier in the destination CTA), warps 4-11 epilogue.
//    Role loops run on the uniform datapath (whole warp loops, one elected lane issues).
//  * int32 products sit in TMEM buffers (BN=256: 2 x 256 columns, BN=128: 4 x 128), so the MMAs of product p+1
//    overlap the epilogue of product p.  tcgen05.commit multicasts "stage free" / "product ready" to both CTAs.
//  * Epilogue per product: tcgen05.ld 16 columns (128-wide tiles keep the next load in flight), (double)p by the
//    magic-number trick (exact), then acc = fma(p, 2^(32-rshift), acc).  After the last pair: x = acc*2^-44*amax[r]*bmax[c];
//    C = alpha*x (+ beta*C).  The "buffer drained" hand-off to the MMA warp is a relaxed arrive (no release fence).
//  * Soft lockstep between CTA pairs keeps the pairs of one wave within a few k-steps of each other so that they
//    share slice panels through L2.
//  * Tile queue: entry x tiles_m x tiles_n (entry = index in a strided batch, ozk_gemm_i8_fused_batched), static
//    round-robin over ceil(tiles / rounds) CTA pairs so that every pair owns the same number of tiles; block launches
//    (ozk_gemm_i8_fused_block) address a rectangle of C inside the operands' slice planes and may run one CTA pair per
//    tile (OZK_FUSED_ONE_TILE_PER_PAIR), which makes the hardware CTA scheduler the tile queue of the block pipelines.
//    (A device-side tile queue gated by ready flags was built and measured in rounds 1-2 and removed: 28 ms against
//    26 ms for the multi-launch pipeline at 8192^3 end to end, profiles/r2_e2e_sweep.txt.)
#include <cmath>
#include <cstdlib>
#include <mutex>

#include <cuda.h>
#include <cuda_runtime.h>

#include "oz_common.cuh"
#include "ozimmu_b200.h"
#include "ptx.cuh"

namespace oz {
namespace {

constexpr uint32_t BK = 128;  // bytes (== int8 elements) of K per stage; rows per CTA: template parameter BMC_ (128 or 64)
constexpr uint32_t kThreads = 384;
constexpr uint32_t kEpiWarps = 8;
constexpr uint32_t kUmmaK = 32;
constexpr uint32_t kSyncEvery = 16;     // k-steps per soft-lockstep interval

struct FusedParams {
  uint32_t m, n;
  uint32_t k_blocks;           // pitch / 128
  uint32_t num_split;
  int32_t bits;                // L = bits per int8 slice
  uint32_t single_a, single_b; // != 0: raw mode, only this (1-based) pair, int32 output
  uint32_t tiles_m, tiles_n;   // grid of pair tiles: 256 rows x BN columns
  uint32_t group_m;            // rasterisation band height in tiles
  uint32_t rt_a, rt_b;         // 128-row tiles per slice plane of A / B (rows padded to 256)
  const int8_t *a_slices;
  const int8_t *b_slices;
  double alpha, beta;
  double *c;
  unsigned long long ldc;
  const double *amax;
  const double *bmax;
  int32_t *c_i32;
  // complex GEMM (reference src/gemm.cu:412-521): C is cuDoubleComplex (ldc in complex elements), each operand is two
  // planes of slices (real, imaginary) `*_plane_bytes` apart with row scales `*max_plane` apart, and every tile runs
  // the FOUR real plane products back to back -- (im,im) -> -alpha, (re,re) -> +alpha, (im,re), (re,im) -> i*alpha,
  // the reference's order (:479-518) -- each followed by y = fma(x, coef, y) on the tile of C (axy_complex, :160-186);
  // the first one starts from y = beta*y (init_c_complex, :188-239).  (alpha, alpha_im), (beta, beta_im) = the scalars.
  uint32_t cplx;
  uint32_t groups;             // plane products per tile: 1 (real) or 4 (complex)
  double alpha_im, beta_im;
  unsigned long long a_plane_bytes, b_plane_bytes, amax_plane, bmax_plane;
  // cuBLAS device pointer mode: alpha / beta (1 double each, 2 for complex) are read from device memory in the
  // finalize step, in stream order, instead of being passed by value
  const double *alpha_dev, *beta_dev;
  // soft lockstep between CTA pairs
  uint32_t *sync_ctr;          // one arrival counter per kSyncEvery k-steps, zeroed before launch (or null)
  uint32_t sync_window;        // a pair starts sync interval j only after interval j - window is complete
  uint32_t sync_len;           // counters available
  uint32_t no_lockstep;        // host only: never pace this launch (it shares the GPU with other launches)
  uint32_t one_tile_per_pair;  // host only: grid = one CTA pair per tile (non-persistent; the hardware CTA scheduler is
                               // the tile queue, and higher-priority kernels get SMs whenever a tile ends)
  uint32_t b_rows;             // rows of B's slice planes that exist from b_slices on (multiple of 128): a tile
                               // column narrower than 256 may reach past the plane's end
  // strided batch (grouped launch): tile index = entry * tiles_m * tiles_n + tile inside the entry; every entry
  // has its own slices / row scales / C at these distances (bytes for the slices, doubles for the rest)
  uint32_t batch;
  unsigned long long a_batch_bytes, b_batch_bytes, amax_batch, bmax_batch, c_batch;
};

// reference src/config.cu:85-92: for sum = 2..s+1, for j = 1..sum-1: (A_id=j, B_id=sum-j)
struct PairIter {
  uint32_t s, sum, a;
  bool single, done;
  __device__ PairIter(const FusedParams &p) {
    s = p.num_split;
    single = p.single_a != 0;
    done = false;
    if (single) {
      sum = p.single_a + p.single_b;
      a = p.single_a;
    } else {
      sum = 2;
      a = 1;
    }
  }
  __device__ bool valid() const { return !done; }
  __device__ uint32_t a_id() const { return a; }
  __device__ uint32_t b_id() const { return sum - a; }
  __device__ void next() {
    if (single) {
      done = true;
      return;
    }
    a++;
    if (a >= sum || a > s) {
      sum++;
      a = (sum - 1 > s) ? sum - s : 1;  // keep B_id = sum - a <= s (never hit for sum <= s+1)
    }
    if (sum > s + 1) done = true;
  }
  // 2^(32 - rshift), rshift = L*(A_id+B_id-2) - 2*(7-L)   (reference src/gemm.cu:394-401, :96-99)
  __device__ double scale(int32_t L) const {
    const int32_t e = 32 - (L * static_cast<int32_t>(sum - 2) - 2 * (7 - L));
    return __longlong_as_double(static_cast<long long>(static_cast<uint64_t>(1023 + e) << 52));
  }
};

// Complex GEMM: plane (0 = real, 1 = imaginary) of A and of B that plane product g multiplies -- reference
// src/gemm.cu:479-480: (im,im), (re,re), (im,re), (re,im); always plane 0 for a real GEMM.
__device__ __forceinline__ uint32_t group_plane_a(const FusedParams &p, uint32_t g) { return p.cplx ? (0x5u >> g) & 1u : 0u; }
__device__ __forceinline__ uint32_t group_plane_b(const FusedParams &p, uint32_t g) { return p.cplx ? (0x9u >> g) & 1u : 0u; }

// tile index -> (tile row, tile column): bands of group_m tile rows, column-major inside a band, so the tiles
// that run concurrently cover a compact block of C and share their A / B panels through L2
__device__ __forceinline__ void tile_coords(const FusedParams &p, uint32_t t, uint32_t &tm, uint32_t &tn,
                                            uint32_t &entry) {
  const uint32_t per_entry = p.tiles_m * p.tiles_n;
  entry = t / per_entry;
  t -= entry * per_entry;
  const uint32_t group_size = p.group_m * p.tiles_n;
  const uint32_t g = t / group_size;
  const uint32_t first = g * p.group_m;
  const uint32_t rows = min(p.tiles_m - first, p.group_m);
  const uint32_t r = t - g * group_size;
  tm = first + r % rows;
  tn = r / rows;
}

// BMC_ = rows of C per CTA: 128 (UMMA M = 256, the default) or 64 (UMMA M = 128: twice as many, half as tall tiles for
// problems that would leave SMs idle, each CTA moving 8 KB of A per k-step instead of 16).  With M = 128 the pair MMA
// folds a CTA's 64 x N block of D into all 128 TMEM lanes: lanes 0-63 hold columns [0, N/2), lanes 64-127 hold columns
// [N/2, N) of the same 64 rows (measured: tools/ubench/umma_m128_probe.cu, profiles/r2_umma_m128_probe.txt), so a product
// occupies N/2 TMEM columns.
template <uint32_t BN_, uint32_t BMC_ = 128>
struct PairCfg {
  static_assert(BN_ % 16 == 0 && BN_ >= 128 && BN_ <= 256, "UMMA M=256 needs N % 16 == 0, N <= 256");
  static_assert(BMC_ == 128 || (BMC_ == 64 && BN_ == 128), "rows per CTA: 128, or 64 with a 128-wide tile");
  static constexpr uint32_t kABytes = BMC_ * BK;
  static constexpr uint32_t kBRows = BN_ / 2;                        // this CTA's half of the B tile
  static constexpr uint32_t kBBytes = kBRows * BK;
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;         // per CTA (a multiple of 1 KB: SW128 atoms)
  static constexpr uint32_t kAccBufs = (BN_ == 128) ? 4 : 2;
  static constexpr uint32_t kBufStride = (BMC_ == 128) ? BN_ : BN_ / 2;   // TMEM columns between buffers
  // epilogue: 4 lane quarters x 2 warps; BMC_ = 128: the two warps of a quarter take the column halves of the tile;
  // BMC_ = 64: quarters 2, 3 already are the upper column half, the two warps halve the N/2 TMEM columns
  static constexpr uint32_t kColsPerThread = (BMC_ == 128) ? BN_ / 2 : BN_ / 4;
  // FP64 accumulators: up to 96 columns per thread in registers (192 registers); a wider tile keeps
  // the rest in shared memory ([column][row] doubles, conflict-free for lane <-> row)
  static constexpr uint32_t kRegCols = kColsPerThread < 96 ? kColsPerThread : 96;
  static constexpr uint32_t kSpillCols = kColsPerThread - kRegCols;
  static constexpr uint32_t kSpillBytes = 2 * kSpillCols * 128 * 8;
  // the operand ring takes what is left of the 227 KB: BN=256 -> 5 stages, 240 -> 5, 224 -> 6, 208 -> 7, 192 / 128 -> 8;
  // 16 KB stages (BMC_ = 64): 12
  static constexpr uint32_t kSmemMax = 227 * 1024, kBarReserve = 512, kAlign = 1024;
  static constexpr uint32_t kStagesFit = (kSmemMax - kSpillBytes - kBarReserve - kAlign) / kStageBytes;
  static constexpr uint32_t kStagesCap = (BMC_ == 128) ? 8 : 12;
  static constexpr uint32_t kStages = kStagesFit > kStagesCap ? kStagesCap : kStagesFit;
  static constexpr uint32_t kBarBytes = 8 * (3 * kStages + 2 * kAccBufs) + 16;
  static_assert(kBarBytes <= kBarReserve, "barrier block");
  static constexpr uint32_t kSmemBytes = kStages * kStageBytes + kSpillBytes + kBarBytes + kAlign;
  // registers to spare for a second TMEM load buffer in the epilogue (acc 128 + 2 x 16 + 16 of 224)
  static constexpr bool kLdPipe = (BN_ == 128);
  static constexpr uint32_t kRegsOther = (BN_ == 128) ? 56 : 40;
  static constexpr uint32_t kRegsEpi = (BN_ == 128) ? 224 : 232;
};

template <uint32_t BN_, uint32_t BMC_>
__global__ void __launch_bounds__(kThreads, 1)
oz_gemm_pair_kernel(const FusedParams p) {
  using Cfg = PairCfg<BN_, BMC_>;
  constexpr uint32_t BM = BMC_;
  constexpr uint32_t kStagesP = Cfg::kStages, kBufs = Cfg::kAccBufs, kCols = Cfg::kColsPerThread;
  constexpr uint32_t kRegCols = Cfg::kRegCols, kSpillCols = Cfg::kSpillCols;
  constexpr bool kLdPipe = Cfg::kLdPipe;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t spill_base = smem_base + kStagesP * Cfg::kStageBytes;
  const uint32_t bar_base = spill_base + Cfg::kSpillBytes;
  auto full_bar = [&](uint32_t s) { return bar_base + 8u * s; };                       // this CTA's stage landed
  auto pfull_bar = [&](uint32_t s) { return bar_base + 8u * (kStagesP + s); };          // leader: peer's stage landed
  auto empty_bar = [&](uint32_t s) { return bar_base + 8u * (2 * kStagesP + s); };
  auto tfull_bar = [&](uint32_t b) { return bar_base + 8u * (3 * kStagesP + b); };
  auto tempty_bar = [&](uint32_t b) { return bar_base + 8u * (3 * kStagesP + kBufs + b); };
  const uint32_t tmem_slot = bar_base + 8u * (3 * kStagesP + 2 * kBufs);
  volatile uint32_t *tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));

  // warp-uniform by construction (shfl from lane 0), so that the single-warp role loops below stay on the
  // uniform datapath: tcgen05.mma / bulk-copy operands must be uniform registers, and values produced under
  // lane-divergent control flow force a per-instruction ELECT/R2UR waterfall.
  const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t rank = __shfl_sync(0xffffffffu, ptx::cluster_ctarank(), 0) & 1u;  // 0 = leader (issues the MMAs)
  const uint32_t pair_id = blockIdx.x >> 1;
  const uint32_t num_pairs = gridDim.x >> 1;
  const uint32_t num_tiles = p.tiles_m * p.tiles_n * p.batch;
  const uint32_t steps_per_tile = (p.single_a != 0 ? 1u : p.num_split * (p.num_split + 1) / 2) * p.k_blocks * p.groups;

  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < kStagesP; s++) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(pfull_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    for (uint32_t b = 0; b < kBufs; b++) {
      ptx::mbar_init(tfull_bar(b), 1);
      ptx::mbar_init(tempty_bar(b), 2 * kEpiWarps);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) ptx::tmem_alloc_2sm<512>(tmem_slot);
  ptx::tc_fence_before();
  ptx::cluster_arrive();
  ptx::cluster_wait();
  ptx::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot_ptr, 0);

  if (warp < 4) {
    ptx::reg_dealloc<Cfg::kRegsOther>();
    if (warp == 0) {
      // ===================== producer (both CTAs; whole warp loops, one lane issues) ================
      const bool issuer = ptx::elect_one();
      uint32_t stage = 0, ph = 0;
      // Soft lockstep: the CTA pairs of one wave share slice panels through L2 only while they stay within a
      // few k-steps of each other; left alone they drift apart by whole products and every panel is re-fetched
      // from HBM.  Leaders count arrivals per kSyncEvery-step interval and do not start interval j before all
      // pairs have started interval j - window.  A performance hint only: the wait is bounded and abandoned for
      // the rest of the kernel on the first timeout.
      const uint32_t tiles_max = (num_tiles + num_pairs - 1) / num_pairs;
      const uint32_t pairs_last = num_tiles - (tiles_max - 1) * num_pairs;  // pairs that own tiles_max tiles
      bool lockstep = p.sync_ctr != nullptr && rank == 0;
      uint32_t g = 0;
      for (uint32_t t = pair_id; t < num_tiles; t += num_pairs) {   // static round-robin tile list of this pair
        uint32_t tm, tn, entry = 0;
        tile_coords(p, t, tm, tn, entry);
        const int8_t *a_base = p.a_slices + static_cast<size_t>(entry) * p.a_batch_bytes;
        const int8_t *b_base = p.b_slices + static_cast<size_t>(entry) * p.b_batch_bytes;
        // this CTA's 128 rows of A (one 16 KB tile per k-block) and its BN/2 rows of B: rows [b_row0, b_row0 + BN/2) of the
        // plane, which lie in one 128-row tile (BN = 256, 128) or straddle two (other widths): up to two linear pieces,
        // each a whole number of 8-row swizzle atoms; rows past the end of the plane are not fetched (their columns of
        // C do not exist)
        // BM = 128: this CTA's rows are 128-row slice tile 2 tm + rank; BM = 64: half `rank` of slice tile tm
        const size_t a_tile = BM == 128 ? static_cast<size_t>(tm) * 2 + rank : static_cast<size_t>(tm);
        const size_t a_sub = BM == 128 ? 0 : static_cast<size_t>(rank) * BM * BK;
        const uint32_t b_row0 = tn * BN_ + rank * Cfg::kBRows;
        const uint32_t b_avail = b_row0 < p.b_rows ? p.b_rows - b_row0 : 0u;
        const uint32_t b_rows1 = min(min(Cfg::kBRows, 128u - (b_row0 & 127u)), b_avail);
        // (BN = 256 / 128: the half-tile never straddles two slice tiles -- one piece, known at compile time)
        constexpr bool kTwoPieces = (BN_ != 256 && BN_ != 128);
        const uint32_t b_rows2 = kTwoPieces ? min(Cfg::kBRows - b_rows1, b_avail - b_rows1) : 0u;
        const size_t b_tile = b_row0 >> 7;
        const size_t b_sub = static_cast<size_t>(b_row0 & 127u) * BK;
        const uint32_t stage_tx = BM * BK + (b_rows1 + b_rows2) * BK;
        for (uint32_t grp = 0; grp < p.groups; grp++)
        for (PairIter it(p); it.valid(); it.next()) {
          // running source pointers: one 16 KB slice tile further per k-step
          const int8_t *a_ptr = a_base + group_plane_a(p, grp) * p.a_plane_bytes +
                                ((it.a_id() - 1) * static_cast<size_t>(p.rt_a) + a_tile) * p.k_blocks * kTileBytes + a_sub;
          const int8_t *b_ptr = b_base + group_plane_b(p, grp) * p.b_plane_bytes +
                                ((it.b_id() - 1) * static_cast<size_t>(p.rt_b) + b_tile) * p.k_blocks * kTileBytes + b_sub;
          const size_t b2_off = static_cast<size_t>(p.k_blocks) * kTileBytes - b_sub;   // same k-step, next row tile
          auto fill_stage = [&]() {
            ptx::mbar_wait(empty_bar(stage), ph ^ 1u);
            const uint32_t dst = smem_base + stage * Cfg::kStageBytes;
            if (issuer) {
              ptx::mbar_expect_tx(full_bar(stage), stage_tx);
              ptx::bulk_load(dst, a_ptr, BM * BK, full_bar(stage));
              if (b_rows1) ptx::bulk_load(dst + BM * BK, b_ptr, b_rows1 * BK, full_bar(stage));
              if (kTwoPieces && b_rows2)
                ptx::bulk_load(dst + BM * BK + b_rows1 * BK, b_ptr + b2_off, b_rows2 * BK, full_bar(stage));
            }
            a_ptr += kTileBytes;
            b_ptr += kTileBytes;
            if (++stage == kStagesP) { stage = 0; ph ^= 1u; }
          };
          if (!lockstep) {
            // unpaced (single-round launches, block pipelines, batches): nothing but the ring in the loop
            for (uint32_t kb = 0; kb < p.k_blocks; kb++) fill_stage();
            g += p.k_blocks;
            continue;
          }
          for (uint32_t kb = 0; kb < p.k_blocks; kb++, g++) {
            if (lockstep && (g % kSyncEvery) == 0) {
              const uint32_t j = g / kSyncEvery;
              if (j >= p.sync_len) {
                lockstep = false;
              } else {
                uint32_t ok = 1;
                if (issuer) {
                  atomicAdd(p.sync_ctr + j, 1u);
                  if (j >= p.sync_window) {
                    const uint32_t jw = j - p.sync_window;
                    const uint32_t want =
                        (static_cast<uint64_t>(jw) * kSyncEvery < static_cast<uint64_t>(tiles_max - 1) * steps_per_tile)
                            ? num_pairs : pairs_last;
                    const long long t0 = clock64();
                    while (ptx::ld_relaxed_gpu(p.sync_ctr + jw) < want) {
                      if (clock64() - t0 > 400000) { ok = 0; break; }
                    }
                  }
                }
                ok = __shfl_sync(0xffffffffu, ok, __ffs(__ballot_sync(0xffffffffu, issuer)) - 1);
                if (!ok) lockstep = false;
              }
            }
            fill_stage();
          }
        }
      }
    } else if (warp == 1 && rank == 0) {
      // ===================== MMA issuer (leader CTA; whole warp loops, one lane issues) ============
      constexpr uint32_t idesc = ptx::make_i8_idesc(2 * BM, BN_);
      const bool issuer = ptx::elect_one();
      uint32_t stage = 0, ph = 0, buf = 0, bph = 0;
      for (uint32_t t = pair_id; t < num_tiles; t += num_pairs) {
        for (uint32_t grp = 0; grp < p.groups; grp++)
        for (PairIter it(p); it.valid(); it.next()) {
          // plain wait: the peer's epilogue warps arrive relaxed (nothing to acquire: they only finished READING the
          // buffer), and an acquire.cluster wait costs an L1 invalidate per product
          ptx::mbar_wait(tempty_bar(buf), bph ^ 1u);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * Cfg::kBufStride;
          for (uint32_t kb = 0; kb < p.k_blocks; kb++) {
            ptx::mbar_wait(full_bar(stage), ph);            // my tiles landed
            // the peer CTA's tiles landed (relayed with a relaxed remote arrive: there is nothing to acquire -- the
            // payload was written by the async proxy into the peer's SMEM -- and an acquire.cluster wait costs an L1
            // invalidate, CCTL.IVALL, per k-step)
            ptx::mbar_wait(pfull_bar(stage), ph);
            ptx::tc_fence_after();
            const uint32_t a_smem = smem_base + stage * Cfg::kStageBytes;
            const uint64_t a_desc = ptx::make_sw128_kmajor_desc(a_smem);
            const uint64_t b_desc = ptx::make_sw128_kmajor_desc(a_smem + BM * BK);
            if (issuer) {
#pragma unroll
              for (uint32_t kk = 0; kk < BK / kUmmaK; kk++)
                ptx::mma_i8_ss_2sm(d_tmem, a_desc + kk * (kUmmaK >> 4), b_desc + kk * (kUmmaK >> 4), idesc,
                                   (kb | kk) != 0 ? 1u : 0u);
              ptx::tc_commit_2sm_mc(empty_bar(stage), 0x3);
            }
            __syncwarp();
            if (++stage == kStagesP) { stage = 0; ph ^= 1u; }
          }
          if (issuer) ptx::tc_commit_2sm_mc(tfull_bar(buf), 0x3);
          __syncwarp();
          if (++buf == kBufs) { buf = 0; bph ^= 1u; }
        }
      }
    } else if (warp == 3 && rank == 1) {
      // ===================== relay (non-leader CTA): "my stage landed" -> leader's pfull barrier =====
      uint32_t stage = 0, ph = 0;
      auto relay_steps = [&](const uint64_t steps) {
        for (uint64_t i = 0; i < steps; i++) {
          ptx::mbar_wait(full_bar(stage), ph);
          if (lane == 0) ptx::mbar_arrive_remote_relaxed(ptx::mapa(pfull_bar(stage), 0));
          __syncwarp();
          if (++stage == kStagesP) { stage = 0; ph ^= 1u; }
        }
      };
      const uint32_t tiles_mine = (num_tiles > pair_id) ? (num_tiles - pair_id + num_pairs - 1) / num_pairs : 0;
      relay_steps(static_cast<uint64_t>(tiles_mine) * steps_per_tile);
    }
  } else {
    // ===================== epilogue (both CTAs): 8 warps, FP64 accumulators in registers (+ SMEM) ==========
    ptx::reg_alloc<Cfg::kRegsEpi>();
    const uint32_t q = warp & 3u;            // TMEM lane quarter this warp may touch
    const uint32_t half = (warp - 4u) >> 2;  // which half of the quarter's columns
    // BM = 128: lane = row, TMEM column = tile column.  BM = 64 (UMMA M = 128): quarters 0, 1 hold rows 0-63 x columns
    // [0, BN/2), quarters 2, 3 the same rows x columns [BN/2, BN), both in TMEM columns [0, BN/2) of the buffer
    const uint32_t row_in_cta = BM == 128 ? q * 32u + lane : (q & 1u) * 32u + lane;
    const uint32_t col_in_tile = BM == 128 ? half * kCols : (q >> 1) * (BN_ / 2) + half * kCols;
    const bool raw = p.single_a != 0;
    uint32_t pc = 0;
    for (uint32_t t = pair_id; t < num_tiles; t += num_pairs) {
      uint32_t tm, tn, entry;
      tile_coords(p, t, tm, tn, entry);
      const uint32_t row = tm * 2 * BM + rank * BM + row_in_cta;
      const uint32_t col0 = tn * BN_ + col_in_tile;
      double acc[kRegCols];
      // this thread's spill accumulators: columns [half*kSpillCols, +kSpillCols) of the [col][row] array
      double *spill = reinterpret_cast<double *>(smem_raw + (spill_base - ptx::smem_u32(smem_raw))) +
                      static_cast<size_t>(half * kSpillCols) * 128 + (q * 32u + lane);
      // one plane product per group: 1 for a real GEMM, the reference's 4 for a complex one (src/gemm.cu:479-518),
      // each folded into C by its own finalize before the next starts from a zero accumulator
      for (uint32_t grp = 0; grp < p.groups; grp++) {
#pragma unroll
        for (uint32_t j = 0; j < kRegCols; j++) acc[j] = 0.0;
        bool first = true;
        for (PairIter it(p); it.valid(); it.next(), pc++) {
          const uint32_t buf = pc % kBufs, bph = (pc / kBufs) & 1u;
          ptx::mbar_wait(tfull_bar(buf), bph);
          ptx::tc_fence_after();
          const uint32_t taddr = tmem_base + ((q * 32u) << 16) + buf * Cfg::kBufStride + half * kCols;
          const double scale = it.scale(p.bits);
          // 16 columns per TMEM load; a tile width that is not a multiple of 32 ends with one 8-column load.  A tile whose
          // accumulators leave registers to spare (kLdPipe) keeps the NEXT load in flight while it folds the current one:
          // a short product (k <= 2048) otherwise spends more time in four load round trips than in FP64 work.
          constexpr uint32_t kChunks = (kCols + 15) / 16;
          uint32_t vbuf[kLdPipe ? 2 : 1][16];
          auto issue_ld = [&](const uint32_t c, uint32_t (&dst)[16]) {
            if (c * 16 + 16 <= kCols) ptx::tmem_ld_x16(taddr + c * 16, dst);
            else ptx::tmem_ld_x8(taddr + c * 16, dst);
          };
          if (kLdPipe) issue_ld(0, vbuf[0]);
#pragma unroll
          for (uint32_t c = 0; c < kChunks; c++) {
            const uint32_t nv = (c * 16 + 16 <= kCols) ? 16u : 8u;
            uint32_t (&v)[16] = vbuf[kLdPipe ? (c & 1u) : 0u];
            if (!kLdPipe) issue_ld(c, v);
            ptx::tmem_ld_wait();
            if (kLdPipe && c + 1 < kChunks) issue_ld(c + 1, vbuf[(c + 1) & 1u]);
            if (!raw) {
              // (double)p without I2F.F64 (a quarter-rate conversion, 15/clk/SM measured, that would make the
              // epilogue as slow as the MMAs): 2^52 + 2^31 + p is the bit pattern {0x43300000, p ^ 0x80000000};
              // subtracting 2^52 + 2^31 is exact.
#pragma unroll
              for (uint32_t g = 0; g < nv; g += 8) {
                double d[8];
#pragma unroll
                for (uint32_t j = 0; j < 8; j++)
                  d[j] = __dadd_rn(__hiloint2double(0x43300000, static_cast<int>(v[g + j] ^ 0x80000000u)),
                                   -4503601774854144.0);
                if (c * 16 + g < kRegCols) {
#pragma unroll
                  for (uint32_t j = 0; j < 8; j++)
                    acc[(c * 16 + g + j) % kRegCols] = __fma_rn(d[j], scale, acc[(c * 16 + g + j) % kRegCols]);
                } else {
#pragma unroll
                  for (uint32_t j = 0; j < 8; j++) {
                    double *sp = spill + static_cast<size_t>(c * 16 + g + j - kRegCols) * 128;
                    *sp = __fma_rn(d[j], scale, first ? 0.0 : *sp);
                  }
                }
              }
            } else if (row < p.m) {
#pragma unroll
              for (uint32_t j = 0; j < nv; j++) {
                const uint32_t col = col0 + c * 16 + j;
                if (col < p.n)
                  p.c_i32[(static_cast<size_t>(entry) * p.n + col) * p.m + row] = static_cast<int32_t>(v[j]);
              }
            }
          }
          first = false;
          ptx::tc_fence_before();
          __syncwarp();
          // "buffer drained": this warp's TMEM loads have completed (tcgen05.wait::ld above) and there is no memory
          // write to publish, so the hand-off to the leader's MMA warp needs no release fence -- a .release.cluster
          // arrive costs an ERRBAR (~1 us per product and warp, profiles/r2_small_products_ncu.txt)
          if (lane == 0) {
            if (rank == 0) ptx::mbar_arrive(tempty_bar(buf));
            else ptx::mbar_arrive_remote_relaxed(ptx::mapa(tempty_bar(buf), 0));
          }
        }
        if (!raw && row < p.m) {
          // alpha / beta by value, or read here from device memory (cuBLAS device pointer mode)
          double alpha = p.alpha, alpha_im = p.alpha_im, beta = p.beta, beta_im = p.beta_im;
          if (p.alpha_dev != nullptr) {
            alpha = __ldg(p.alpha_dev);
            beta = __ldg(p.beta_dev);
            if (p.cplx) {
              alpha_im = __ldg(p.alpha_dev + 1);
              beta_im = __ldg(p.beta_dev + 1);
            }
          }
          // reference src/gemm.cu:124-148: x = acc / 2^44 * amax[mi] * bmax[ni]
          const double am = p.amax[static_cast<size_t>(entry) * p.amax_batch + group_plane_a(p, grp) * p.amax_plane + row];
          const double *bmax = p.bmax + static_cast<size_t>(entry) * p.bmax_batch + group_plane_b(p, grp) * p.bmax_plane;
          double *cbase = p.c + static_cast<size_t>(entry) * p.c_batch * (p.cplx ? 2 : 1);
          double *crow = cbase + row;
          // complex: the coefficient of this plane product (reference src/gemm.cu:497-509):
          // (im,im) -> -alpha, (re,re) -> +alpha, (im,re) and (re,im) -> i*alpha
          const double coef_re = grp == 1 ? alpha : (grp == 0 ? -alpha : -alpha_im);
          const double coef_im = grp == 1 ? alpha_im : (grp == 0 ? -alpha_im : alpha);
#pragma unroll
          for (uint32_t j = 0; j < kCols; j++) {
            const uint32_t col = col0 + j;
            if (col < p.n) {
              const double a_j = (j < kRegCols) ? acc[j % kRegCols] : spill[static_cast<size_t>(j - kRegCols) * 128];
              double x = __dmul_rn(a_j, 0x1p-44);
              x = __dmul_rn(x, am);
              x = __dmul_rn(x, __ldg(bmax + col));
              if (p.cplx) {
                double2 *dst = reinterpret_cast<double2 *>(cbase) + static_cast<size_t>(col) * p.ldc + row;
                double2 y = make_double2(0.0, 0.0);
                if (grp == 0) {
                  if (beta != 0 || beta_im != 0) {
                    // C = beta * C (reference init_c_complex_kernel<false>, src/gemm.cu:214-222), contracted as the
                    // reference compiles it: t = y.y*b.y; x' = fma(y.x, b.x, -t); t = y.x*b.y; y' = fma(y.y, b.x, t).
                    // The reference's source reads the already-updated real part in the second product (an aliasing
                    // bug: Im(beta*C) is wrong whenever Im(beta) != 0); this uses the original one.
                    // Identical bits when Im(beta) == 0.
                    y = *dst;
                    const double yx = __fma_rn(y.x, beta, -__dmul_rn(y.y, beta_im));
                    y.y = __fma_rn(y.y, beta, __dmul_rn(y.x, beta_im));
                    y.x = yx;
                  }
                } else {
                  y = *dst;   // written by this very thread after the previous plane product
                }
                y.x = __fma_rn(x, coef_re, y.x);      // axy_complex_kernel (reference src/gemm.cu:160-186)
                y.y = __fma_rn(x, coef_im, y.y);
                *dst = y;
              } else {
                double *dst = crow + static_cast<size_t>(col) * p.ldc;
                if (beta != 0) {
                  *dst = __fma_rn(alpha, x, __dmul_rn(beta, *dst));
                } else {
                  *dst = __dmul_rn(alpha, x);
                }
              }
            }
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  ptx::cluster_arrive();
  ptx::cluster_wait();
  if (warp == 2) ptx::tmem_dealloc_2sm<512>(tmem_base);
}

// k == 0: every product is empty, C = beta * C (beta == 0: C is not read, reference src/gemm.cu:143-147); complex C:
// the beta pre-scale of the complex path (init_c_complex, src/gemm.cu:188-239, with the App. B.6 fix).  beta by value
// or read from device memory (cuBLAS device pointer mode).
__global__ void __launch_bounds__(256)
oz_scale_c_kernel(double *__restrict__ c, const size_t ldc, const uint32_t m, const uint32_t n, double beta,
                  double beta_im, const double *__restrict__ beta_dev, const bool cplx) {
  const uint32_t r = blockIdx.x * 256 + threadIdx.x, col = blockIdx.y;
  if (r >= m || col >= n) return;
  if (beta_dev != nullptr) {
    beta = __ldg(beta_dev);
    if (cplx) beta_im = __ldg(beta_dev + 1);
  }
  if (!cplx) {
    double *p = c + static_cast<size_t>(col) * ldc + r;
    *p = (beta != 0) ? __dmul_rn(beta, *p) : 0.0;
  } else {
    double2 *p = reinterpret_cast<double2 *>(c) + static_cast<size_t>(col) * ldc + r;
    double2 y = make_double2(0.0, 0.0);
    if (beta != 0 || beta_im != 0) {
      y = *p;
      const double yx = __fma_rn(y.x, beta, -__dmul_rn(y.y, beta_im));
      y.y = __fma_rn(y.y, beta, __dmul_rn(y.x, beta_im));
      y.x = yx;
    }
    *p = y;
  }
}

// Complex GEMM, "plane products first" variant (ozk_zgemm_combine): the four real plane products x = Re/Im(A) * Re/Im(B)
// were written to scratch as ordinary real products; this folds them into C exactly as the fused complex epilogue does
// (same operations in the same order per element, reference src/gemm.cu:160-239,479-518).
// x4: [4][n][m] doubles, plane e = (A plane) + 2 * (B plane): (re,re), (im,re), (re,im), (im,im).
__global__ void __launch_bounds__(256)
oz_zgemm_combine_kernel(const double *__restrict__ x4, const uint32_t m, const uint32_t n, double alpha, double alpha_im,
                        double beta, double beta_im, const double *__restrict__ alpha_dev,
                        const double *__restrict__ beta_dev, double2 *__restrict__ c, const size_t ldc) {
  const uint32_t r = blockIdx.x * 256 + threadIdx.x, col = blockIdx.y;
  if (r >= m || col >= n) return;
  if (alpha_dev != nullptr) {
    alpha = __ldg(alpha_dev), alpha_im = __ldg(alpha_dev + 1);
    beta = __ldg(beta_dev), beta_im = __ldg(beta_dev + 1);
  }
  const size_t plane = static_cast<size_t>(m) * n, at = static_cast<size_t>(col) * m + r;
  double2 *dst = c + static_cast<size_t>(col) * ldc + r;
  double2 y = make_double2(0.0, 0.0);
  if (beta != 0 || beta_im != 0) {
    y = *dst;
    const double yx = __fma_rn(y.x, beta, -__dmul_rn(y.y, beta_im));
    y.y = __fma_rn(y.y, beta, __dmul_rn(y.x, beta_im));
    y.x = yx;
  }
  // (im,im) -> -alpha, (re,re) -> +alpha, (im,re), (re,im) -> i*alpha
  double x = __ldg(x4 + 3 * plane + at);
  y.x = __fma_rn(x, -alpha, y.x);
  y.y = __fma_rn(x, -alpha_im, y.y);
  x = __ldg(x4 + at);
  y.x = __fma_rn(x, alpha, y.x);
  y.y = __fma_rn(x, alpha_im, y.y);
  x = __ldg(x4 + plane + at);
  y.x = __fma_rn(x, -alpha_im, y.x);
  y.y = __fma_rn(x, alpha, y.y);
  x = __ldg(x4 + 2 * plane + at);
  y.x = __fma_rn(x, -alpha_im, y.x);
  y.y = __fma_rn(x, alpha, y.y);
  *dst = y;
}

// ---- host side --------------------------------------------------------------------------------
// One-time, per-device kernel setup (dynamic SMEM opt-in, resident cluster count), safe for concurrent
// callers and for one process driving several GPUs.
constexpr int kMaxDevices = 64;
struct PerDeviceOnce {
  std::mutex mu;
  bool done[kMaxDevices] = {};
  int value[kMaxDevices] = {};
};

// 0 = default (tile width chosen per problem); 128 ... 256 = forced (test/tuning hook, see ozk_set_cluster_shape)
int g_tile_override = 0;
// 128 x 128 tiles (64 rows per CTA, UMMA M = 128): -1 = chosen per problem, 1 = forced, 0 = never (a forced tile width)
int g_half_rows_override = -1;

// ---- tuning knob (read once; OZIMMU_B200_LOCKSTEP overrides) ----
// rasterisation band height in tiles (OZIMMU_B200_GROUP_M, read once; default 8: a wave of 74 pairs covers 8 x 9.25
// tiles, the squarest block and with it the smallest panel footprint per product pass)
uint32_t raster_group_m() {
  static const uint32_t g = [] {
    uint32_t v = 8;
    if (const char *e = std::getenv("OZIMMU_B200_GROUP_M")) v = static_cast<uint32_t>(std::atoi(e));
    return v == 0 ? 8u : v;
  }();
  return g;
}

uint32_t lockstep_window() {
  static const uint32_t w = [] {
    uint32_t v = 2;
    if (const char *e = std::getenv("OZIMMU_B200_LOCKSTEP")) v = static_cast<uint32_t>(std::atoi(e));
    return v;
  }();
  return w;
}

constexpr uint32_t kSyncCounters = 1u << 16;  // per buffer: 64 Ki intervals = 1 Mi k-steps per CTA pair
constexpr int kSyncBuffers = 8;               // launches that may be in flight at once without sharing
// per device: a small ring of counter buffers, allocated TOGETHER on the first paced launch on that device and kept for
// the life of the process (allocating them one by one made each of the first eight multi-round launches pay a
// cudaMalloc, i.e. a device synchronisation: 2048^3 measured 0.56 ms instead of 0.37 ms over its first ten calls,
// profiles/r2_sizes_vs_reference.txt).  More than kSyncBuffers launches in flight at once share counters, which only
// costs pacing time-outs.
uint32_t *next_sync_buffer() {
  static std::mutex mu;
  static uint32_t *pool[kMaxDevices] = {};
  static bool failed[kMaxDevices] = {};
  static int next[kMaxDevices] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  if (pool[dev] == nullptr) {
    if (failed[dev]) return nullptr;
    if (cudaMalloc(&pool[dev], static_cast<size_t>(kSyncBuffers) * kSyncCounters * sizeof(uint32_t)) != cudaSuccess) {
      pool[dev] = nullptr;
      failed[dev] = true;
      cudaGetLastError();
      return nullptr;
    }
  }
  const int i = next[dev];
  next[dev] = (i + 1) % kSyncBuffers;
  return pool[dev] + static_cast<size_t>(i) * kSyncCounters;
}

template <uint32_t BN_, uint32_t BMC_ = 128>
int launch_pair(const FusedParams &p0, cudaStream_t stream) {
  using Cfg = PairCfg<BN_, BMC_>;
  constexpr uint32_t BM = BMC_;
  FusedParams p = p0;
  p.tiles_m = ceil_div_u32(p.m, 2 * BM);
  p.tiles_n = ceil_div_u32(p.n, BN_);
  p.group_m = raster_group_m();
  if (p.rt_a == 0) p.rt_a = static_cast<uint32_t>(slice_row_tiles(p.m));  // != 0: block of a larger plane
  if (p.rt_b == 0) p.rt_b = static_cast<uint32_t>(slice_row_tiles(p.n));
  if (p.b_rows == 0) p.b_rows = p.rt_b * static_cast<uint32_t>(kTileRows);
  p.sync_window = lockstep_window();

  auto kern = oz_gemm_pair_kernel<BN_, BMC_>;
  int dev = 0, sms = 0;
  OZ_CUDA_TRY(cudaGetDevice(&dev));
  OZ_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (dev < 0 || dev >= kMaxDevices) return static_cast<int>(cudaErrorInvalidDevice);

  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  static PerDeviceOnce once;
  int max_pairs = 0;
  {
    std::lock_guard<std::mutex> lock(once.mu);
    if (!once.done[dev]) {
      OZ_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
      cfg.gridDim = dim3(static_cast<unsigned>(sms) / 2 * 2);
      int nc = 0;
      if (cudaOccupancyMaxActiveClusters(&nc, kern, &cfg) == cudaSuccess && nc > 0) once.value[dev] = nc;
      else once.value[dev] = sms / 2;
      cudaGetLastError();
      once.done[dev] = true;
      if (std::getenv("OZIMMU_B200_DEBUG"))
        std::fprintf(stderr, "[ozimmu_b200] pair kernel %ux%u: %d CTA pairs resident (%d SMs)\n", 2 * BM, BN_, once.value[dev], sms);
    }
    max_pairs = once.value[dev];
  }
  if (p.batch == 0) p.batch = 1;
  const uint32_t num_tiles = p.tiles_m * p.tiles_n * p.batch;
  if (num_tiles == 0) return 0;
  // Every CTA pair gets the same number of tiles (+-1): the launch needs `rounds` rounds either way, and a pair with a
  // short tile list would only exit early when the launch's own grid is wider than needed.  Launching
  // ceil(tiles / rounds) pairs instead leaves the other SMs to a concurrent launch for the WHOLE duration (the block
  // pipelines rotate launches over several streams); with the full grid and 128 tiles, 20 pairs run one tile and 54
  // run two, and whichever CTAs of the next launch start late still own two tiles -- 2 rounds per 1.73 rounds of work.
  const uint32_t rounds = ceil_div_u32(num_tiles, static_cast<uint32_t>(max_pairs));
  // one_tile_per_pair: a non-persistent launch, one CTA pair per tile.  The hardware CTA scheduler then is the tile
  // queue: pairs of concurrent launches fill the SMs in launch order with no ragged last round per launch, and a
  // higher-priority kernel (the split of the operand block that has just arrived) gets SMs whenever a tile ends
  // instead of waiting for a persistent pair's whole tile list.
  const uint32_t pairs = p.one_tile_per_pair ? num_tiles : ceil_div_u32(num_tiles, rounds);
  cfg.gridDim = dim3(pairs * 2);
  // lockstep counters only pay off when several rounds of tiles stream through L2 (pairs cannot drift apart
  // within a single round, and the polling costs ~7 % there)
  p.sync_ctr = nullptr;
  if (p.sync_window > 0 && pairs > 1 && num_tiles > pairs && p.single_a == 0 && !p.no_lockstep) {
    const uint64_t steps = static_cast<uint64_t>(ceil_div_u32(num_tiles, pairs)) * (p.num_split * (p.num_split + 1) / 2) * p.k_blocks * p.groups;
    const uint64_t need = steps / kSyncEvery + 1;
    if (need <= kSyncCounters) {
      uint32_t *buf = next_sync_buffer();
      if (buf) {
        OZ_CUDA_TRY(cudaMemsetAsync(buf, 0, need * sizeof(uint32_t), stream));
        p.sync_ctr = buf;
        p.sync_len = static_cast<uint32_t>(need);
      }
    }
  }
  OZ_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, p));
  count_launch(1);
  return 0;
}

// tile widths built into the library (UMMA N of the CTA pair): 256 and 128 are the defaults, the widths in between
// trade MACs per delivered byte for a deeper operand ring (PairCfg) and a different tile count
constexpr int kTileWidths[] = {256, 240, 224, 208, 192, 128};

bool tile_width_ok(int bn) {
  for (int w : kTileWidths)
    if (w == bn) return true;
  return false;
}

// OZIMMU_B200_TILE_N=<width>: force the tile width (read once)
int env_tile_width() {
  static const int v = [] {
    const char *e = std::getenv("OZIMMU_B200_TILE_N");
    const int w = e ? std::atoi(e) : 0;
    return tile_width_ok(w) ? w : 0;
  }();
  return v;
}

// ---- which tile for which problem ------------------------------------------------------------------------------
// Measured time of one round of tiles -- 72 of the 74 CTA pairs busy, s = 9 -- per tile shape and k, in microseconds
// (tools/tile_cost_probe.py, profiles/r2_tile_cost_probe.txt): `first` = a launch of one round, `next` = what a second
// round adds (128-wide tiles lose ~14 % at k = 8192 once several rounds stream through L2 without lockstep help; the
// k = 8192 `next` column is the per-round time of whole 8192^3 products, 14-19 rounds, profiles/r2_sweep_tile_width.txt).  Per
// area of C the narrow tiles win at small k (the per-product epilogue weighs more than the operand bytes: 128-wide costs
// 0.79 x of 256-wide at k = 1024) and lose at large k.  Other split counts scale every entry alike.
struct TileCost {
  int rows, width;          // rows of C per CTA pair, tile width
  float first[4], next[4];  // k = 1024, 2048, 4096, 8192
};
constexpr TileCost kTileCosts[] = {
    {256, 256, {250, 360, 634, 1243}, {254, 368, 634, 1280}}, {256, 240, {232, 340, 585, 1150}, {233, 340, 583, 1190}},
    {256, 224, {195, 315, 558, 1087}, {239, 317, 607, 1200}}, {256, 208, {170, 285, 535, 1024}, {205, 286, 535, 1080}},
    {256, 192, {155, 257, 498, 946}, {164, 259, 500, 1020}},  {256, 128, {99, 157, 321, 631}, {93, 152, 345, 718}},
    {128, 128, {71, 124, 273, 579}, {81, 157, 368, 770}},
};
// piecewise linear in k between the measured points, proportional below the first and along the last slope beyond
double interp_k(const float (&v)[4], double k) {
  const double ks[4] = {1024, 2048, 4096, 8192};
  if (k <= ks[0]) return v[0] * (0.35 + 0.65 * k / ks[0]);   // a tile keeps ~1/3 of its k = 1024 time as k -> 0
  for (int i = 0; i < 3; i++)
    if (k <= ks[i + 1]) return v[i] + (v[i + 1] - v[i]) * (k - ks[i]) / (ks[i + 1] - ks[i]);
  return v[3] + (v[3] - v[2]) * (k - ks[3]) / (ks[3] - ks[2]);
}

// SM count of the current device (queried once per device: the choice below runs on every launch)
int current_device_sms() {
  static int cached[kMaxDevices] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 148;
  if (cached[dev] == 0) {
    int sms = 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
    cached[dev] = sms;   // benign race: every thread writes the same value
  }
  return cached[dev];
}

// the tile with the lowest modelled cost: (rows of C per CTA pair, width)
void choose_tile(uint32_t m, uint32_t n, double k, double batch, bool one_tile_per_pair, int sms, bool allow_half_rows,
                 int &rows_out, int &width_out) {
  // A persistent launch costs first + (rounds - 1) x next with rounds = ceil(tiles / resident pairs); a one-tile-per-
  // pair launch shares the GPU with its neighbours in the block pipelines, so what counts is the SM time of its tiles:
  // tiles x first / (tiles of the measured round).  (Round 1 / early round 2 used rounds x (55 + width), fitted at k = 8192 only.)
  const double pairs = static_cast<double>(sms > 1 ? sms / 2 : 1);
  double best_cost = 1e300;
  rows_out = 256, width_out = 256;
  for (const TileCost &t : kTileCosts) {   // ties go to the earlier (wider) entry
    if (t.rows == 128 && !allow_half_rows) continue;
    const double tiles = static_cast<double>(ceil_div_u32(m, t.rows)) * ceil_div_u32(n, t.width) * batch;
    const double rounds = std::ceil(tiles / pairs);
    const double probe_tiles = t.rows == 128 ? 64.0 : 72.0;   // tiles in the measured round
    const double cost = one_tile_per_pair ? tiles * interp_k(t.first, k) / probe_tiles
                                          : interp_k(t.first, k) + (rounds - 1) * interp_k(t.next, k);
    if (cost < best_cost) {
      best_cost = cost;
      rows_out = t.rows;
      width_out = t.width;
    }
  }
}

int dispatch_fused(const FusedParams &p, cudaStream_t stream) {
  int bn = g_tile_override ? g_tile_override : env_tile_width();
  bool half_rows = g_half_rows_override > 0;
  if (bn == 0 && g_half_rows_override <= 0) {
    int rows = 256;
    choose_tile(p.m, p.n, static_cast<double>(p.k_blocks) * BK, p.batch ? p.batch : 1, p.one_tile_per_pair != 0,
                current_device_sms(), g_half_rows_override < 0, rows, bn);
    half_rows = rows == 128;
  }
  if (half_rows) return launch_pair<128, 64>(p, stream);
  switch (bn) {
    case 256: return launch_pair<256>(p, stream);
    case 240: return launch_pair<240>(p, stream);
    case 224: return launch_pair<224>(p, stream);
    case 208: return launch_pair<208>(p, stream);
    case 192: return launch_pair<192>(p, stream);
    case 128: return launch_pair<128>(p, stream);
    default: return static_cast<int>(cudaErrorInvalidValue);
  }
}

bool valid_common(size_t m, size_t n, size_t k, size_t pitch, unsigned num_split, unsigned bits) {
  return m > 0 && n > 0 && k > 0 && m < (1ull << 31) && n < (1ull << 31) && pitch % 128 == 0 &&
         pitch >= k && pitch < (1ull << 31) && num_split >= 1 && num_split <= 18 && bits >= 1 && bits <= 7;
}

FusedParams base_params(size_t m, size_t n, size_t pitch, const int8_t *a_slices, const int8_t *b_slices,
                        unsigned num_split, unsigned bits) {
  FusedParams p{};
  p.m = static_cast<uint32_t>(m);
  p.n = static_cast<uint32_t>(n);
  p.k_blocks = static_cast<uint32_t>(pitch / BK);
  p.num_split = num_split;
  p.bits = static_cast<int32_t>(bits);
  p.a_slices = a_slices;
  p.b_slices = b_slices;
  p.groups = 1;
  return p;
}

}  // namespace
}  // namespace oz

// Test/tuning hook: force the tile of the fused kernel (see ozimmu_b200.h); anything else restores the per-problem choice.
extern "C" int ozk_fused_tile_choice(size_t m, size_t n, size_t k, size_t batch, int one_tile_per_pair, int sms, int *rows,
                                     int *width) {
  if (rows == nullptr || width == nullptr) return 1;
  if (sms <= 0) sms = oz::current_device_sms();
  const size_t pitch = (k + 127) / 128 * 128;
  oz::choose_tile(static_cast<uint32_t>(m), static_cast<uint32_t>(n), static_cast<double>(pitch),
                  static_cast<double>(batch ? batch : 1), one_tile_per_pair != 0, sms, true, *rows, *width);
  cudaGetLastError();   // a process without a device: the SM-count query failed, which is fine here
  return 0;
}

extern "C" int ozk_set_cluster_shape(int cm, int cn) {
  oz::g_tile_override = (cm == 0 && oz::tile_width_ok(cn)) ? cn : 0;
  // (64, 128): force the 128 x 128 tile (64 rows per CTA); a forced width excludes it; anything else: per-problem choice
  oz::g_half_rows_override = (cm == 64 && cn == 128) ? 1 : (oz::g_tile_override ? 0 : -1);
  return 0;
}

// The general launcher: every other ozk_gemm_i8_fused* entry point fills this struct.
extern "C" int ozk_gemm_i8_fused_ex(const ozk_fused_args_t *a, void *stream) {
  if (a == nullptr) return static_cast<int>(cudaErrorInvalidValue);
  const size_t batch = a->batch ? a->batch : 1;
  if (a->m == 0 || a->n == 0) return 0;
  const size_t a_plane_rows = a->a_plane_rows ? a->a_plane_rows : a->m;
  const size_t b_plane_rows = a->b_plane_rows ? a->b_plane_rows : a->n;
  const size_t tiles = ((a->m + 255) / 256) * ((a->n + 127) / 128);  // upper bound (128-wide tiles)
  if (!oz::valid_common(a->m, a->n, a->k, a->pitch, a->num_split, a->bits_per_int8) || a->ldc < a->m ||
      a->row0 % 256 != 0 || a->col0 % 256 != 0 || a->row0 + a->m > a_plane_rows || a->col0 + a->n > b_plane_rows ||
      a_plane_rows >= (1ull << 31) || b_plane_rows >= (1ull << 31) || a->a_batch_bytes % 16 != 0 ||
      a->b_batch_bytes % 16 != 0 || a->a_plane_bytes % 16 != 0 || a->b_plane_bytes % 16 != 0 || batch >= (1ull << 31) ||
      tiles * batch >= (1ull << 31) || (a->alpha_dev == nullptr) != (a->beta_dev == nullptr) || a->c == nullptr)
    return static_cast<int>(cudaErrorInvalidValue);
  // a block starts on a row-tile boundary of its plane: the kernel only needs the plane's slice stride
  const size_t tile_row_bytes = a->pitch * oz::kTileRows;
  oz::FusedParams p = oz::base_params(a->m, a->n, a->pitch, a->a_slices + (a->row0 / oz::kTileRows) * tile_row_bytes,
                                      a->b_slices + (a->col0 / oz::kTileRows) * tile_row_bytes, a->num_split,
                                      a->bits_per_int8);
  p.rt_a = static_cast<uint32_t>(oz::slice_row_tiles(a_plane_rows));
  p.rt_b = static_cast<uint32_t>(oz::slice_row_tiles(b_plane_rows));
  p.b_rows = p.rt_b * static_cast<uint32_t>(oz::kTileRows) - static_cast<uint32_t>(a->col0);
  // entries of a batch share no operands: nothing to gain from pacing the CTA pairs
  p.no_lockstep = ((a->flags & OZK_FUSED_NO_LOCKSTEP) || batch > 1) ? 1u : 0u;
  p.one_tile_per_pair = (a->flags & OZK_FUSED_ONE_TILE_PER_PAIR) ? 1u : 0u;
  p.alpha = a->alpha[0];
  p.alpha_im = a->alpha[1];
  p.beta = a->beta[0];
  p.beta_im = a->beta[1];
  p.alpha_dev = a->alpha_dev;
  p.beta_dev = a->beta_dev;
  p.c = static_cast<double *>(a->c);
  p.ldc = a->ldc;
  p.amax = a->amax;
  p.bmax = a->bmax;
  p.batch = static_cast<uint32_t>(batch);
  p.a_batch_bytes = a->a_batch_bytes;
  p.b_batch_bytes = a->b_batch_bytes;
  p.amax_batch = a->amax_batch;
  p.bmax_batch = a->bmax_batch;
  p.c_batch = a->c_batch;
  if (a->complex_c) {
    p.cplx = 1;
    p.groups = 4;
    p.a_plane_bytes = a->a_plane_bytes;
    p.b_plane_bytes = a->b_plane_bytes;
    p.amax_plane = a->amax_plane;
    p.bmax_plane = a->bmax_plane;
  }
  return oz::dispatch_fused(p, static_cast<cudaStream_t>(stream));
}

namespace {
ozk_fused_args_t real_args(size_t m, size_t n, size_t k, const int8_t *a_slices, const int8_t *b_slices, size_t pitch,
                           const double *amax, const double *bmax, unsigned num_split, unsigned bits_per_int8,
                           double alpha, double beta, double *c, size_t ldc) {
  ozk_fused_args_t a{};
  a.m = m, a.n = n, a.k = k, a.pitch = pitch;
  a.a_slices = a_slices, a.b_slices = b_slices;
  a.amax = amax, a.bmax = bmax;
  a.num_split = num_split, a.bits_per_int8 = bits_per_int8;
  a.alpha[0] = alpha, a.beta[0] = beta;
  a.c = c, a.ldc = ldc;
  return a;
}
}  // namespace

extern "C" int ozk_gemm_i8_fused(size_t m, size_t n, size_t k, const int8_t *a_slices,
                                 const int8_t *b_slices, size_t pitch, const double *amax,
                                 const double *bmax, unsigned num_split, unsigned bits_per_int8,
                                 double alpha, double beta, double *c, size_t ldc, void *stream) {
  const ozk_fused_args_t a = real_args(m, n, k, a_slices, b_slices, pitch, amax, bmax, num_split, bits_per_int8, alpha,
                                       beta, c, ldc);
  return ozk_gemm_i8_fused_ex(&a, stream);
}

extern "C" int ozk_gemm_i8_fused_block(size_t m, size_t n, size_t k, const int8_t *a_slices, size_t a_plane_rows,
                                       size_t row0, const int8_t *b_slices, size_t b_plane_rows, size_t col0,
                                       size_t pitch, const double *amax, const double *bmax, unsigned num_split,
                                       unsigned bits_per_int8, double alpha, double beta, double *c, size_t ldc,
                                       unsigned flags, void *stream) {
  ozk_fused_args_t a = real_args(m, n, k, a_slices, b_slices, pitch, amax, bmax, num_split, bits_per_int8, alpha, beta,
                                 c, ldc);
  a.a_plane_rows = a_plane_rows, a.b_plane_rows = b_plane_rows;
  a.row0 = row0, a.col0 = col0;
  a.flags = flags;
  return ozk_gemm_i8_fused_ex(&a, stream);
}

extern "C" int ozk_gemm_i8_fused_batched(size_t m, size_t n, size_t k, size_t batch, const int8_t *a_slices,
                                         size_t a_batch_bytes, const int8_t *b_slices, size_t b_batch_bytes,
                                         size_t pitch, const double *amax, size_t amax_batch, const double *bmax,
                                         size_t bmax_batch, unsigned num_split, unsigned bits_per_int8,
                                         double alpha, double beta, double *c, size_t ldc, size_t c_batch,
                                         void *stream) {
  if (batch == 0) return 0;
  ozk_fused_args_t a = real_args(m, n, k, a_slices, b_slices, pitch, amax, bmax, num_split, bits_per_int8, alpha, beta,
                                 c, ldc);
  a.batch = batch;
  a.a_batch_bytes = a_batch_bytes, a.b_batch_bytes = b_batch_bytes;
  a.amax_batch = amax_batch, a.bmax_batch = bmax_batch, a.c_batch = c_batch;
  return ozk_gemm_i8_fused_ex(&a, stream);
}

extern "C" int ozk_scale_c_ex(size_t m, size_t n, const double beta[2], const double *beta_dev, int complex_c, void *c,
                              size_t ldc, void *stream) {
  if (m == 0 || n == 0) return 0;
  if (ldc < m || m >= (1ull << 31) || n >= 65536ull * 32768ull || c == nullptr || (beta == nullptr && beta_dev == nullptr))
    return static_cast<int>(cudaErrorInvalidValue);
  const size_t es = complex_c ? 2 : 1;
  for (size_t j0 = 0; j0 < n; j0 += 65535) {
    const unsigned nj = static_cast<unsigned>(n - j0 < 65535 ? n - j0 : 65535);
    dim3 grid(static_cast<unsigned>((m + 255) / 256), nj);
    oz::oz_scale_c_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<double *>(c) + j0 * ldc * es, ldc, static_cast<uint32_t>(m), nj, beta ? beta[0] : 0.0,
        (beta && complex_c) ? beta[1] : 0.0, beta_dev, complex_c != 0);
    oz::count_launch(1);
  }
  return static_cast<int>(cudaGetLastError());
}

extern "C" int ozk_zgemm_combine(size_t m, size_t n, const double *x4, const double alpha[2], const double beta[2],
                                 const double *alpha_dev, const double *beta_dev, void *c, size_t ldc, void *stream) {
  if (m == 0 || n == 0) return 0;
  if (ldc < m || m >= (1ull << 31) || n >= 65536ull * 32768ull || c == nullptr || x4 == nullptr ||
      (alpha_dev == nullptr) != (beta_dev == nullptr) || (alpha_dev == nullptr && (alpha == nullptr || beta == nullptr)))
    return static_cast<int>(cudaErrorInvalidValue);
  for (size_t j0 = 0; j0 < n; j0 += 65535) {
    const unsigned nj = static_cast<unsigned>(n - j0 < 65535 ? n - j0 : 65535);
    dim3 grid(static_cast<unsigned>((m + 255) / 256), nj);
    // the column offset applies to every scratch plane alike: pass the full n so that the plane stride stays m * n
    oz::oz_zgemm_combine_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x4 + j0 * m, static_cast<uint32_t>(m), static_cast<uint32_t>(n), alpha ? alpha[0] : 0.0, alpha ? alpha[1] : 0.0,
        beta ? beta[0] : 0.0, beta ? beta[1] : 0.0, alpha_dev, beta_dev, static_cast<double2 *>(c) + j0 * ldc, ldc);
    oz::count_launch(1);
  }
  return static_cast<int>(cudaGetLastError());
}

extern "C" int ozk_scale_c(size_t m, size_t n, double beta, double *c, size_t ldc, void *stream) {
  const double b[2] = {beta, 0.0};
  return ozk_scale_c_ex(m, n, b, nullptr, 0, c, ldc, stream);
}

extern "C" int ozk_gemm_i8_pair(size_t m, size_t n, size_t k, const int8_t *a_slices,
                                const int8_t *b_slices, size_t pitch, unsigned num_split,
                                unsigned a_id, unsigned b_id, int32_t *c_i32, void *stream) {
  if (m == 0 || n == 0) return 0;
  if (!oz::valid_common(m, n, k, pitch, num_split, 7) || a_id < 1 || b_id < 1 || a_id > num_split ||
      b_id > num_split)
    return static_cast<int>(cudaErrorInvalidValue);
  oz::FusedParams p = oz::base_params(m, n, pitch, a_slices, b_slices, num_split, 7);
  p.single_a = a_id;
  p.single_b = b_id;
  p.c_i32 = c_i32;
  return oz::dispatch_fused(p, static_cast<cudaStream_t>(stream));
}
