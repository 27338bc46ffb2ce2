// gemm_fused.cu -- the hot path: s(s+1)/2 exact int8 x int8 -> int32 slice products on tcgen05
// tensor cores, each folded into a per-element FP64 accumulator in the reference's exact op
// order, then the exponent rescale and alpha/beta -- ONE persistent kernel.
//
// Replaces, fused: reference src/gemm.cu:266-334 (matmul_core -> cublasGemmEx int8, 45 library
// GEMMs at s=9), :77-102 (accumulate_in_f64, one 20 B/element HBM pass per product), :104-122
// (init_accumulator_buffer) and :124-158 (axby).  Arithmetic per output element follows SURVEY
// App. A.4/A.5 exactly (same FMA sequence => bit-identical C).
//
// Design (B200 / sm_100a):
//  * Persistent CTAs, one per SM.  Each CTA owns a 128x128 tile of C for the whole pair list, so
//    the FP64 accumulator never leaves the SM: it lives in the registers of 8 epilogue warps
//    (256 threads x 64 doubles).  HBM sees the int8 slices (mostly from L2) and C exactly once.
//  * Warp roles: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (one thread), warp 2 = TMEM
//    allocator, warps 4-11 = epilogue.  setmaxnreg moves registers from warps 0-3 to the epilogue.
//  * Operands: K-major int8 slices [slice][row][pitch]; TMA (3-D tensor map: k, row, slice) stages
//    128x128-byte tiles with the 128-byte swizzle into a 6-deep SMEM ring; UMMA M=128,N=128,K=32
//    kind::i8 accumulates a full-K product into one of 4 TMEM buffers (128 columns each), so the
//    MMA pipe runs up to 3 products ahead of the epilogue.
//  * Optional thread-block clusters (CM x CN CTAs): the A tile is TMA-multicast across the CN CTAs
//    that share a row block and the B tile across the CM CTAs that share a column block, cutting
//    L2->SMEM traffic per SM from 128 to 128*(1/CN+1/CM)/2 bytes per MMA-cycle.
//  * Epilogue per product: tcgen05.ld 64 int32 columns of the thread's row, release the TMEM
//    buffer, acc = fma((double)p, 2^(32-rshift), acc).  After the last pair:
//    x = acc*2^-44*amax[r]*bmax[c]; C = alpha*x (+ beta*C), coalesced along rows of column-major C.
#include <cstdlib>
#include <mutex>

#include <cuda.h>
#include <cuda_runtime.h>

#include "oz_common.cuh"
#include "ozimmu_b200.h"
#include "ptx.cuh"

namespace oz {
namespace {

constexpr uint32_t BM = 128, BN = 128, BK = 128;  // tile (BK in bytes == int8 elements)
constexpr uint32_t kStages = 6;
constexpr uint32_t kAccBufs = 4;                  // TMEM accumulator ring (4 x 128 columns = 512)
constexpr uint32_t kThreads = 384;
constexpr uint32_t kEpiWarps = 8;
constexpr uint32_t kStageBytes = (BM + BN) * BK;  // 32 KB
constexpr uint32_t kBarBytes = 8 * (2 * kStages + 2 * kAccBufs) + 16;
constexpr uint32_t kSmemBytes = kStages * kStageBytes + kBarBytes + 1024;  // +1024: manual alignment
constexpr uint32_t kUmmaK = 32;

struct FusedParams {
  uint32_t m, n;
  uint32_t k_blocks;           // ceil(pitch / BK)
  uint32_t num_split;
  int32_t bits;                // L = bits per int8 slice
  uint32_t single_a, single_b; // != 0: raw mode, only this (1-based) pair, int32 output
  uint32_t super_m, super_n;   // cluster-tile grid
  uint32_t group_m;            // rasterisation band height in cluster tiles
  double alpha, beta;
  double *c;
  unsigned long long ldc;
  const double *amax;
  const double *bmax;
  int32_t *c_i32;
  // complex accumulation (reference src/gemm.cu:160-239,479-518): C is cuDoubleComplex (ldc in complex
  // elements) and this launch adds one of the four real products: y = fma(x, (alpha, alpha_im), y),
  // after y = beta * y (the reference's init_c_complex, applied once, by the first of the four launches)
  uint32_t cplx, cplx_init;
  double alpha_im, beta_im;
  // pair kernel tuning (see launch_pair): L2 prefetch lead in k-blocks; soft lockstep between CTA pairs
  uint32_t prefetch_ahead;
  uint32_t *sync_ctr;          // one arrival counter per kSyncEvery k-steps, zeroed before launch (or null)
  uint32_t sync_window;        // a pair starts sync interval j only after interval j - window is complete
  uint32_t sync_len;           // counters available
  uint32_t idle_sleep_ns;      // __nanosleep between barrier probes of the idle roles (0 = spin)
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

__device__ __forceinline__ void super_tile_coords(const FusedParams &p, uint32_t st, uint32_t &sm,
                                                  uint32_t &sn) {
  const uint32_t group_size = p.group_m * p.super_n;
  const uint32_t g = st / group_size;
  const uint32_t first = g * p.group_m;
  const uint32_t rows = min(p.super_m - first, p.group_m);
  const uint32_t r = st - g * group_size;
  sm = first + r % rows;
  sn = r / rows;
}

template <uint32_t CM, uint32_t CN>
__global__ void __launch_bounds__(kThreads, 1)
oz_gemm_fused_kernel(const __grid_constant__ CUtensorMap tmap_a,
                     const __grid_constant__ CUtensorMap tmap_b, const FusedParams p) {
  constexpr uint32_t CSZ = CM * CN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + kStages * kStageBytes;
  auto full_bar = [&](uint32_t s) { return bar_base + 8u * s; };
  auto empty_bar = [&](uint32_t s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](uint32_t b) { return bar_base + 8u * (2 * kStages + b); };
  auto tempty_bar = [&](uint32_t b) { return bar_base + 8u * (2 * kStages + kAccBufs + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 2 * kAccBufs);
  volatile uint32_t *tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t crank = (CSZ > 1) ? ptx::cluster_ctarank() : 0u;
  const uint32_t cm = crank % CM, cn = crank / CM;
  const uint32_t cluster_id = blockIdx.x / CSZ;
  const uint32_t num_clusters = gridDim.x / CSZ;
  const uint32_t num_super = p.super_m * p.super_n;

  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < kStages; s++) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), CM + CN - 1);
    }
    for (uint32_t b = 0; b < kAccBufs; b++) {
      ptx::mbar_init(tfull_bar(b), 1);
      ptx::mbar_init(tempty_bar(b), kEpiWarps);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_b);
  }
  if (warp == 2) ptx::tmem_alloc<512>(tmem_slot);
  ptx::tc_fence_before();
  if (CSZ > 1) {
    ptx::cluster_arrive();
    ptx::cluster_wait();
  } else {
    __syncthreads();
  }
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp < 4) {
    ptx::reg_dealloc<56>();
    if (warp == 0 && lane == 0) {
      // ===================== TMA producer =====================
      uint16_t mask_a = 0, mask_b = 0;
      for (uint32_t j = 0; j < CN; j++) mask_a |= static_cast<uint16_t>(1u << (cm + CM * j));
      for (uint32_t i = 0; i < CM; i++) mask_b |= static_cast<uint16_t>(1u << (i + CM * cn));
      uint32_t ks = 0;
      for (uint32_t st = cluster_id; st < num_super; st += num_clusters) {
        uint32_t sm, sn;
        super_tile_coords(p, st, sm, sn);
        const int row_a = static_cast<int>((sm * CM + cm) * BM + cn * (BM / CN));
        const int row_b = static_cast<int>((sn * CN + cn) * BN + cm * (BN / CM));
        for (PairIter it(p); it.valid(); it.next()) {
          const int sa = static_cast<int>(it.a_id() - 1), sb = static_cast<int>(it.b_id() - 1);
          for (uint32_t kb = 0; kb < p.k_blocks; kb++, ks++) {
            const uint32_t stage = ks % kStages, ph = (ks / kStages) & 1u;
            ptx::mbar_wait(empty_bar(stage), ph ^ 1u);
            ptx::mbar_expect_tx(full_bar(stage), kStageBytes);
            const uint32_t a_dst = smem_base + stage * kStageBytes + cn * (BM / CN) * BK;
            const uint32_t b_dst = smem_base + stage * kStageBytes + BM * BK + cm * (BN / CM) * BK;
            if (CN > 1) {
              ptx::tma_load_3d_mc(a_dst, &tmap_a, full_bar(stage), static_cast<int>(kb * BK), row_a, sa, mask_a);
            } else {
              ptx::tma_load_3d(a_dst, &tmap_a, full_bar(stage), static_cast<int>(kb * BK), row_a, sa);
            }
            if (CM > 1) {
              ptx::tma_load_3d_mc(b_dst, &tmap_b, full_bar(stage), static_cast<int>(kb * BK), row_b, sb, mask_b);
            } else {
              ptx::tma_load_3d(b_dst, &tmap_b, full_bar(stage), static_cast<int>(kb * BK), row_b, sb);
            }
          }
        }
      }
    } else if (warp == 1 && lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc = ptx::make_i8_idesc(BM, BN);
      uint16_t mask_e = 0;
      for (uint32_t j = 0; j < CN; j++) mask_e |= static_cast<uint16_t>(1u << (cm + CM * j));
      for (uint32_t i = 0; i < CM; i++) mask_e |= static_cast<uint16_t>(1u << (i + CM * cn));
      uint32_t ks = 0, pc = 0;
      for (uint32_t st = cluster_id; st < num_super; st += num_clusters) {
        for (PairIter it(p); it.valid(); it.next(), pc++) {
          const uint32_t buf = pc % kAccBufs, bph = (pc / kAccBufs) & 1u;
          ptx::mbar_wait(tempty_bar(buf), bph ^ 1u);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * BN;
          for (uint32_t kb = 0; kb < p.k_blocks; kb++, ks++) {
            const uint32_t stage = ks % kStages, ph = (ks / kStages) & 1u;
            ptx::mbar_wait(full_bar(stage), ph);
            ptx::tc_fence_after();
            const uint32_t a_smem = smem_base + stage * kStageBytes;
            const uint64_t a_desc = ptx::make_sw128_kmajor_desc(a_smem);
            const uint64_t b_desc = ptx::make_sw128_kmajor_desc(a_smem + BM * BK);
#pragma unroll
            for (uint32_t kk = 0; kk < BK / kUmmaK; kk++) {
              // advance kUmmaK bytes along K inside the 128-byte swizzle atom: +32 B => +2 (16-B units)
              ptx::mma_i8_ss(d_tmem, a_desc + kk * (kUmmaK >> 4), b_desc + kk * (kUmmaK >> 4), idesc,
                             (kb | kk) != 0 ? 1u : 0u);
            }
            if (CSZ > 1) {
              ptx::tc_commit_mc(empty_bar(stage), mask_e);
            } else {
              ptx::tc_commit(empty_bar(stage));
            }
          }
          ptx::tc_commit(tfull_bar(buf));
        }
      }
    }
  } else {
    // ===================== epilogue: 8 warps, FP64 accumulators in registers =====================
    ptx::reg_alloc<224>();
    const uint32_t q = warp & 3u;            // TMEM lane quarter this warp may touch
    const uint32_t half = (warp - 4u) >> 2;  // which 64-column half of the tile
    const bool raw = p.single_a != 0;
    uint32_t pc = 0;
    for (uint32_t st = cluster_id; st < num_super; st += num_clusters) {
      uint32_t sm, sn;
      super_tile_coords(p, st, sm, sn);
      const uint32_t row = (sm * CM + cm) * BM + q * 32u + lane;
      const uint32_t col0 = (sn * CN + cn) * BN + half * 64u;
      double acc[64];
#pragma unroll
      for (int j = 0; j < 64; j++) acc[j] = 0.0;
      for (PairIter it(p); it.valid(); it.next(), pc++) {
        const uint32_t buf = pc % kAccBufs, bph = (pc / kAccBufs) & 1u;
        ptx::mbar_wait(tfull_bar(buf), bph);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + ((q * 32u) << 16) + buf * BN + half * 64u;
        uint32_t v[4][16];
#pragma unroll
        for (int c = 0; c < 4; c++) ptx::tmem_ld_x16(taddr + c * 16, v[c]);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(tempty_bar(buf));
        if (!raw) {
          const double scale = it.scale(p.bits);
#pragma unroll
          for (int c = 0; c < 4; c++)
#pragma unroll
            for (int j = 0; j < 16; j++)
              acc[c * 16 + j] = __fma_rn(__int2double_rn(static_cast<int32_t>(v[c][j])), scale, acc[c * 16 + j]);
        } else if (row < p.m) {
#pragma unroll
          for (int c = 0; c < 4; c++)
#pragma unroll
            for (int j = 0; j < 16; j++) {
              const uint32_t col = col0 + c * 16 + j;
              if (col < p.n) p.c_i32[static_cast<size_t>(col) * p.m + row] = static_cast<int32_t>(v[c][j]);
            }
        }
      }
      if (!raw && row < p.m) {
        // reference src/gemm.cu:124-148: x = acc / 2^44 * amax[mi] * bmax[ni]
        const double am = p.amax[row];
        double *crow = p.c + row;
#pragma unroll
        for (int j = 0; j < 64; j++) {
          const uint32_t col = col0 + j;
          if (col < p.n) {
            double x = __dmul_rn(acc[j], 0x1p-44);
            x = __dmul_rn(x, am);
            x = __dmul_rn(x, __ldg(p.bmax + col));
            double *dst = crow + static_cast<size_t>(col) * p.ldc;
            if (p.beta != 0) {
              *dst = __fma_rn(p.alpha, x, __dmul_rn(p.beta, *dst));
            } else {
              *dst = __dmul_rn(p.alpha, x);
            }
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  if (CSZ > 1) {
    ptx::cluster_arrive();
    ptx::cluster_wait();
  } else {
    __syncthreads();
  }
  if (warp == 2) ptx::tmem_dealloc<512>(tmem_base);
}

// =================================================================================================
// CTA-pair kernel (the default): tcgen05.mma.cta_group::2, UMMA M=256 x N=BN x K=32.
//
// Why: with one CTA per 128x128 tile every MMA-cycle moves 256 B through shared memory (TMA writes
// + tensor-core reads of (128+128) x 32 B per 64 cycles) against a 128 B/clk port, which pins the
// tensor pipe at 50 % (ncu profiles/r1_fused_1cta.txt).  A CTA pair shares B: each CTA stages its own
// 128 rows of A and only BN/2 rows of B, and reads the same -- (128 + BN/2) x 64 B per BN/2 cycles:
// 192 B/clk at BN=128, 149 B/clk at BN=192.  BN=192 is the widest tile whose FP64 accumulators
// (128 x 192 per CTA = 192 registers per epilogue thread) still fit the register file.
//
// Per CTA: 128 x BN outputs, FP64 accumulators in the registers of 8 epilogue warps, int32 products
// in kAccBufs TMEM buffers.  Leader CTA (cluster rank 0) issues all MMAs; both CTAs run a TMA
// producer (completing on the LEADER's full barrier) and an epilogue (arriving on the LEADER's
// tmem-empty barrier); tcgen05.commit multicasts "stage free" / "product ready" to both CTAs.
// =================================================================================================
constexpr uint32_t kSyncEvery = 16;      // k-steps per soft-lockstep interval of the pair kernel

// Position in one CTA's operand stream of the pair kernel: (tile, slice pair, k-block), in issue order.
// With a PM x PN cluster of CTA pairs each CTA fetches only its 1/PN share of the A rows and 1/PM
// share of the B rows it needs (the rest arrives by TMA multicast from its cluster neighbours).
template <uint32_t BN_, uint32_t PM, uint32_t PN>
struct KCursor {
  const FusedParams &p;
  uint32_t t, step, num_tiles, kb, rank, pm, pn;
  PairIter it;
  int row_a, row_b;
  __device__ KCursor(const FusedParams &p_, uint32_t first, uint32_t step_, uint32_t num_tiles_, uint32_t rank_,
                     uint32_t pm_, uint32_t pn_)
      : p(p_), t(first), step(step_), num_tiles(num_tiles_), kb(0), rank(rank_), pm(pm_), pn(pn_), it(p_) {
    set_rows();
  }
  __device__ void set_rows() {
    if (t >= num_tiles) return;
    uint32_t tm, tn;
    super_tile_coords(p, t, tm, tn);
    row_a = static_cast<int>((tm * PM + pm) * 2 * BM + rank * BM + pn * (BM / PN));
    row_b = static_cast<int>((tn * PN + pn) * BN_ + rank * (BN_ / 2) + pm * (BN_ / 2 / PM));
  }
  __device__ bool valid() const { return t < num_tiles; }
  __device__ int k0() const { return static_cast<int>(kb * BK); }
  __device__ int sa() const { return static_cast<int>(it.a_id() - 1); }
  __device__ int sb() const { return static_cast<int>(it.b_id() - 1); }
  __device__ void next() {
    if (++kb < p.k_blocks) return;
    kb = 0;
    it.next();
    if (it.valid()) return;
    it = PairIter(p);
    t += step;
    set_rows();
  }
  __device__ void prefetch(const CUtensorMap *ta, const CUtensorMap *tb) const {
    ptx::tma_prefetch_l2_3d(ta, k0(), row_a, sa());
    ptx::tma_prefetch_l2_3d(tb, k0(), row_b, sb());
  }
};

template <uint32_t BN_>
struct PairCfg {
  static constexpr uint32_t kStageBytes = (BM + BN_ / 2) * BK;       // per CTA
  static constexpr uint32_t kStages = (BN_ == 128) ? 8 : (BN_ == 192 ? 7 : 5);
  static constexpr uint32_t kAccBufs = (BN_ == 128) ? 4 : 2;
  static constexpr uint32_t kBufStride = (BN_ == 128) ? 128 : 256;   // TMEM columns between buffers
  static constexpr uint32_t kColsPerThread = BN_ / 2;                // epilogue: 2 column halves
  // FP64 accumulators: up to 96 columns per thread in registers (192 registers); a wider tile keeps
  // the rest in shared memory ([column][row] doubles, conflict-free for lane <-> row)
  static constexpr uint32_t kRegCols = kColsPerThread < 96 ? kColsPerThread : 96;
  static constexpr uint32_t kSpillCols = kColsPerThread - kRegCols;
  static constexpr uint32_t kSpillBytes = 2 * kSpillCols * BM * 8;
  static constexpr uint32_t kBarBytes = 8 * (2 * kStages + 2 * kAccBufs) + 16;
  static constexpr uint32_t kSmemBytes = kStages * kStageBytes + kSpillBytes + kBarBytes + 1024;
  static constexpr uint32_t kRegsOther = (BN_ == 128) ? 56 : 40;
  static constexpr uint32_t kRegsEpi = (BN_ == 128) ? 224 : 232;
};

template <uint32_t BN_, uint32_t PM, uint32_t PN>
__global__ void __launch_bounds__(kThreads, 1)
oz_gemm_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const FusedParams p) {
  using Cfg = PairCfg<BN_>;
  constexpr uint32_t CSZ = 2 * PM * PN;  // CTAs per cluster: PM x PN CTA pairs
  constexpr uint32_t kStagesP = Cfg::kStages, kBufs = Cfg::kAccBufs, kCols = Cfg::kColsPerThread;
  constexpr uint32_t kRegCols = Cfg::kRegCols, kSpillCols = Cfg::kSpillCols;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t spill_base = smem_base + kStagesP * Cfg::kStageBytes;
  const uint32_t bar_base = spill_base + Cfg::kSpillBytes;
  auto full_bar = [&](uint32_t s) { return bar_base + 8u * s; };
  auto empty_bar = [&](uint32_t s) { return bar_base + 8u * (kStagesP + s); };
  auto tfull_bar = [&](uint32_t b) { return bar_base + 8u * (2 * kStagesP + b); };
  auto tempty_bar = [&](uint32_t b) { return bar_base + 8u * (2 * kStagesP + kBufs + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStagesP + 2 * kBufs);
  volatile uint32_t *tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));

  // warp-uniform by construction (shfl from lane 0), so that the single-warp role loops below stay
  // on the uniform datapath: tcgen05.mma / TMA operands must be uniform registers, and values
  // produced under lane-divergent control flow force a per-instruction ELECT/R2UR waterfall.
  const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t crank = __shfl_sync(0xffffffffu, ptx::cluster_ctarank(), 0);
  const uint32_t rank = crank & 1u;      // within the CTA pair: 0 = leader (issues the MMAs)
  const uint32_t pidx = crank >> 1;      // pair index in the cluster
  const uint32_t pm = pidx % PM, pn = pidx / PM;
  const uint32_t lead = crank & ~1u;     // cluster rank of this pair's leader
  const uint32_t pair_id = blockIdx.x / CSZ;       // cluster index (one cluster tile at a time)
  const uint32_t num_pairs = gridDim.x / CSZ;
  const uint32_t num_tiles = p.super_m * p.super_n;  // cluster tiles: (PM*256) rows x (PN*BN) columns

  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < kStagesP; s++) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), PM + PN - 1);  // one release per pair that reads what this CTA stages
    }
    for (uint32_t b = 0; b < kBufs; b++) {
      ptx::mbar_init(tfull_bar(b), 1);
      ptx::mbar_init(tempty_bar(b), 2 * kEpiWarps);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_b);
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
      // ===================== TMA producer (both CTAs; whole warp loops, one lane issues) ==========
      // Optional (OZIMMU_B200_PREFETCH=n, default off): a second cursor runs n k-blocks ahead of the ring
      // and pulls the boxes into L2 (cp.async.bulk.prefetch).  It hides HBM latency, but the kernel is
      // bound by L2->SM delivery and the extra L2 lookups cost more than they save: -5 % at 8192^3
      // (profiles/r1_sweep_prefetch_lockstep_256.txt).
      const bool issuer = ptx::elect_one();
      uint32_t stage = 0, ph = 0;
      KCursor<BN_, PM, PN> ahead(p, pair_id, num_pairs, num_tiles, rank, pm, pn);
      // TMA multicast: this CTA's share of the A rows goes to the same-rank CTA of every pair in its
      // cluster row, its share of the B rows to the same-rank CTA of every pair in its cluster column
      uint16_t mask_a = 0, mask_b = 0;
      for (uint32_t j = 0; j < PN; j++) mask_a |= static_cast<uint16_t>(1u << (2 * (pm + PM * j) + rank));
      for (uint32_t i = 0; i < PM; i++) mask_b |= static_cast<uint16_t>(1u << (2 * (i + PM * pn) + rank));
      for (uint32_t i = 0; i < p.prefetch_ahead && ahead.valid(); i++, ahead.next()) {
        if (issuer) ahead.prefetch(&tmap_a, &tmap_b);
      }
      // Soft lockstep: the CTA pairs of one wave share slice panels through L2 only while they stay
      // within a few k-steps of each other; left alone they drift apart by whole products and every
      // panel is re-fetched from HBM (ncu: 103 GB of DRAM reads for a 1.2 GB working set at 8192^3).
      // Leaders count arrivals per kSyncEvery-step interval and do not start interval j before all
      // pairs have started interval j - window.  It is a performance hint only: the wait is bounded
      // and abandoned for the rest of the kernel on the first timeout.
      const uint32_t steps_per_tile = (p.single_a != 0 ? 1u : p.num_split * (p.num_split + 1) / 2) * p.k_blocks;
      const uint32_t tiles_max = (num_tiles + num_pairs - 1) / num_pairs;
      const uint32_t pairs_last = num_tiles - (tiles_max - 1) * num_pairs;  // pairs that own tiles_max tiles
      bool lockstep = p.sync_ctr != nullptr && crank == 0;
      uint32_t g = 0;
      for (KCursor<BN_, PM, PN> cur(p, pair_id, num_pairs, num_tiles, rank, pm, pn); cur.valid(); cur.next(), g++) {
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
                const uint32_t want = (static_cast<uint64_t>(jw) * kSyncEvery < static_cast<uint64_t>(tiles_max - 1) * steps_per_tile)
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
        ptx::mbar_wait_relaxed(empty_bar(stage), ph ^ 1u, p.idle_sleep_ns);
        const uint32_t leader_full = ptx::mapa(full_bar(stage), lead);
        const uint32_t a_dst = smem_base + stage * Cfg::kStageBytes + pn * (BM / PN) * BK;
        const uint32_t b_dst = smem_base + stage * Cfg::kStageBytes + BM * BK + pm * (BN_ / 2 / PM) * BK;
        if (issuer) {
          if (rank == 0) ptx::mbar_expect_tx(full_bar(stage), 2 * Cfg::kStageBytes);
          if (PN > 1) ptx::tma_load_3d_2sm_mc(a_dst, &tmap_a, leader_full, cur.k0(), cur.row_a, cur.sa(), mask_a);
          else ptx::tma_load_3d_2sm(a_dst, &tmap_a, leader_full, cur.k0(), cur.row_a, cur.sa());
          if (PM > 1) ptx::tma_load_3d_2sm_mc(b_dst, &tmap_b, leader_full, cur.k0(), cur.row_b, cur.sb(), mask_b);
          else ptx::tma_load_3d_2sm(b_dst, &tmap_b, leader_full, cur.k0(), cur.row_b, cur.sb());
          if (ahead.valid()) ahead.prefetch(&tmap_a, &tmap_b);
        }
        if (ahead.valid()) ahead.next();
        if (++stage == kStagesP) { stage = 0; ph ^= 1u; }
      }
    } else if (warp == 1 && rank == 0) {
      // ===================== MMA issuer (leader CTA; whole warp loops, one lane issues) ============
      constexpr uint32_t idesc = ptx::make_i8_idesc(2 * BM, BN_);
      const bool issuer = ptx::elect_one();
      // "stage consumed" goes to every CTA that stages data for this pair: both CTAs of all pairs in
      // this pair's cluster row and column; "product ready" to the two CTAs of this pair
      uint16_t mask_e = 0;
      for (uint32_t j = 0; j < PN; j++) mask_e |= static_cast<uint16_t>(3u << (2 * (pm + PM * j)));
      for (uint32_t i = 0; i < PM; i++) mask_e |= static_cast<uint16_t>(3u << (2 * (i + PM * pn)));
      const uint16_t mask_t = static_cast<uint16_t>(3u << (2 * pidx));
      uint32_t stage = 0, ph = 0, buf = 0, bph = 0;
      for (uint32_t t = pair_id; t < num_tiles; t += num_pairs) {
        for (PairIter it(p); it.valid(); it.next()) {
          ptx::mbar_wait_cluster(tempty_bar(buf), bph ^ 1u);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * Cfg::kBufStride;
          for (uint32_t kb = 0; kb < p.k_blocks; kb++) {
            ptx::mbar_wait_cluster(full_bar(stage), ph);
            ptx::tc_fence_after();
            const uint32_t a_smem = smem_base + stage * Cfg::kStageBytes;
            const uint64_t a_desc = ptx::make_sw128_kmajor_desc(a_smem);
            const uint64_t b_desc = ptx::make_sw128_kmajor_desc(a_smem + BM * BK);
            if (issuer) {
#pragma unroll
              for (uint32_t kk = 0; kk < BK / kUmmaK; kk++)
                ptx::mma_i8_ss_2sm(d_tmem, a_desc + kk * (kUmmaK >> 4), b_desc + kk * (kUmmaK >> 4), idesc,
                                   (kb | kk) != 0 ? 1u : 0u);
              ptx::tc_commit_2sm_mc(empty_bar(stage), mask_e);
            }
            __syncwarp();
            if (++stage == kStagesP) { stage = 0; ph ^= 1u; }
          }
          if (issuer) ptx::tc_commit_2sm_mc(tfull_bar(buf), mask_t);
          __syncwarp();
          if (++buf == kBufs) { buf = 0; bph ^= 1u; }
        }
      }
    }
  } else {
    // ===================== epilogue (both CTAs): 8 warps, FP64 accumulators in registers ==========
    ptx::reg_alloc<Cfg::kRegsEpi>();
    const uint32_t q = warp & 3u;            // TMEM lane quarter this warp may touch
    const uint32_t half = (warp - 4u) >> 2;  // which column half of the tile
    const bool raw = p.single_a != 0;
    uint32_t pc = 0;
    for (uint32_t t = pair_id; t < num_tiles; t += num_pairs) {
      uint32_t tm, tn;
      super_tile_coords(p, t, tm, tn);
      const uint32_t row = (tm * PM + pm) * 2 * BM + rank * BM + q * 32u + lane;
      const uint32_t col0 = (tn * PN + pn) * BN_ + half * kCols;
      double acc[kRegCols];
#pragma unroll
      for (uint32_t j = 0; j < kRegCols; j++) acc[j] = 0.0;
      // this thread's spill accumulators: columns [half*kSpillCols, +kSpillCols) of the [col][row] array
      double *spill = reinterpret_cast<double *>(smem_raw + (spill_base - ptx::smem_u32(smem_raw))) +
                      static_cast<size_t>(half * kSpillCols) * BM + (q * 32u + lane);
      bool first = true;
      for (PairIter it(p); it.valid(); it.next(), pc++) {
        const uint32_t buf = pc % kBufs, bph = (pc / kBufs) & 1u;
        ptx::mbar_wait_relaxed(tfull_bar(buf), bph, p.idle_sleep_ns);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + ((q * 32u) << 16) + buf * Cfg::kBufStride + half * kCols;
        const double scale = it.scale(p.bits);
#pragma unroll
        for (uint32_t c = 0; c < kCols / 16; c++) {
          uint32_t v[16];
          ptx::tmem_ld_x16(taddr + c * 16, v);
          ptx::tmem_ld_wait();
          if (!raw) {
            // (double)p without I2F.F64 (a quarter-rate conversion, 15/clk/SM measured, that would
            // make the epilogue as slow as the MMAs): 2^52 + 2^31 + p is the bit pattern
            // {0x43300000, p ^ 0x80000000}; subtracting 2^52 + 2^31 is exact.
#pragma unroll
            for (uint32_t g = 0; g < 16; g += 8) {
              double d[8];
#pragma unroll
              for (uint32_t j = 0; j < 8; j++)
                d[j] = __dadd_rn(__hiloint2double(0x43300000, static_cast<int>(v[g + j] ^ 0x80000000u)),
                                 -4503601774854144.0);
              if (c * 16 < kRegCols) {
#pragma unroll
                for (uint32_t j = 0; j < 8; j++)
                  acc[(c * 16 + g + j) % kRegCols] = __fma_rn(d[j], scale, acc[(c * 16 + g + j) % kRegCols]);
              } else {
#pragma unroll
                for (uint32_t j = 0; j < 8; j++) {
                  double *sp = spill + static_cast<size_t>(c * 16 + g + j - kRegCols) * BM;
                  *sp = __fma_rn(d[j], scale, first ? 0.0 : *sp);
                }
              }
            }
          } else if (row < p.m) {
#pragma unroll
            for (uint32_t j = 0; j < 16; j++) {
              const uint32_t col = col0 + c * 16 + j;
              if (col < p.n) p.c_i32[static_cast<size_t>(col) * p.m + row] = static_cast<int32_t>(v[j]);
            }
          }
        }
        first = false;
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (rank == 0) ptx::mbar_arrive(tempty_bar(buf));
          else ptx::mbar_arrive_remote(ptx::mapa(tempty_bar(buf), lead));
        }
      }
      if (!raw && row < p.m) {
        // reference src/gemm.cu:124-148: x = acc / 2^44 * amax[mi] * bmax[ni]
        const double am = p.amax[row];
        double *crow = p.c + row;
#pragma unroll
        for (uint32_t j = 0; j < kCols; j++) {
          const uint32_t col = col0 + j;
          if (col < p.n) {
            const double a_j = (j < kRegCols) ? acc[j % kRegCols] : spill[static_cast<size_t>(j - kRegCols) * BM];
            double x = __dmul_rn(a_j, 0x1p-44);
            x = __dmul_rn(x, am);
            x = __dmul_rn(x, __ldg(p.bmax + col));
            if (p.cplx) {
              double2 *dst = reinterpret_cast<double2 *>(p.c) + static_cast<size_t>(col) * p.ldc + row;
              double2 y = make_double2(0.0, 0.0);
              if (p.cplx_init) {
                if (p.beta != 0 || p.beta_im != 0) {
                  // init_c_complex_kernel<false> as compiled (reference src/gemm.cu:214-222, incl. its use
                  // of the already-updated real part): t = y.y*b.y; y.x = fma(y.x, b.x, -t);
                  // t = y.x*b.y; y.y = fma(y.y, b.x, t)
                  y = *dst;
                  const double yx = __fma_rn(y.x, p.beta, -__dmul_rn(y.y, p.beta_im));
                  y.y = __fma_rn(y.y, p.beta, __dmul_rn(yx, p.beta_im));
                  y.x = yx;
                }
              } else {
                y = *dst;
              }
              y.x = __fma_rn(x, p.alpha, y.x);      // axy_complex_kernel (reference src/gemm.cu:160-186)
              y.y = __fma_rn(x, p.alpha_im, y.y);
              *dst = y;
            } else {
              double *dst = crow + static_cast<size_t>(col) * p.ldc;
              if (p.beta != 0) {
                *dst = __fma_rn(p.alpha, x, __dmul_rn(p.beta, *dst));
              } else {
                *dst = __dmul_rn(p.alpha, x);
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

// k == 0: every product is empty, C = beta * C (beta == 0: C is not read, reference src/gemm.cu:143-147)
__global__ void __launch_bounds__(256)
oz_scale_c_kernel(double *__restrict__ c, const size_t ldc, const uint32_t m, const uint32_t n, const double beta) {
  const uint32_t r = blockIdx.x * 256 + threadIdx.x, col = blockIdx.y;
  if (r >= m || col >= n) return;
  double *p = c + static_cast<size_t>(col) * ldc + r;
  *p = (beta != 0) ? __dmul_rn(beta, *p) : 0.0;
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

using EncodeTiledFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                   const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                   const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// slices: [num_split][rows][pitch] int8 -> 3-D map (k, row, slice), box (128, box_rows, 1), SW128
int make_slice_tmap(CUtensorMap *map, const int8_t *base, size_t rows, size_t pitch,
                    unsigned num_split, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return static_cast<int>(cudaErrorNotSupported);
  cuuint64_t dims[3] = {pitch, rows, num_split};
  cuuint64_t strides[2] = {pitch, static_cast<cuuint64_t>(rows) * pitch};
  cuuint32_t box[3] = {BK, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<int8_t *>(base), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : static_cast<int>(cudaErrorInvalidValue);
}

// 0 = default (CTA-pair kernel, BN chosen per problem); 128/192/256 = CTA-pair kernel with that BN; CM*10+CN = the
// single-CTA kernel with a CM x CN multicast cluster (test/tuning hook, see ozk_set_cluster_shape)
int g_cluster_override = 0;

template <uint32_t CM, uint32_t CN>
int launch_fused(const FusedParams &p0, const int8_t *a_slices, const int8_t *b_slices, size_t pitch,
                 cudaStream_t stream) {
  FusedParams p = p0;
  CUtensorMap ta, tb;
  int rc = make_slice_tmap(&ta, a_slices, p.m, pitch, p.num_split, BM / CN);
  if (rc) return rc;
  rc = make_slice_tmap(&tb, b_slices, p.n, pitch, p.num_split, BN / CM);
  if (rc) return rc;
  const uint32_t tiles_m = ceil_div_u32(p.m, BM), tiles_n = ceil_div_u32(p.n, BN);
  p.super_m = ceil_div_u32(tiles_m, CM);
  p.super_n = ceil_div_u32(tiles_n, CN);
  p.group_m = (12 / CM) > 0 ? 12 / CM : 1;

  auto kern = oz_gemm_fused_kernel<CM, CN>;
  int dev = 0, sms = 0;
  OZ_CUDA_TRY(cudaGetDevice(&dev));
  OZ_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (dev < 0 || dev >= kMaxDevices) return static_cast<int>(cudaErrorInvalidDevice);

  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = stream;
  uint32_t max_clusters = static_cast<uint32_t>(sms) / (CM * CN);
  if (CM * CN > 1) {
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CM * CN;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  static PerDeviceOnce once;
  {
    std::lock_guard<std::mutex> lock(once.mu);
    if (!once.done[dev]) {
      OZ_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
      once.value[dev] = static_cast<int>(max_clusters);
      if (CM * CN > 1) {
        cfg.gridDim = dim3(static_cast<unsigned>(sms) / (CM * CN) * (CM * CN));
        int nc = 0;
        if (cudaOccupancyMaxActiveClusters(&nc, kern, &cfg) == cudaSuccess && nc > 0) once.value[dev] = nc;
        cudaGetLastError();
      }
      once.done[dev] = true;
    }
    max_clusters = static_cast<uint32_t>(once.value[dev]);
  }
  const uint32_t num_super = p.super_m * p.super_n;
  const uint32_t clusters = num_super < max_clusters ? num_super : max_clusters;
  if (clusters == 0) return 0;
  cfg.gridDim = dim3(clusters * CM * CN);
  OZ_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, ta, tb, p));
  count_launch(1);
  return 0;
}


// ---- pair-kernel tuning knobs (read once; OZIMMU_B200_PREFETCH / OZIMMU_B200_LOCKSTEP override) ----
struct PairTuning {
  uint32_t prefetch_ahead;  // k-blocks of L2 prefetch lead (0 = off)
  uint32_t sync_window;     // soft-lockstep window in kSyncEvery-step intervals (0 = off)
  uint32_t idle_sleep_ns;   // OZIMMU_B200_IDLE_SLEEP: nanosleep of idle roles between barrier probes
};
const PairTuning &pair_tuning() {
  static const PairTuning t = [] {
    PairTuning v{0, 2, 0};
    if (const char *e = std::getenv("OZIMMU_B200_IDLE_SLEEP")) v.idle_sleep_ns = static_cast<uint32_t>(std::atoi(e));
    if (const char *e = std::getenv("OZIMMU_B200_PREFETCH")) v.prefetch_ahead = static_cast<uint32_t>(std::atoi(e));
    if (const char *e = std::getenv("OZIMMU_B200_LOCKSTEP")) v.sync_window = static_cast<uint32_t>(std::atoi(e));
    if (v.prefetch_ahead > 256) v.prefetch_ahead = 256;
    return v;
  }();
  return t;
}

constexpr uint32_t kSyncCounters = 1u << 16;  // per buffer: 64 Ki intervals = 1 Mi k-steps per CTA pair
constexpr int kSyncBuffers = 8;               // launches that may be in flight at once without sharing
uint32_t *next_sync_buffer() {
  static std::mutex mu;
  static uint32_t *pool[kSyncBuffers] = {};
  static int pool_dev[kSyncBuffers] = {};
  static int next = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  const int i = next;
  next = (next + 1) % kSyncBuffers;
  if (pool[i] != nullptr && pool_dev[i] != dev) {
    pool[i] = nullptr;  // another device's buffer: leak-free enough for a per-process pool of 8 x 256 KB
  }
  if (pool[i] == nullptr) {
    if (cudaMalloc(&pool[i], kSyncCounters * sizeof(uint32_t)) != cudaSuccess) {
      pool[i] = nullptr;
      cudaGetLastError();
      return nullptr;
    }
    pool_dev[i] = dev;
  }
  return pool[i];
}

template <uint32_t BN_, uint32_t PM, uint32_t PN>
int launch_pair(const FusedParams &p0, const int8_t *a_slices, const int8_t *b_slices, size_t pitch,
                cudaStream_t stream) {
  using Cfg = PairCfg<BN_>;
  constexpr uint32_t CSZ = 2 * PM * PN;
  FusedParams p = p0;
  CUtensorMap ta, tb;
  int rc = make_slice_tmap(&ta, a_slices, p.m, pitch, p.num_split, BM / PN);
  if (rc) return rc;
  rc = make_slice_tmap(&tb, b_slices, p.n, pitch, p.num_split, BN_ / 2 / PM);
  if (rc) return rc;
  p.super_m = ceil_div_u32(p.m, 2 * BM * PM);
  p.super_n = ceil_div_u32(p.n, BN_ * PN);
  p.group_m = 8 / PM;
  const PairTuning &tune = pair_tuning();
  p.prefetch_ahead = tune.prefetch_ahead;
  p.sync_window = tune.sync_window;
  p.idle_sleep_ns = tune.idle_sleep_ns;

  auto kern = oz_gemm_pair_kernel<BN_, PM, PN>;
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
  attr[0].val.clusterDim.x = CSZ;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  static PerDeviceOnce once;
  int cached_max = 0;
  {
    std::lock_guard<std::mutex> lock(once.mu);
    if (!once.done[dev]) {
      OZ_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
      cfg.gridDim = dim3(static_cast<unsigned>(sms) / CSZ * CSZ);
      int nc = 0;
      if (cudaOccupancyMaxActiveClusters(&nc, kern, &cfg) == cudaSuccess && nc > 0) once.value[dev] = nc;
      else once.value[dev] = sms / static_cast<int>(CSZ);
      cudaGetLastError();
      once.done[dev] = true;
      if (std::getenv("OZIMMU_B200_DEBUG"))
        std::fprintf(stderr, "[ozimmu_b200] pair kernel BN=%u cluster %ux%u pairs: %d clusters resident (%d of %d SMs)\n", BN_,
                     PM, PN, once.value[dev], once.value[dev] * static_cast<int>(CSZ), sms);
    }
    cached_max = once.value[dev];
  }
  const uint32_t num_tiles = p.super_m * p.super_n;
  const uint32_t pairs = num_tiles < static_cast<uint32_t>(cached_max) ? num_tiles : static_cast<uint32_t>(cached_max);
  if (pairs == 0) return 0;
  cfg.gridDim = dim3(pairs * CSZ);
  // lockstep counters only pay off when several rounds of tiles stream through L2 (pairs cannot drift apart
  // within a single round, and the polling costs ~7 % there)
  p.sync_ctr = nullptr;
  if (tune.sync_window > 0 && pairs > 1 && num_tiles > pairs && p.single_a == 0) {
    const uint64_t steps = static_cast<uint64_t>(ceil_div_u32(num_tiles, pairs)) * (p.num_split * (p.num_split + 1) / 2) * p.k_blocks;
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
  OZ_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, ta, tb, p));
  count_launch(1);
  return 0;
}

int dispatch_fused(const FusedParams &p, const int8_t *a_slices, const int8_t *b_slices, size_t pitch,
                   cudaStream_t stream) {
  int shape = g_cluster_override;
  if (p.cplx && shape != 128 && shape != 192 && shape != 256) shape = 0;  // complex epilogue: CTA-pair kernel only
  if (shape == 0) {
    // The kernel is bound by L2->SM delivery, so a launch costs about
    // rounds x bytes-per-k-block-per-SM = ceil(tiles / resident pairs) x (128 + BN/2).  BN=256 delivers
    // the fewest bytes per MAC; narrower tiles win when they fill more SMs or avoid a ragged last round.
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint64_t pairs = static_cast<uint64_t>(sms > 1 ? sms / 2 : 1);
    uint64_t best_cost = ~0ull;
    for (int bn : {256, 192, 128}) {
      const uint64_t tiles = static_cast<uint64_t>(ceil_div_u32(p.m, 2 * BM)) * ceil_div_u32(p.n, bn);
      const uint64_t cost = ((tiles + pairs - 1) / pairs) * (128 + bn / 2);
      if (cost < best_cost) {
        best_cost = cost;
        shape = bn;
      }
    }
  }
  switch (shape) {
    case 192: return launch_pair<192, 1, 1>(p, a_slices, b_slices, pitch, stream);
    case 128: return launch_pair<128, 1, 1>(p, a_slices, b_slices, pitch, stream);
    case 256: return launch_pair<256, 1, 1>(p, a_slices, b_slices, pitch, stream);
    case 1120: return launch_pair<192, 2, 1>(p, a_slices, b_slices, pitch, stream);
    case 1210: return launch_pair<192, 1, 2>(p, a_slices, b_slices, pitch, stream);
    case 1220: return launch_pair<192, 2, 2>(p, a_slices, b_slices, pitch, stream);
    case 11: return launch_fused<1, 1>(p, a_slices, b_slices, pitch, stream);
    case 21: return launch_fused<2, 1>(p, a_slices, b_slices, pitch, stream);
    case 12: return launch_fused<1, 2>(p, a_slices, b_slices, pitch, stream);
    case 22: return launch_fused<2, 2>(p, a_slices, b_slices, pitch, stream);
    default: return static_cast<int>(cudaErrorInvalidValue);
  }
}

bool valid_common(size_t m, size_t n, size_t k, size_t pitch, unsigned num_split, unsigned bits) {
  return m > 0 && n > 0 && k > 0 && m < (1ull << 31) && n < (1ull << 31) && pitch % 16 == 0 &&
         pitch >= k && pitch < (1ull << 31) && num_split >= 1 && num_split <= 18 && bits >= 1 && bits <= 7;
}

}  // namespace
}  // namespace oz

// Test/tuning hook: force the cluster shape of the fused kernel (0 = default heuristic).
extern "C" int ozk_set_cluster_shape(int cm, int cn) {
  if (cm == 0 && (cn == 128 || cn == 192 || cn == 256)) oz::g_cluster_override = cn;  // CTA-pair kernel, BN = cn
  else if (cm == 100 && (cn == 21 || cn == 12 || cn == 22))              // BN=192, PM x PN pairs multicast
    oz::g_cluster_override = 1000 + cn * 10;
  else oz::g_cluster_override = (cm <= 0 || cn <= 0) ? 0 : cm * 10 + cn;
  return 0;
}

extern "C" int ozk_gemm_i8_fused(size_t m, size_t n, size_t k, const int8_t *a_slices,
                                 const int8_t *b_slices, size_t pitch, const double *amax,
                                 const double *bmax, unsigned num_split, unsigned bits_per_int8,
                                 double alpha, double beta, double *c, size_t ldc, void *stream) {
  if (m == 0 || n == 0) return 0;
  if (!oz::valid_common(m, n, k, pitch, num_split, bits_per_int8) || ldc < m)
    return static_cast<int>(cudaErrorInvalidValue);
  oz::FusedParams p{};
  p.m = static_cast<uint32_t>(m);
  p.n = static_cast<uint32_t>(n);
  p.k_blocks = oz::ceil_div_u32(static_cast<uint32_t>(pitch), oz::BK);
  p.num_split = num_split;
  p.bits = static_cast<int32_t>(bits_per_int8);
  p.alpha = alpha;
  p.beta = beta;
  p.c = c;
  p.ldc = ldc;
  p.amax = amax;
  p.bmax = bmax;
  return oz::dispatch_fused(p, a_slices, b_slices, pitch, static_cast<cudaStream_t>(stream));
}

extern "C" int ozk_gemm_i8_fused_complex(size_t m, size_t n, size_t k, const int8_t *a_slices,
                                         const int8_t *b_slices, size_t pitch, const double *amax,
                                         const double *bmax, unsigned num_split, unsigned bits_per_int8,
                                         double coef_re, double coef_im, int apply_beta, double beta_re,
                                         double beta_im, void *c, size_t ldc, void *stream) {
  if (m == 0 || n == 0) return 0;
  if (!oz::valid_common(m, n, k, pitch, num_split, bits_per_int8) || ldc < m)
    return static_cast<int>(cudaErrorInvalidValue);
  oz::FusedParams p{};
  p.m = static_cast<uint32_t>(m);
  p.n = static_cast<uint32_t>(n);
  p.k_blocks = oz::ceil_div_u32(static_cast<uint32_t>(pitch), oz::BK);
  p.num_split = num_split;
  p.bits = static_cast<int32_t>(bits_per_int8);
  p.alpha = coef_re;
  p.alpha_im = coef_im;
  p.beta = beta_re;
  p.beta_im = beta_im;
  p.cplx = 1;
  p.cplx_init = apply_beta ? 1u : 0u;
  p.c = static_cast<double *>(c);
  p.ldc = ldc;
  p.amax = amax;
  p.bmax = bmax;
  return oz::dispatch_fused(p, a_slices, b_slices, pitch, static_cast<cudaStream_t>(stream));
}

extern "C" int ozk_scale_c(size_t m, size_t n, double beta, double *c, size_t ldc, void *stream) {
  if (m == 0 || n == 0) return 0;
  if (ldc < m || m >= (1ull << 31) || n >= 65536ull * 32768ull) return static_cast<int>(cudaErrorInvalidValue);
  for (size_t j0 = 0; j0 < n; j0 += 65535) {
    const unsigned nj = static_cast<unsigned>(n - j0 < 65535 ? n - j0 : 65535);
    dim3 grid(static_cast<unsigned>((m + 255) / 256), nj);
    oz::oz_scale_c_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(c + j0 * ldc, ldc, static_cast<uint32_t>(m),
                                                                              nj, beta);
    oz::count_launch(1);
  }
  return static_cast<int>(cudaGetLastError());
}

extern "C" int ozk_gemm_i8_pair(size_t m, size_t n, size_t k, const int8_t *a_slices,
                                const int8_t *b_slices, size_t pitch, unsigned num_split,
                                unsigned a_id, unsigned b_id, int32_t *c_i32, void *stream) {
  if (m == 0 || n == 0) return 0;
  if (!oz::valid_common(m, n, k, pitch, num_split, 7) || a_id < 1 || b_id < 1 || a_id > num_split ||
      b_id > num_split)
    return static_cast<int>(cudaErrorInvalidValue);
  oz::FusedParams p{};
  p.m = static_cast<uint32_t>(m);
  p.n = static_cast<uint32_t>(n);
  p.k_blocks = oz::ceil_div_u32(static_cast<uint32_t>(pitch), oz::BK);
  p.num_split = num_split;
  p.bits = 7;
  p.single_a = a_id;
  p.single_b = b_id;
  p.c_i32 = c_i32;
  return oz::dispatch_fused(p, a_slices, b_slices, pitch, static_cast<cudaStream_t>(stream));
}
