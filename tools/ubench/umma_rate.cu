// Microbenchmark (sm_100a): what bounds the int8 CTA-pair MMA -- the shared-memory byte model of DESIGN 3.2, and would
// the M-side operand through tensor memory (TS mode), fed from registers, lift the bound?
// One CTA pair per SM pair; per "k-step" (128 bytes of K) the leader issues 4 x tcgen05.mma.cta_group::2.kind::i8
// (M = 256, N = kN, K = 32) into one accumulator.  Data is garbage (never read back); only the rate matters.
//   mode 0: SS (A and B in SMEM), nothing else running                     -> SMEM reads only
//   mode 1: SS + the producer ring of the product kernel (A 16 KB + B kN/2 x 128 B per k-step, full/empty barriers)
//   mode 2: TS (A in TMEM, static) + the ring carrying B only
//   mode 3: mode 2 + four warps that stream 16 KB of A per k-step from global memory (ld.global.nc.L1::no_allocate)
//           and write it to the A columns of TMEM with tcgen05.st
//   mode 4: mode 3 with plain ld.global (L1-allocating)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -DKN=256 -o tools/ubench/umma_rate tools/ubench/umma_rate.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "../../ozimmu_b200/csrc/ptx.cuh"

using namespace oz;

#ifndef KN
#define KN 256
#endif
#ifndef KM
#define KM 256   // UMMA M of the pair: 256 (128 rows of A per CTA) or 128 (64 rows per CTA; SS modes only)
#endif
constexpr uint32_t kN = KN, kStages = 5, kThreads = 224;   // warps: 0 producer, 1 MMA, 2-5 A stream, 6 relay
constexpr uint32_t kM = KM;
constexpr uint32_t kABytes = (kM / 2) * 128, kBBytes = (kN / 2) * 128;
constexpr uint32_t kStageBytes = kABytes + kBBytes;
constexpr uint32_t kSmem = kStages * kStageBytes + 1024 + 256;
constexpr uint32_t kAColsTmem = 512 - 64;   // two A buffers of 32 columns at the top of TMEM

__device__ __forceinline__ void mma_i8_ts_2sm(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_st_x32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
        "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
        "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
}
__device__ __forceinline__ uint4 ldg_na(const uint4 *p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
k(const uint8_t *__restrict__ g, uint32_t chunks, uint32_t steps, int mode, unsigned long long *out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + kStages * kStageBytes;
  auto full_bar = [&](uint32_t s) { return bar_base + 8u * s; };
  auto pfull_bar = [&](uint32_t s) { return bar_base + 8u * (kStages + s); };
  auto empty_bar = [&](uint32_t s) { return bar_base + 8u * (2 * kStages + s); };
  const uint32_t done_bar = bar_base + 8u * 3 * kStages, tmem_slot = done_bar + 8;
  volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));
  const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const uint32_t rank = __shfl_sync(0xffffffffu, ptx::cluster_ctarank(), 0) & 1u;
  const bool ring = mode >= 1, ts = mode >= 2, ldg = mode >= 3;
  const uint32_t stage_bytes = ts ? kBBytes : kStageBytes;
  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < kStages; s++) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(pfull_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    ptx::mbar_init(done_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc_2sm<512>(tmem_slot);
  ptx::tc_fence_before();
  ptx::cluster_arrive();
  ptx::cluster_wait();
  ptx::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot_ptr, 0);
  const long long t0 = clock64();
  if (warp == 0) {
    if (ring) {
      const bool issuer = ptx::elect_one();
      uint32_t stage = 0, ph = 0;
      for (uint32_t i = 0; i < steps; i++) {
        ptx::mbar_wait(empty_bar(stage), ph ^ 1u);
        const uint32_t dst = smem_base + stage * kStageBytes;
        if (issuer) {
          // chunks shared by the 8-9 pairs of a "tile row / column", as in the product kernel
          const uint8_t *sa = g + static_cast<size_t>(((blockIdx.x >> 4) * 977u + i * 2u) % chunks) * 16384u;
          const uint8_t *sb = g + static_cast<size_t>(((blockIdx.x & 15u) * 331u + i * 2u + 1u) % chunks) * 16384u;
          ptx::mbar_expect_tx(full_bar(stage), stage_bytes);
          if (!ts) ptx::bulk_load(dst, sa, kABytes, full_bar(stage));
          ptx::bulk_load(dst + kABytes, sb, kBBytes, full_bar(stage));
        }
        if (++stage == kStages) { stage = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      constexpr uint32_t idesc = ptx::make_i8_idesc(kM, kN);
      const bool issuer = ptx::elect_one();
      uint32_t stage = 0, ph = 0;
      for (uint32_t i = 0; i < steps; i++) {
        if (ring) {
          ptx::mbar_wait(full_bar(stage), ph);
          ptx::mbar_wait_cluster(pfull_bar(stage), ph);
          ptx::tc_fence_after();
        }
        const uint32_t a_smem = smem_base + stage * kStageBytes;
        const uint64_t a_desc = ptx::make_sw128_kmajor_desc(a_smem);
        const uint64_t b_desc = ptx::make_sw128_kmajor_desc(a_smem + kABytes);
        const uint32_t a_tmem = tmem_base + kAColsTmem + (i & 1u) * 32u;
        if (issuer) {
#pragma unroll
          for (uint32_t kk = 0; kk < 4; kk++) {
            if (ts) mma_i8_ts_2sm(tmem_base, a_tmem + kk * 8u, b_desc + kk * 2u, idesc, 1u);
            else ptx::mma_i8_ss_2sm(tmem_base, a_desc + kk * 2u, b_desc + kk * 2u, idesc, 1u);
          }
          if (ring) ptx::tc_commit_2sm_mc(empty_bar(stage), 0x3);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; ph ^= 1u; }
      }
      if (issuer) ptx::tc_commit_2sm_mc(done_bar, 0x3);
      __syncwarp();
    }
  } else if (warp == 6) {
    // relay of the non-leader CTA ("my stage landed" -> the leader's pfull barrier), as in the product kernel
    if (rank == 1 && ring) {
      uint32_t stage = 0, ph = 0;
      for (uint32_t i = 0; i < steps; i++) {
        ptx::mbar_wait(full_bar(stage), ph);
        if (lane == 0) ptx::mbar_arrive_remote_relaxed(ptx::mapa(pfull_bar(stage), 0));
        __syncwarp();
        if (++stage == kStages) { stage = 0; ph ^= 1u; }
      }
    }
  } else if (ldg) {
    // warps 2..5: lane quarter (warp & 3); every thread owns one row of A: 128 B per k-step
    const uint32_t q = warp & 3u;
    // chunk-major A tile ([8 chunks][128 rows][16 B]): a warp's LDG.128 reads 512 contiguous bytes
    const uint4 *src = reinterpret_cast<const uint4 *>(g) + (static_cast<size_t>(blockIdx.x >> 4) * 1777u % chunks) * 1024u +
                       (q * 32u + lane);
    uint32_t keep = 0;
    auto chunk = [&](uint32_t i) { return src + static_cast<size_t>((i * 7u) % 512u) * 1024u; };   // a 16 KB chunk per k-step
    auto load8 = [&](const uint4 *p, uint4 (&x)[8]) {
#pragma unroll
      for (uint32_t c = 0; c < 8; c++) x[c] = (mode == 3) ? ldg_na(p + c * 128u) : __ldg(p + c * 128u);
    };
    // three k-steps of loads in flight per thread (the L2 round trip is longer than a k-step)
    uint4 x0[8], x1[8], x2[8];
    load8(chunk(0), x0);
    load8(chunk(1), x1);
    for (uint32_t i = 0; i < steps; i += 3) {
      auto put = [&](uint32_t step, const uint4 (&x)[8]) {
        uint32_t v[32];
#pragma unroll
        for (uint32_t c = 0; c < 8; c++) v[4 * c] = x[c].x, v[4 * c + 1] = x[c].y, v[4 * c + 2] = x[c].z, v[4 * c + 3] = x[c].w;
        tmem_st_x32(tmem_base + ((q * 32u) << 16) + kAColsTmem + (step & 1u) * 32u, v);
        keep ^= v[0];
      };
      load8(chunk(i + 2), x2);
      put(i, x0);
      load8(chunk(i + 3), x0);
      put(i + 1, x1);
      load8(chunk(i + 4), x1);
      put(i + 2, x2);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    if (keep == 0x12345678u) out[1023] = keep;
  }
  ptx::mbar_wait_cluster(done_bar, 0);
  const long long t1 = clock64();
  if (threadIdx.x == 32 && rank == 0) {
    out[blockIdx.x] = static_cast<unsigned long long>(t1 - t0);
  }
  ptx::tc_fence_before();
  ptx::cluster_arrive();
  ptx::cluster_wait();
  if (warp == 1) ptx::tmem_dealloc_2sm<512>(tmem_base);
}

int main(int argc, char **argv) {
  int sms;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const uint32_t chunks = 4096, steps = argc > 1 ? atoi(argv[1]) : 20000;
  uint8_t *g;
  unsigned long long *out, h[1024];
  cudaMalloc(&g, static_cast<size_t>(chunks) * 16384);
  {  // random bytes: the toggle rate of real slices (power is what limits the real kernel)
    const size_t nbytes = static_cast<size_t>(chunks) * 16384;
    uint8_t *hbuf = static_cast<uint8_t *>(malloc(nbytes));
    uint64_t x = 88172645463325252ull;
    for (size_t i = 0; i < nbytes; i += 8) {
      x ^= x << 13; x ^= x >> 7; x ^= x << 17;
      *reinterpret_cast<uint64_t *>(hbuf + i) = x & 0x7f7f7f7f7f7f7f7full ? x : 1;
    }
    cudaMemcpy(g, hbuf, nbytes, cudaMemcpyHostToDevice);
    free(hbuf);
  }
  cudaMalloc(&out, sizeof(h));
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
  const int grid = sms / 2 * 2;
  const char *names[] = {"SS, no loads", "SS + ring (A 16 KB + B)", "TS + ring (B only)", "TS + ring (B) + LDG.nc.no_allocate -> STTM (A)",
                         "TS + ring (B) + LDG (L1) -> STTM (A)"};
  for (int mode = 0; mode < (kM == 256 ? 5 : 2); mode++) {
    cudaMemset(out, 0, sizeof(h));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<<<grid, kThreads, kSmem>>>(g, chunks, steps, mode, out);
    cudaEventRecord(e1);
    const cudaError_t err = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    double cyc = 0; int n = 0;
    for (int i = 0; i < grid; i += 2) { cyc += static_cast<double>(h[i]); n++; }
    printf("M=%u N=%u mode %d (%s): %s  %.3f ms, %.1f clk per k-step (tensor time %u), %.0f int8 TOP/s\n", kM, kN, mode, names[mode],
           cudaGetErrorString(err), ms, cyc / n / steps, kM * kN / 128, 2.0 * kM * kN * 128 * steps * (grid / 2) / (ms * 1e-3) / 1e12);
    fflush(stdout);
    if (err != cudaSuccess) return 1;
  }
  return 0;
}
