// Microbenchmark (sm_100a): the operand path of a tile-contiguous slice layout.  Cluster of 2 CTAs (one per SM).
// Each CTA's producer thread issues TWO linear 16 KB bulk copies per stage (an A tile and a B tile) through a
// 5-stage ring; `shared` > 1 makes groups of `shared` CTAs read the same chunks at the same time (as the CTAs of one
// tile row / tile column of the GEMM do).  remote = 1: the non-leader CTA completes its copies on the LEADER's
// mbarrier (shared::cluster address) -- is that legal for non-tensor bulk copies?
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#ifndef CHUNK
#define CHUNK 16384
#endif
#ifndef RING
#define RING 5
#endif
constexpr uint32_t kChunk = CHUNK, kRing = RING;   // -DCHUNK=8192 -DRING=12: the stages of the 128 x 128 tile
constexpr uint32_t kSmem = kRing * 2 * kChunk + 256;
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void wait_parity(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
k(const uint8_t *__restrict__ g, uint32_t chunks, uint32_t iters, uint32_t shared, int remote, unsigned long long *out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem), bar0 = base + kRing * 2 * kChunk;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) {
    for (uint32_t i = 0; i < kRing; i++)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8 * i) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    const uint32_t grp = blockIdx.x / shared;
    // remote mode: rank 1 signals rank 0's barriers; rank 0 expects both CTAs' bytes and is the only one that waits
    for (uint32_t it = 0; it < iters; it++) {
      const uint32_t slot = it % kRing, bar = bar0 + 8 * slot;
      uint32_t tgt = bar;
      if (remote) asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(tgt) : "r"(bar));
      if (it >= kRing && !(remote && rank == 1)) wait_parity(bar, ((it / kRing) - 1) & 1u);
      if (!remote) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(2 * kChunk) : "memory");
      else if (rank == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(4 * kChunk) : "memory");
      // chunks is a power of two: no division in the issuing thread's loop
      const uint8_t *sa = g + static_cast<size_t>((grp * 977u + it * 2u) & (chunks - 1u)) * kChunk;
      const uint8_t *sb = g + static_cast<size_t>((grp * 977u + it * 2u + 1u + 31u * (blockIdx.x & (shared - 1u))) & (chunks - 1u)) * kChunk;
#ifdef ONECOPY
      (void)sb;   // one copy of 2 x kChunk contiguous bytes per stage
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(base + slot * 2 * kChunk), "l"(g + static_cast<size_t>((grp * 977u + it * 2u) & (chunks - 2u)) * kChunk),
                     "r"(2 * kChunk), "r"(tgt) : "memory");
#else
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(base + slot * 2 * kChunk), "l"(sa), "r"(kChunk), "r"(tgt) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(base + slot * 2 * kChunk + kChunk), "l"(sb), "r"(kChunk), "r"(tgt) : "memory");
#endif
    }
    if (!(remote && rank == 1))
      for (uint32_t it = (iters > kRing ? iters - kRing : 0); it < iters; it++) wait_parity(bar0 + 8 * (it % kRing), (it / kRing) & 1u);
    const long long t1 = clock64();
    out[blockIdx.x * 2] = static_cast<unsigned long long>(iters) * 2 * kChunk;
    out[blockIdx.x * 2 + 1] = static_cast<unsigned long long>(t1 - t0);
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// usage: bulk_pair [grid (CTAs, even; default: all SMs)] [chunks in the working set (default 4096)]
int main(int argc, char **argv) {
  int sms;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const uint32_t chunks = argc > 2 ? static_cast<uint32_t>(atoi(argv[2])) : 4096;
  uint8_t *g;
  unsigned long long *out, *h = new unsigned long long[512];
  cudaMalloc(&g, static_cast<size_t>(chunks) * kChunk);
  cudaMemset(g, 1, static_cast<size_t>(chunks) * kChunk);
  cudaMalloc(&out, sizeof(unsigned long long) * 512);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
  const int grid = argc > 1 && atoi(argv[1]) > 0 ? atoi(argv[1]) / 2 * 2 : sms / 2 * 2;
  struct Cfg { uint32_t shared; int remote; uint32_t iters; } cfgs[] = {{1, 0, 4000}, {1, 0, 4000}, {8, 0, 4000}, {16, 0, 4000},
                                                                        {1, 1, 5}};  // remote: legality only (one ring pass)
  for (auto c : cfgs) {
    cudaMemset(out, 0, sizeof(unsigned long long) * 512);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<<<grid, 128, kSmem>>>(g, chunks, c.iters, c.shared, c.remote, out);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(h, out, sizeof(unsigned long long) * 2 * grid, cudaMemcpyDeviceToHost);
    double b = 0, cyc = 0;
    for (int i = 0; i < grid; i++) { b += h[2 * i]; cyc += h[2 * i + 1]; }
    printf("2 x %u KB linear bulk copies per stage, %u stages, %d CTAs, working set %.0f MB, %u CTAs share each A chunk, remote-barrier=%d: %s %.3f ms, %.1f B/clk/SM, %.2f TB/s chip\n",
           kChunk / 1024, kRing, grid, chunks * (kChunk / 1048576.0), c.shared, c.remote, cudaGetErrorString(err), ms, cyc > 0 ? b / cyc : 0.0, b / (ms * 1e-3) / 1e12);
    fflush(stdout);
    if (err != cudaSuccess) break;
  }
  return 0;
}
