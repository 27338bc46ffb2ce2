// Microbenchmark (sm_100a): is SM ingest from a peer SM's shared memory (DSMEM) additive to ingest from L2?
// Cluster of 2 CTAs, one CTA per SM.  Role L: one thread streams 16 KB bulk copies L2 -> SMEM through an 8-deep
// ring (the fused GEMM kernel's operand path, minus the tensor maps).  Role D: 8 warps read the peer CTA's shared
// memory with 16-byte ld.shared::cluster loads.  Modes: 1 = L only, 2 = D only, 3 = both.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sm_ingest sm_ingest.cu && ./sm_ingest
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

constexpr uint32_t kChunk = 16384, kRing = 8, kPeerBytes = 65536;
constexpr uint32_t kSmem = kRing * kChunk + kPeerBytes + 256;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(288, 1)
ingest(const uint8_t *__restrict__ g, uint32_t chunks_in_buffer, uint32_t iters_l, uint32_t iters_d, int mode,
       unsigned long long *out, const __grid_constant__ CUtensorMap tmap, uint32_t tiles_x, uint32_t tiles_y) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  const uint32_t bar0 = base + kRing * kChunk + kPeerBytes;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) {
    for (uint32_t i = 0; i < kRing; i++)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8 * i) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (uint32_t i = threadIdx.x; i < kPeerBytes / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem + kRing * kChunk)[i] = i;
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  long long t0 = clock64(), t1 = t0;
  unsigned long long bytes = 0;
  if (threadIdx.x == 0) {
    if (mode & 1) {
      uint32_t c = blockIdx.x * 7u;
      for (uint32_t it = 0; it < iters_l; it++) {
        const uint32_t slot = it % kRing, bar = bar0 + 8 * slot;
        if (it >= kRing) {
          const uint32_t parity = ((it / kRing) - 1) & 1u;
          uint32_t ok;
          do {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                         : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
          } while (!ok);
        }
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kChunk) : "memory");
        const uint8_t *src = g + static_cast<size_t>(c % chunks_in_buffer) * kChunk;
        c += 13;
        if (mode & 4) {  // tiled TMA box of 128 B x 128 rows (what the GEMM kernel issues)
          const uint32_t ty = tiles_y & 0xFFFFu, bx = (tiles_y >> 16) & 0xFFu ? ((tiles_y >> 16) & 0xFFu) : 256u, by = tiles_y >> 24;
          const uint32_t t = c % (tiles_x * ty);
          const int cx = static_cast<int>((t % tiles_x) * bx), cy = static_cast<int>((t / tiles_x) * by);
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                       ::"r"(base + slot * kChunk), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(bar), "r"(cx), "r"(cy) : "memory");
        } else {
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(base + slot * kChunk), "l"(src), "r"(kChunk), "r"(bar) : "memory");
        }
        bytes += kChunk;
      }
      for (uint32_t it = (iters_l > kRing ? iters_l - kRing : 0); it < iters_l; it++) {  // drain
        const uint32_t slot = it % kRing, bar = bar0 + 8 * slot, parity = (it / kRing) & 1u;
        uint32_t ok;
        do {
          asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                       : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        } while (!ok);
      }
    }
    t1 = clock64();
    out[blockIdx.x * 4 + 0] = bytes;
    out[blockIdx.x * 4 + 1] = static_cast<unsigned long long>(t1 - t0);
  } else if (threadIdx.x >= 32) {
    if (mode & 2) {
      uint32_t peer_base;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(peer_base) : "r"(base + kRing * kChunk), "r"(rank ^ 1u));
      const uint32_t tid = threadIdx.x - 32;  // 0..255
      uint32_t acc = 0;
      for (uint32_t it = 0; it < iters_d; it++) {
#pragma unroll 4
        for (uint32_t off = tid * 16; off < kPeerBytes; off += 256 * 16) {
          uint32_t a, b, c, d;
          asm volatile("ld.shared::cluster.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(peer_base + off));
          acc += a ^ b ^ c ^ d;
        }
      }
      t1 = clock64();
      if (tid == 0) {
        out[blockIdx.x * 4 + 2] = static_cast<unsigned long long>(iters_d) * kPeerBytes;
        out[blockIdx.x * 4 + 3] = static_cast<unsigned long long>(t1 - t0);
      }
      if (acc == 0x12345678u) out[0] = acc;
    }
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

int main() {
  int sms;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const uint32_t chunks = 4096;  // 64 MB buffer: L2-resident after the first sweep
  uint8_t *g;
  unsigned long long *out, *h;
  cudaMalloc(&g, static_cast<size_t>(chunks) * kChunk);
  cudaMemset(g, 1, static_cast<size_t>(chunks) * kChunk);
  cudaMalloc(&out, sizeof(unsigned long long) * 4 * 256);
  h = new unsigned long long[4 * 256];
  cudaFuncSetAttribute(ingest, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
  const int grid = sms / 2 * 2;
  using EncodeFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void *fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeFn enc = reinterpret_cast<EncodeFn>(fp);
  struct Shape { const char *name; cuuint64_t inner, rows; CUtensorMapDataType dt; cuuint32_t esz, bx, by; CUtensorMapSwizzle sw; } shapes[] = {
      {"tiled TMA u8 box 128 B x 128 rows SW128, matrix [8192][8192 B]", 8192, 8192, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, 128, 128, CU_TENSOR_MAP_SWIZZLE_128B},
      {"tiled TMA u8 box 128 B x 128 rows SW128, matrix [524288][128 B] (contiguous 16 KB)", 128, 524288, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, 128, 128, CU_TENSOR_MAP_SWIZZLE_128B},
      {"tiled TMA u64 box 2 KB x 8 rows no swizzle, matrix [32768][2 KB] (contiguous 16 KB)", 256, 32768, CU_TENSOR_MAP_DATA_TYPE_UINT64, 8, 256, 8, CU_TENSOR_MAP_SWIZZLE_NONE},
      {"tiled TMA u32 box 1 KB x 16 rows no swizzle, matrix [65536][1 KB] (contiguous 16 KB)", 256, 65536, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, 256, 16, CU_TENSOR_MAP_SWIZZLE_NONE},
      {"tiled TMA u8 box 256 B x 64 rows no swizzle, matrix [262144][256 B] (contiguous 16 KB)", 256, 262144, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, 256, 64, CU_TENSOR_MAP_SWIZZLE_NONE}};
  for (auto &sh : shapes) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {sh.inner, sh.rows};
    cuuint64_t strides[1] = {sh.inner * sh.esz};
    cuuint32_t box[2] = {sh.bx, sh.by}, es[2] = {1, 1};
    CUresult r = enc(&tm, sh.dt, 2, g, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     sh.sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
    for (int rep = 0; rep < 2; rep++) {
      cudaMemset(out, 0, sizeof(unsigned long long) * 4 * 256);
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      ingest<<<grid, 288, kSmem>>>(g, chunks, 6000, 0, 5, out, tm, static_cast<uint32_t>(sh.inner / sh.bx), static_cast<uint32_t>(sh.rows / sh.by) | (sh.bx << 16) | (sh.by << 24));
      cudaEventRecord(e1);
      cudaError_t err = cudaDeviceSynchronize();
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      cudaMemcpy(h, out, sizeof(unsigned long long) * 4 * grid, cudaMemcpyDeviceToHost);
      double lb = 0, lc = 0;
      for (int b = 0; b < grid; b++) { lb += h[4 * b]; lc += h[4 * b + 1]; }
      printf("%s: %s  %.3f ms | L2->SMEM %.1f B/clk/SM (%.2f TB/s chip)\n", sh.name, cudaGetErrorString(err), ms, lb / lc,
             lb / (ms * 1e-3) / 1e12);
    }
  }
  CUtensorMap tm0{};
  for (int mode : {1, 2, 3, 1, 3}) {
    cudaMemset(out, 0, sizeof(unsigned long long) * 4 * 256);
    const uint32_t iters_l = 6000, iters_d = (mode == 2) ? 1500 : 1500;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    ingest<<<grid, 288, kSmem>>>(g, chunks, iters_l, iters_d, mode, out, tm0, 1, 1);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(h, out, sizeof(unsigned long long) * 4 * grid, cudaMemcpyDeviceToHost);
    double lb = 0, lc = 0, db = 0, dc = 0;
    for (int b = 0; b < grid; b++) { lb += h[4 * b]; lc += h[4 * b + 1]; db += h[4 * b + 2]; dc += h[4 * b + 3]; }
    printf("mode %d (%s): %s  kernel %.3f ms | L2->SMEM %.1f B/clk/SM (%.2f TB/s chip) | DSMEM reads %.1f B/clk/SM (%.2f TB/s chip)\n",
           mode, mode == 1 ? "L2 only" : mode == 2 ? "DSMEM only" : "both", cudaGetErrorString(err), ms,
           lc > 0 ? lb / lc * 1.0 : 0.0, (mode & 1) ? lb / (ms * 1e-3) / 1e12 : 0.0, dc > 0 ? db / dc : 0.0,
           (mode & 2) ? db / (ms * 1e-3) / 1e12 : 0.0);
  }
  return 0;
}
