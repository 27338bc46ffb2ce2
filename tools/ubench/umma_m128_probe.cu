// Probe (sm_100a): where does tcgen05.mma.cta_group::2.kind::i8 with M = 128 (64 rows of A per CTA) put D in tensor
// memory?  One CTA pair, one MMA (M = 128, N = 128, K = 32).  A[i][0] = i - 64, A[i][1] = i - 64, B[j][0] = 127,
// B[j][1] = 1, B[j][2] = j - 64 against A[i][2] = 1  =>  D[i][j] = 128 * (i - 64) + (j - 64): every element names its
// own coordinates.  TMEM is pre-filled with a sentinel through tcgen05.st, then all 128 lanes x 256 columns of both
// CTAs are dumped and the host prints which (lane, column) holds which (row, column) of D.
// Also runs M = 256 (the shipped kernel's shape) as a control: lane = row, column = column.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/ubench/umma_m128_probe tools/ubench/umma_m128_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#include "../../ozimmu_b200/csrc/ptx.cuh"

using namespace oz;

constexpr uint32_t kN = 128, kDumpCols = 256;
constexpr uint32_t kSentinel = 0x7f7f7f7fu;

__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, uint32_t v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
      ::"r"(taddr), "r"(v) : "memory");
}

// a_rows: rows of A this CTA holds (64 for M = 128, 128 for M = 256); out: [cta][128 lanes][kDumpCols]
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe(uint32_t m_total, uint32_t a_rows, uint32_t *out) {
  __shared__ __align__(1024) uint8_t a_tile[128 * 128];
  __shared__ __align__(1024) uint8_t b_tile[64 * 128];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t rank = ptx::cluster_ctarank() & 1u;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t i = threadIdx.x; i < 128 * 128; i += 128) a_tile[i] = 0;
  for (uint32_t i = threadIdx.x; i < 64 * 128; i += 128) b_tile[i] = 0;
  __syncthreads();
  auto put = [](uint8_t *tile, uint32_t r, uint32_t k, int v) {
    tile[r * 128 + (((k >> 4) ^ (r & 7)) << 4) + (k & 15)] = static_cast<uint8_t>(static_cast<int8_t>(v));
  };
  for (uint32_t r = threadIdx.x; r < a_rows; r += 128) {
    const int i = static_cast<int>(rank * a_rows + r);           // global row of D
    // rows are encoded modulo 128 so that M = 256 stays inside int8
    const int code = (i % 128) - 64;
    put(a_tile, r, 0, code);
    put(a_tile, r, 1, code);
    put(a_tile, r, 2, 1);
  }
  for (uint32_t r = threadIdx.x; r < kN / 2; r += 128) {
    const int j = static_cast<int>(rank * (kN / 2) + r);         // global column of D
    put(b_tile, r, 0, 127);
    put(b_tile, r, 1, 1);
    put(b_tile, r, 2, j - 64);
  }
  if (threadIdx.x == 0) {
    ptx::mbar_init(ptx::smem_u32(&bar), 1);
    ptx::fence_mbar_init();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (warp == 0) ptx::tmem_alloc_2sm<512>(ptx::smem_u32(&tmem_slot));
  ptx::tc_fence_before();
  ptx::cluster_arrive();
  ptx::cluster_wait();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(&tmem_slot);
  // sentinel in this warp's 32 lanes, all dumped columns
  for (uint32_t c = 0; c < kDumpCols; c += 16) tmem_st_x16(tmem_base + ((warp * 32u) << 16) + c, kSentinel);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  ptx::tc_fence_before();
  ptx::cluster_arrive();
  ptx::cluster_wait();
  ptx::tc_fence_after();
  if (rank == 0 && warp == 1) {
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::make_i8_idesc(m_total, kN);
      ptx::mma_i8_ss_2sm(tmem_base, ptx::make_sw128_kmajor_desc(ptx::smem_u32(a_tile)),
                         ptx::make_sw128_kmajor_desc(ptx::smem_u32(b_tile)), idesc, 0u);
      ptx::tc_commit_2sm_mc(ptx::smem_u32(&bar), 0x3);
    }
    __syncwarp();
  }
  ptx::mbar_wait_cluster(ptx::smem_u32(&bar), 0);
  ptx::tc_fence_after();
  for (uint32_t c = 0; c < kDumpCols; c += 16) {
    uint32_t v[16];
    ptx::tmem_ld_x16(tmem_base + ((warp * 32u) << 16) + c, v);
    ptx::tmem_ld_wait();
    for (uint32_t j = 0; j < 16; j++)
      out[(static_cast<size_t>(rank) * 128 + warp * 32 + lane) * kDumpCols + c + j] = v[j];
  }
  ptx::tc_fence_before();
  ptx::cluster_arrive();
  ptx::cluster_wait();
  if (warp == 0) ptx::tmem_dealloc_2sm<512>(tmem_base);
}

static void report(const char *name, const std::vector<uint32_t> &h) {
  std::printf("== %s ==\n", name);
  const int lanes[] = {0, 1, 15, 16, 31, 32, 33, 63, 64, 65, 95, 96, 127};
  for (int cta = 0; cta < 2; cta++) {
    for (int lane : lanes) {
      // the row of D this lane holds (rows are encoded modulo 128), the TMEM columns that hold data, and the D column
      // found at a few TMEM columns
      int row = -1, first = -1, last = -1, filled = 0;
      int at[5] = {-1, -1, -1, -1, -1};
      const int probe_cols[5] = {0, 1, 63, 64, 127};
      for (int col = 0; col < static_cast<int>(kDumpCols); col++) {
        const uint32_t u = h[(static_cast<size_t>(cta) * 128 + lane) * kDumpCols + col];
        if (u == kSentinel) continue;
        const int t = static_cast<int>(u) + 64 * 128 + 64;   // u = 128 * (i % 128 - 64) + (j - 64)
        const int i = t / 128, j = t % 128;
        filled++;
        if (first < 0) first = col;
        last = col;
        if (row < 0) row = i;
        else if (row != i) row = -2;   // a lane holding several rows
        for (int q = 0; q < 5; q++)
          if (col == probe_cols[q]) at[q] = j;
      }
      std::printf("cta %d lane %3d: row %4d  %3d tmem cols in [%d, %d]  D col at tmem col 0/1/63/64/127 = %d / %d / %d / %d / %d\n",
                  cta, lane, row, filled, first, last, at[0], at[1], at[2], at[3], at[4]);
    }
  }
}

int main() {
  uint32_t *d = nullptr;
  const size_t n = 2 * 128 * kDumpCols;
  cudaMalloc(&d, n * sizeof(uint32_t));
  std::vector<uint32_t> h(n);
  for (uint32_t m_total : {256u, 128u}) {
    cudaMemset(d, 0xff, n * sizeof(uint32_t));
    probe<<<2, 128>>>(m_total, m_total / 2, d);
    const cudaError_t e = cudaDeviceSynchronize();
    std::printf("M = %u: %s\n", m_total, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    cudaMemcpy(h.data(), d, n * sizeof(uint32_t), cudaMemcpyDeviceToHost);
    report(m_total == 256 ? "cta_group::2 M=256 N=128 (control)" : "cta_group::2 M=128 N=128", h);
  }
  return 0;
}
