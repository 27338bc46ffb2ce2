// Microbenchmark: per-SM throughput of the instructions the fused epilogue needs (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_rates fp64_rates.cu && ./fp64_rates
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kIters = 4096;

template <int MODE>
__global__ void __launch_bounds__(256) k(double *out, const int *in, double s) {
  int p[8];
  double acc[8];
#pragma unroll
  for (int j = 0; j < 8; j++) { p[j] = in[threadIdx.x + j * 256]; acc[j] = j; }
  for (int it = 0; it < kIters; it++) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
      if (MODE == 0) {            // DFMA only
        acc[j] = __fma_rn(acc[j], s, 1.0);
      } else if (MODE == 1) {     // I2F.F64 + DFMA (current epilogue)
        acc[j] = __fma_rn(__int2double_rn(p[j]), s, acc[j]);
        p[j] += it;
      } else if (MODE == 2) {     // magic-number conversion: LOP3 + DADD + DFMA
        const double m = __hiloint2double(0x43300000, p[j] ^ 0x80000000);
        acc[j] = __fma_rn(m - 4503601774854144.0, s, acc[j]);
        p[j] += it;
      } else if (MODE == 3) {     // DADD only
        acc[j] = __dadd_rn(acc[j], s);
      } else if (MODE == 4) {     // I2F.F64 only (result consumed by integer xor to keep it alive)
        const double d = __int2double_rn(p[j]);
        p[j] = (p[j] + it) ^ __double2hiint(d);
      } else if (MODE == 5) {     // integer-only int32 -> double bit pattern, then DFMA
        const int v = p[j];
        const uint32_t sgn = static_cast<uint32_t>(v) & 0x80000000u;
        const uint32_t a = static_cast<uint32_t>(v < 0 ? -v : v);
        const int lz = __clz(a);
        const uint32_t t = a << lz;
        uint32_t hi = sgn | (static_cast<uint32_t>(1023 + 31 - lz) << 20) | ((t & 0x7FFFFFFFu) >> 11);
        uint32_t lo = t << 21;
        if (a == 0) { hi = 0; lo = 0; }
        acc[j] = __fma_rn(__hiloint2double(hi, lo), s, acc[j]);
        p[j] += it;
      } else if (MODE == 6) {     // FP32 FFMA reference
        float f = __int_as_float(p[j]);
        f = __fmaf_rn(f, 1.0001f, 0.5f);
        p[j] = __float_as_int(f);
      }
    }
  }
  double r = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) r += acc[j] + p[j];
  out[blockIdx.x * 256 + threadIdx.x] = r;
}

template <int MODE>
void run(const char *name, int ops_per_elem, double *out, int *in, int sms, int warps_per_sm) {
  const int blocks = sms * warps_per_sm / 8;
  k<MODE><<<blocks, 256>>>(out, in, 1.0000001);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<blocks, 256>>>(out, in, 1.0000001);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const double elems = double(blocks) * 256 * 8 * kIters;
  printf("%-34s warps/SM=%2d  %8.3f ms  %7.2f Gelem/s  => %6.2f elem/clk/SM at %d MHz nominal (%d fp64-pipe ops/elem)\n",
         name, warps_per_sm, ms, elems / ms / 1e6, elems / (ms * 1e-3) / sms / (clk_khz * 1e3), clk_khz / 1000, ops_per_elem);
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double *out; int *in;
  cudaMalloc(&out, sizeof(double) * sms * 64 * 256);
  cudaMalloc(&in, sizeof(int) * 256 * 8);
  cudaMemset(in, 1, sizeof(int) * 256 * 8);
  for (int w : {8, 32}) {
    run<0>("DFMA", 1, out, in, sms, w);
    run<3>("DADD", 1, out, in, sms, w);
    run<4>("I2F.F64.S32", 1, out, in, sms, w);
    run<1>("I2F.F64 + DFMA", 2, out, in, sms, w);
    run<2>("xor + DADD(magic) + DFMA", 2, out, in, sms, w);
    run<5>("int-only convert + DFMA", 1, out, in, sms, w);
    run<6>("FFMA (fp32 reference)", 0, out, in, sms, w);
  }
  return 0;
}
