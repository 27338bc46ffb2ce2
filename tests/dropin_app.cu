// A plain cuBLAS application (no knowledge of ozIMMU): C = alpha*A*B + beta*C with cublasDgemm, operands read
// from / result written to raw files.  tests/test_gpu_dropin_cpp.py runs it with and without
// LD_PRELOAD=libozimmu.so.   usage: dropin_app n in.bin out.bin   (in.bin: A | B | C, column-major n x n doubles)
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cublas_v2.h>
#include <cuda_runtime.h>

int main(int argc, char **argv) {
  if (argc < 4) return 2;
  const int n = std::atoi(argv[1]);
  const size_t cnt = static_cast<size_t>(n) * n;
  std::vector<double> h(3 * cnt);
  FILE *f = std::fopen(argv[2], "rb");
  if (!f || std::fread(h.data(), sizeof(double), 3 * cnt, f) != 3 * cnt) return 3;
  std::fclose(f);
  double *d = nullptr;
  if (cudaMalloc(&d, 3 * cnt * sizeof(double)) != cudaSuccess) return 4;
  cudaMemcpy(d, h.data(), 3 * cnt * sizeof(double), cudaMemcpyHostToDevice);
  cublasHandle_t handle;
  if (cublasCreate(&handle) != CUBLAS_STATUS_SUCCESS) return 5;
  const double alpha = 1.5, beta = -0.5;
  const cublasStatus_t st = cublasDgemm(handle, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &alpha, d, n, d + cnt, n, &beta,
                                        d + 2 * cnt, n);
  if (st != CUBLAS_STATUS_SUCCESS) return 6;
  cudaDeviceSynchronize();
  cudaMemcpy(h.data(), d + 2 * cnt, cnt * sizeof(double), cudaMemcpyDeviceToHost);
  f = std::fopen(argv[3], "wb");
  std::fwrite(h.data(), sizeof(double), cnt, f);
  std::fclose(f);
  cublasDestroy(handle);
  cudaFree(d);
  std::printf("dropin_app done n=%d\n", n);
  return 0;
}
