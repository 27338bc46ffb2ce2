// A plain cuBLAS application (no knowledge of ozIMMU): C = alpha*A*B + beta*C with cublasDgemm, operands read
// from / result written to raw files.  tests/test_gpu_dropin_cpp.py runs it with and without
// LD_PRELOAD=libozimmu.so.   usage: dropin_app n in.bin out.bin [mode]   (in.bin: A | B | C, column-major n x n doubles)
//   mode host     (default) one handle, host pointer mode
//        devptr   alpha / beta live in device memory (CUBLAS_POINTER_MODE_DEVICE)
//        threads  two host threads, each with its own cuBLAS handle and stream, three DGEMMs each, concurrently;
//                 out.bin = C of thread 0 | C of thread 1 (the same product twice)
//        multigpu one process, one cuBLAS handle per GPU (devices 0 and 1); out.bin = C of GPU 0 | C of GPU 1
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include <cublas_v2.h>
#include <cuda_runtime.h>

static int run_one(int device, int n, const std::vector<double> &h, double *out, bool devptr, int repeats) {
  const size_t cnt = static_cast<size_t>(n) * n;
  if (cudaSetDevice(device) != cudaSuccess) return 10;
  double *d = nullptr, *scal = nullptr;
  if (cudaMalloc(&d, 3 * cnt * sizeof(double)) != cudaSuccess) return 4;
  cudaStream_t st;
  cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  cublasHandle_t handle;
  if (cublasCreate(&handle) != CUBLAS_STATUS_SUCCESS) return 5;
  cublasSetStream(handle, st);
  const double alpha = 1.5, beta = -0.5;
  const double *pa = &alpha, *pb = &beta;
  if (devptr) {
    cudaMalloc(&scal, 2 * sizeof(double));
    cudaMemcpy(scal, &alpha, sizeof(double), cudaMemcpyHostToDevice);
    cudaMemcpy(scal + 1, &beta, sizeof(double), cudaMemcpyHostToDevice);
    cublasSetPointerMode(handle, CUBLAS_POINTER_MODE_DEVICE);
    pa = scal, pb = scal + 1;
  }
  for (int r = 0; r < repeats; r++) {
    cudaMemcpyAsync(d, h.data(), 3 * cnt * sizeof(double), cudaMemcpyHostToDevice, st);
    const cublasStatus_t s = cublasDgemm(handle, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, pa, d, n, d + cnt, n, pb, d + 2 * cnt, n);
    if (s != CUBLAS_STATUS_SUCCESS) return 6;
  }
  cudaMemcpyAsync(out, d + 2 * cnt, cnt * sizeof(double), cudaMemcpyDeviceToHost, st);
  if (cudaStreamSynchronize(st) != cudaSuccess) return 7;
  cublasDestroy(handle);
  cudaFree(d);
  cudaFree(scal);
  cudaStreamDestroy(st);
  return 0;
}

int main(int argc, char **argv) {
  if (argc < 4) return 2;
  const int n = std::atoi(argv[1]);
  const char *mode = argc > 4 ? argv[4] : "host";
  const size_t cnt = static_cast<size_t>(n) * n;
  std::vector<double> h(3 * cnt);
  FILE *f = std::fopen(argv[2], "rb");
  if (!f || std::fread(h.data(), sizeof(double), 3 * cnt, f) != 3 * cnt) return 3;
  std::fclose(f);
  std::vector<double> out;
  int rc = 0;
  if (!std::strcmp(mode, "threads") || !std::strcmp(mode, "multigpu")) {
    const bool multi = !std::strcmp(mode, "multigpu");
    int ndev = 0;
    cudaGetDeviceCount(&ndev);
    if (multi && ndev < 2) return 9;
    out.resize(2 * cnt);
    int rcs[2] = {0, 0};
    std::thread t0([&] { rcs[0] = run_one(0, n, h, out.data(), false, 3); });
    std::thread t1([&] { rcs[1] = run_one(multi ? 1 : 0, n, h, out.data() + cnt, false, 3); });
    t0.join();
    t1.join();
    rc = rcs[0] ? rcs[0] : rcs[1];
  } else {
    out.resize(cnt);
    rc = run_one(0, n, h, out.data(), !std::strcmp(mode, "devptr"), 1);
  }
  if (rc) return rc;
  f = std::fopen(argv[3], "wb");
  std::fwrite(out.data(), sizeof(double), out.size(), f);
  std::fclose(f);
  std::printf("dropin_app done n=%d mode=%s\n", n, mode);
  return 0;
}
