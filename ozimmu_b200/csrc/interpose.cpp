// interpose.cpp -- the LD_PRELOAD surface: cuBLAS GEMM entry points that route FP64 GEMMs through
// the Ozaki-scheme path and pass everything else to the real cuBLAS.
//
// Replaces reference src/cublas.cu:18-131 (env parsing, global handle, create/destroy hooks),
// :133-295 (cublasGemmEx, cublasDgemm_v2), :297-313 (cublasZgemm_v2) and :315-512 (strided
// batched).  Same symbols, same env knobs (OZIMMU_COMPUTE_MODE re-read on every call,
// OZIMMU_INTERCEPT_THRESHOLD_{M,N,K}, OZIMMU_AUTO_AVG_MANTISSA_LOSS_THRESHOLD,
// OZIMMU_MALLOC_ASYNC, OZIMMU_ENABLE_CULIP_PROFILING).  Deliberate fixes of reference defects
// (SURVEY App. B.3/B.4/B.7): n is compared with THRESHOLD_N, a failed Ozaki call reports
// CUBLAS_STATUS_INTERNAL_ERROR, the global handle is created on first use (no null deref when
// the application's cublasCreate ran before this library was loaded), and device-pointer-mode
// scalars fall through to the real cuBLAS instead of being dereferenced on the host.
// Complex (CUDA_C_64F) GEMMs take the complex Ozaki path (reference src/gemm.cu:412-521) unless an
// operand is CUBLAS_OP_C: the reference silently treats OP_C as OP_T (src/cublas.cu:50-56), which
// is wrong for complex data, so those calls go to the real cuBLAS instead.
#include <cstring>
#include <mutex>

#include "host.hpp"
#include "ozimmu_b200.h"

using namespace mtk::ozimmu;
namespace H = oz::host;

namespace {

std::mutex g_mu;
handle_t g_handle = nullptr;

// reference src/cublas.cu:18-48: unknown / unset -> dgemm (passthrough)
compute_mode_t env_compute_mode() {
  const char *v = std::getenv("OZIMMU_COMPUTE_MODE");
  if (v == nullptr) return dgemm;
  for (int mode = sgemm; mode <= fp64_int8_auto; mode++)
    if (get_compute_mode_name_str(static_cast<compute_mode_t>(mode)) == v) return static_cast<compute_mode_t>(mode);
  return dgemm;
}

// reference src/cublas.cu:60-86
handle_t global_handle() {
  if (g_handle == nullptr) {
    const malloc_mode_t mm = H::env_enabled("OZIMMU_MALLOC_ASYNC", false) ? malloc_async : malloc_sync;
    H::log_info("Initializing ozIMMU handle...");
    create(&g_handle, mm);
    H::log_info("Successfully initialized");
  }
  if (const char *t = std::getenv("OZIMMU_AUTO_AVG_MANTISSA_LOSS_THRESHOLD")) {
    char *end = nullptr;
    const double v = std::strtod(t, &end);
    if (end == t) throw std::runtime_error(std::string("ERROR: invalid OZIMMU_AUTO_AVG_MANTISSA_LOSS_THRESHOLD = ") + t);
    set_auto_mantissa_loss_threashold(g_handle, v);
  }
  return g_handle;
}

template <class Fn>
Fn real_fn(const char *name) {
  return reinterpret_cast<Fn>(H::real_cublas_symbol(name));
}

using GemmExFn = cublasStatus_t (*)(cublasHandle_t, cublasOperation_t, cublasOperation_t, int, int, int, const void *,
                                    const void *, cudaDataType_t, int, const void *, cudaDataType_t, int, const void *,
                                    void *, cudaDataType_t, int, cublasComputeType_t, cublasGemmAlgo_t);
using GemmStridedBatchedExFn = cublasStatus_t (*)(cublasHandle_t, cublasOperation_t, cublasOperation_t, int, int, int,
                                                  const void *, const void *, cudaDataType_t, int, long long,
                                                  const void *, cudaDataType_t, int, long long, const void *, void *,
                                                  cudaDataType_t, int, long long, int, cublasComputeType_t,
                                                  cublasGemmAlgo_t);

bool host_pointer_mode(cublasHandle_t handle) {
  using Fn = cublasStatus_t (*)(cublasHandle_t, cublasPointerMode_t *);
  auto fn = real_fn<Fn>("cublasGetPointerMode_v2");
  cublasPointerMode_t mode = CUBLAS_POINTER_MODE_HOST;
  if (fn && fn(handle, &mode) != CUBLAS_STATUS_SUCCESS) return false;
  return mode == CUBLAS_POINTER_MODE_HOST;
}

cudaStream_t stream_of(cublasHandle_t handle) {
  using Fn = cublasStatus_t (*)(cublasHandle_t, cudaStream_t *);
  auto fn = real_fn<Fn>("cublasGetStream_v2");
  cudaStream_t s = nullptr;
  if (fn) fn(handle, &s);
  return s;
}

// reference src/cublas.cu:143-148 (with the THRESHOLD_N fix)
bool should_intercept(handle_t h, compute_mode_t mode, int m, int n, int k, cudaDataType_t a, cudaDataType_t b,
                      cudaDataType_t c) {
  return mode != dgemm && m >= 0 && n >= 0 && k >= 0 &&
         static_cast<std::uint32_t>(m) >= h->intercept_threshold_m &&
         static_cast<std::uint32_t>(n) >= h->intercept_threshold_n &&
         static_cast<std::uint32_t>(k) >= h->intercept_threshold_k &&
         ((a == CUDA_R_64F && b == CUDA_R_64F && c == CUDA_R_64F) || (a == CUDA_C_64F && b == CUDA_C_64F && c == CUDA_C_64F));
}

// reference src/culip.cu:14-50: one "[CULiP Result][name] ns" line per intercepted call
struct CulipScope {
  bool on;
  cudaStream_t s;
  std::string name;
  std::chrono::steady_clock::time_point t0;
  CulipScope(cudaStream_t stream, std::string n) : on(H::env_enabled("OZIMMU_ENABLE_CULIP_PROFILING", false)), s(stream), name(std::move(n)) {
    if (!on) return;
    cudaStreamSynchronize(s);
    t0 = std::chrono::steady_clock::now();
  }
  ~CulipScope() {
    if (!on) return;
    cudaStreamSynchronize(s);
    const auto ns = std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
    std::printf("[CULiP Result][%s] %lld ns\n", name.c_str(), static_cast<long long>(ns));
    std::fflush(stdout);
  }
};

const char *op_str(cublasOperation_t op) { return op == CUBLAS_OP_N ? "N" : (op == CUBLAS_OP_T ? "T" : "C"); }

cublasStatus_t ozaki_gemm(cublasHandle_t handle, compute_mode_t mode, cublasOperation_t ta, cublasOperation_t tb, int m,
                          int n, int k, const void *alpha, const void *A, int lda, const void *B, int ldb,
                          const void *beta, void *C, int ldc, element_kind_t kind) {
  std::lock_guard<std::mutex> lock(g_mu);
  try {
    handle_t h = global_handle();
    cudaStream_t s = stream_of(handle);
    set_cuda_stream(h, s);
    CulipScope scope(s, std::string(kind == mtk::ozimmu::real ? "D" : "Z") + get_compute_mode_name_str(mode) + "-" + op_str(ta) + op_str(tb) + "-m" +
                            std::to_string(m) + "-n" + std::to_string(n) + "-k" + std::to_string(k));
    // reference src/cublas.cu:50-56: everything that is not OP_N is treated as OP_T (real data)
    const int err = gemm(h, ta == CUBLAS_OP_N ? op_n : op_t, tb == CUBLAS_OP_N ? op_n : op_t, m, n, k, alpha, A, lda, B,
                         ldb, beta, C, ldc, mode, kind);
    return err ? CUBLAS_STATUS_INVALID_VALUE : CUBLAS_STATUS_SUCCESS;
  } catch (const std::exception &e) {
    H::log_error(e.what());
    return CUBLAS_STATUS_INTERNAL_ERROR;
  }
}

cublasStatus_t ozaki_dgemm(cublasHandle_t handle, compute_mode_t mode, cublasOperation_t ta, cublasOperation_t tb, int m,
                           int n, int k, const double *alpha, const double *A, int lda, const double *B, int ldb,
                           const double *beta, double *C, int ldc) {
  return ozaki_gemm(handle, mode, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, mtk::ozimmu::real);
}

// the whole strided batch through one grouped launch (the reference loops: src/cublas.cu:380-406)
cublasStatus_t ozaki_dgemm_batched(cublasHandle_t handle, compute_mode_t mode, cublasOperation_t ta, cublasOperation_t tb,
                                   int m, int n, int k, const double *alpha, const double *A, int lda, long long strideA,
                                   const double *B, int ldb, long long strideB, const double *beta, double *C, int ldc,
                                   long long strideC, int batch) {
  std::lock_guard<std::mutex> lock(g_mu);
  try {
    handle_t h = global_handle();
    cudaStream_t s = stream_of(handle);
    set_cuda_stream(h, s);
    CulipScope scope(s, std::string("D") + get_compute_mode_name_str(mode) + "-batched" + std::to_string(batch) + "-" +
                            op_str(ta) + op_str(tb) + "-m" + std::to_string(m) + "-n" + std::to_string(n) + "-k" +
                            std::to_string(k));
    const int err = gemm_strided_batched(h, ta == CUBLAS_OP_N ? op_n : op_t, tb == CUBLAS_OP_N ? op_n : op_t, m, n, k,
                                         alpha, A, lda, strideA, B, ldb, strideB, beta, C, ldc, strideC,
                                         static_cast<std::size_t>(batch), mode);
    return err ? CUBLAS_STATUS_INVALID_VALUE : CUBLAS_STATUS_SUCCESS;
  } catch (const std::exception &e) {
    H::log_error(e.what());
    return CUBLAS_STATUS_INTERNAL_ERROR;
  }
}

bool no_conj(cublasOperation_t ta, cublasOperation_t tb) { return ta != CUBLAS_OP_C && tb != CUBLAS_OP_C; }

}  // namespace

extern "C" {

// reference src/cublas.cu:104-115
cublasStatus_t cublasCreate_v2(cublasHandle_t *handle) {
  auto fn = real_fn<cublasStatus_t (*)(cublasHandle_t *)>("cublasCreate_v2");
  if (fn == nullptr) return CUBLAS_STATUS_NOT_INITIALIZED;
  const cublasStatus_t st = fn(handle);
  if (st == CUBLAS_STATUS_SUCCESS && env_compute_mode() != dgemm) {
    std::lock_guard<std::mutex> lock(g_mu);
    try {
      // pre-size the workspace for one 1024^3 fp64_int8_9 product (reference src/cublas.cu:12-16)
      reallocate_working_memory(global_handle(), gemm_list_t{{op_n, op_n, 1024, 1024, 1024, mtk::ozimmu::real, fp64_int8_9}});
    } catch (const std::exception &e) {
      H::log_error(e.what());
    }
  }
  return st;
}

// reference src/cublas.cu:117-131.  The reference tears its global handle down on ANY
// cublasDestroy; here it lives until the process ends (another cuBLAS handle may still be in use).
cublasStatus_t cublasDestroy_v2(cublasHandle_t handle) {
  auto fn = real_fn<cublasStatus_t (*)(cublasHandle_t)>("cublasDestroy_v2");
  if (fn == nullptr) return CUBLAS_STATUS_NOT_INITIALIZED;
  return fn(handle);
}

// reference src/cublas.cu:133-278
cublasStatus_t cublasGemmEx(cublasHandle_t handle, cublasOperation_t transa, cublasOperation_t transb, int m, int n, int k,
                            const void *alpha, const void *A, cudaDataType_t Atype, int lda, const void *B,
                            cudaDataType_t Btype, int ldb, const void *beta, void *C, cudaDataType_t Ctype, int ldc,
                            cublasComputeType_t computeType, cublasGemmAlgo_t algo) {
  const compute_mode_t mode = env_compute_mode();
  const bool is_real = Atype == CUDA_R_64F && Btype == CUDA_R_64F && Ctype == CUDA_R_64F;
  const bool is_cplx = Atype == CUDA_C_64F && Btype == CUDA_C_64F && Ctype == CUDA_C_64F && no_conj(transa, transb);
  if (mode != dgemm && (is_real || is_cplx)) {
    bool take = false;
    {
      std::lock_guard<std::mutex> lock(g_mu);
      try {
        take = should_intercept(global_handle(), mode, m, n, k, Atype, Btype, Ctype) && host_pointer_mode(handle);
      } catch (const std::exception &e) {
        H::log_error(e.what());
      }
    }
    if (take)
      return ozaki_gemm(handle, mode, transa, transb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc,
                        is_real ? mtk::ozimmu::real : complx);
  }
  auto fn = real_fn<GemmExFn>("cublasGemmEx");
  if (fn == nullptr) return CUBLAS_STATUS_NOT_INITIALIZED;
  CulipScope scope(H::env_enabled("OZIMMU_ENABLE_CULIP_PROFILING", false) ? stream_of(handle) : nullptr,
                   std::string("cublasGemmEx-") + op_str(transa) + op_str(transb) + "-m" + std::to_string(m) + "-n" +
                       std::to_string(n) + "-k" + std::to_string(k));
  return fn(handle, transa, transb, m, n, k, alpha, A, Atype, lda, B, Btype, ldb, beta, C, Ctype, ldc, computeType, algo);
}

// reference src/cublas.cu:280-295
cublasStatus_t cublasDgemm_v2(cublasHandle_t handle, cublasOperation_t transa, cublasOperation_t transb, int m, int n,
                              int k, const double *alpha, const double *A, int lda, const double *B, int ldb,
                              const double *beta, double *C, int ldc) {
  const compute_mode_t mode = env_compute_mode();
  if (mode != dgemm) {
    bool take = false;
    {
      std::lock_guard<std::mutex> lock(g_mu);
      try {
        take = should_intercept(global_handle(), mode, m, n, k, CUDA_R_64F, CUDA_R_64F, CUDA_R_64F) &&
               host_pointer_mode(handle);
      } catch (const std::exception &e) {
        H::log_error(e.what());
      }
    }
    if (take) return ozaki_dgemm(handle, mode, transa, transb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
  }
  auto fn = real_fn<cublasStatus_t (*)(cublasHandle_t, cublasOperation_t, cublasOperation_t, int, int, int, const double *,
                                    const double *, int, const double *, int, const double *, double *, int)>(
      "cublasDgemm_v2");
  if (fn == nullptr) return CUBLAS_STATUS_NOT_INITIALIZED;
  return fn(handle, transa, transb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
}

// reference src/cublas.cu:297-313
cublasStatus_t cublasZgemm_v2(cublasHandle_t handle, cublasOperation_t transa, cublasOperation_t transb, int m, int n,
                              int k, const cuDoubleComplex *alpha, const cuDoubleComplex *A, int lda,
                              const cuDoubleComplex *B, int ldb, const cuDoubleComplex *beta, cuDoubleComplex *C,
                              int ldc) {
  const compute_mode_t mode = env_compute_mode();
  if (mode != dgemm && no_conj(transa, transb)) {
    bool take = false;
    {
      std::lock_guard<std::mutex> lock(g_mu);
      try {
        take = should_intercept(global_handle(), mode, m, n, k, CUDA_C_64F, CUDA_C_64F, CUDA_C_64F) &&
               host_pointer_mode(handle);
      } catch (const std::exception &e) {
        H::log_error(e.what());
      }
    }
    if (take) return ozaki_gemm(handle, mode, transa, transb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, complx);
  }
  auto fn = real_fn<cublasStatus_t (*)(cublasHandle_t, cublasOperation_t, cublasOperation_t, int, int, int,
                                    const cuDoubleComplex *, const cuDoubleComplex *, int, const cuDoubleComplex *, int,
                                    const cuDoubleComplex *, cuDoubleComplex *, int)>("cublasZgemm_v2");
  if (fn == nullptr) return CUBLAS_STATUS_NOT_INITIALIZED;
  return fn(handle, transa, transb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
}

// reference src/cublas.cu:315-472 (one Ozaki GEMM per batch entry, :380-406): here one grouped launch
cublasStatus_t cublasGemmStridedBatchedEx(cublasHandle_t handle, cublasOperation_t transa, cublasOperation_t transb, int m,
                                          int n, int k, const void *alpha, const void *A, cudaDataType_t Atype, int lda,
                                          long long strideA, const void *B, cudaDataType_t Btype, int ldb,
                                          long long strideB, const void *beta, void *C, cudaDataType_t Ctype, int ldc,
                                          long long strideC, int batchCount, cublasComputeType_t computeType,
                                          cublasGemmAlgo_t algo) {
  const compute_mode_t mode = env_compute_mode();
  if (mode != dgemm && Atype == CUDA_R_64F && Btype == CUDA_R_64F && Ctype == CUDA_R_64F) {
    bool take = false;
    {
      std::lock_guard<std::mutex> lock(g_mu);
      try {
        take = should_intercept(global_handle(), mode, m, n, k, Atype, Btype, Ctype) && host_pointer_mode(handle);
      } catch (const std::exception &e) {
        H::log_error(e.what());
      }
    }
    if (take && batchCount > 0) {
      const auto *a = static_cast<const double *>(A);
      const auto *b = static_cast<const double *>(B);
      auto *c = static_cast<double *>(C);
      // entries of C that overlap (or walk backwards) cannot run concurrently: entry by entry, as the reference
      const unsigned long long span = static_cast<unsigned long long>(ldc) * (n > 0 ? n - 1 : 0) + m;
      if (batchCount == 1 || strideC < 0 || static_cast<unsigned long long>(strideC) < span) {
        for (int i = 0; i < batchCount; i++) {
          const cublasStatus_t st =
              ozaki_dgemm(handle, mode, transa, transb, m, n, k, static_cast<const double *>(alpha), a + strideA * i, lda,
                          b + strideB * i, ldb, static_cast<const double *>(beta), c + strideC * i, ldc);
          if (st != CUBLAS_STATUS_SUCCESS) return st;
        }
        return CUBLAS_STATUS_SUCCESS;
      }
      return ozaki_dgemm_batched(handle, mode, transa, transb, m, n, k, static_cast<const double *>(alpha), a, lda,
                                 strideA, b, ldb, strideB, static_cast<const double *>(beta), c, ldc, strideC, batchCount);
    }
  }
  if (mode != dgemm && Atype == CUDA_C_64F && Btype == CUDA_C_64F && Ctype == CUDA_C_64F)
    return cublasZgemmStridedBatched(handle, transa, transb, m, n, k, static_cast<const cuDoubleComplex *>(alpha),
                                     static_cast<const cuDoubleComplex *>(A), lda, strideA,
                                     static_cast<const cuDoubleComplex *>(B), ldb, strideB,
                                     static_cast<const cuDoubleComplex *>(beta), static_cast<cuDoubleComplex *>(C), ldc,
                                     strideC, batchCount);
  auto fn = real_fn<GemmStridedBatchedExFn>("cublasGemmStridedBatchedEx");
  if (fn == nullptr) return CUBLAS_STATUS_NOT_INITIALIZED;
  return fn(handle, transa, transb, m, n, k, alpha, A, Atype, lda, strideA, B, Btype, ldb, strideB, beta, C, Ctype, ldc,
            strideC, batchCount, computeType, algo);
}

// reference src/cublas.cu:474-492
cublasStatus_t cublasDgemmStridedBatched(cublasHandle_t handle, cublasOperation_t transa, cublasOperation_t transb, int m,
                                         int n, int k, const double *alpha, const double *A, int lda, long long strideA,
                                         const double *B, int ldb, long long strideB, const double *beta, double *C,
                                         int ldc, long long strideC, int batchCount) {
  const compute_mode_t mode = env_compute_mode();
  if (mode != dgemm)
    return cublasGemmStridedBatchedEx(handle, transa, transb, m, n, k, alpha, A, CUDA_R_64F, lda, strideA, B, CUDA_R_64F,
                                      ldb, strideB, beta, C, CUDA_R_64F, ldc, strideC, batchCount, CUBLAS_COMPUTE_64F,
                                      CUBLAS_GEMM_DEFAULT);
  auto fn = real_fn<cublasStatus_t (*)(cublasHandle_t, cublasOperation_t, cublasOperation_t, int, int, int, const double *,
                                    const double *, int, long long, const double *, int, long long, const double *,
                                    double *, int, long long, int)>("cublasDgemmStridedBatched");
  if (fn == nullptr) return CUBLAS_STATUS_NOT_INITIALIZED;
  return fn(handle, transa, transb, m, n, k, alpha, A, lda, strideA, B, ldb, strideB, beta, C, ldc, strideC, batchCount);
}

// reference src/cublas.cu:494-512 (-> :315-472 with C_64F: one complex Ozaki GEMM per batch entry, :380-406)
cublasStatus_t cublasZgemmStridedBatched(cublasHandle_t handle, cublasOperation_t transa, cublasOperation_t transb, int m,
                                         int n, int k, const cuDoubleComplex *alpha, const cuDoubleComplex *A, int lda,
                                         long long strideA, const cuDoubleComplex *B, int ldb, long long strideB,
                                         const cuDoubleComplex *beta, cuDoubleComplex *C, int ldc, long long strideC,
                                         int batchCount) {
  const compute_mode_t mode = env_compute_mode();
  if (mode != dgemm && no_conj(transa, transb)) {
    bool take = false;
    {
      std::lock_guard<std::mutex> lock(g_mu);
      try {
        take = should_intercept(global_handle(), mode, m, n, k, CUDA_C_64F, CUDA_C_64F, CUDA_C_64F) &&
               host_pointer_mode(handle);
      } catch (const std::exception &e) {
        H::log_error(e.what());
      }
    }
    if (take) {
      for (int i = 0; i < batchCount; i++) {
        const cublasStatus_t st = ozaki_gemm(handle, mode, transa, transb, m, n, k, alpha, A + strideA * i, lda,
                                             B + strideB * i, ldb, beta, C + strideC * i, ldc, complx);
        if (st != CUBLAS_STATUS_SUCCESS) return st;
      }
      return CUBLAS_STATUS_SUCCESS;
    }
  }
  auto fn = real_fn<cublasStatus_t (*)(cublasHandle_t, cublasOperation_t, cublasOperation_t, int, int, int,
                                    const cuDoubleComplex *, const cuDoubleComplex *, int, long long,
                                    const cuDoubleComplex *, int, long long, const cuDoubleComplex *, cuDoubleComplex *,
                                    int, long long, int)>("cublasZgemmStridedBatched");
  if (fn == nullptr) return CUBLAS_STATUS_NOT_INITIALIZED;
  return fn(handle, transa, transb, m, n, k, alpha, A, lda, strideA, B, ldb, strideB, beta, C, ldc, strideC, batchCount);
}

}  // extern "C"
