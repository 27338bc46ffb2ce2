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
// scalars are read on the device by the Ozaki kernels (the reference dereferences them on the host, src/gemm.cu:405).
// State is per device: one ozIMMU handle (workspace, streams, events) and one lock per GPU, created on first use
// with that GPU current, so a process that drives several GPUs -- one cuBLAS handle each, from one or several
// threads -- gets the right workspace on each and calls on different GPUs do not serialise (the reference keeps ONE
// global handle, src/cublas.cu:58-86).
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

// per-device state, indexed by the current CUDA device (the device the application's cuBLAS handle lives on: cuBLAS
// requires it to be current for every call on that handle)
constexpr int kMaxDevices = 64;
struct DeviceState {
  std::mutex mu;
  handle_t handle = nullptr;
};
DeviceState g_state[kMaxDevices];

DeviceState &device_state() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices)
    throw std::runtime_error("ozIMMU: no current CUDA device");
  return g_state[dev];
}

// reference src/cublas.cu:18-48: unknown / unset -> dgemm (passthrough)
compute_mode_t env_compute_mode() {
  const char *v = std::getenv("OZIMMU_COMPUTE_MODE");
  if (v == nullptr) return dgemm;
  for (int mode = sgemm; mode <= fp64_int8_auto; mode++)
    if (get_compute_mode_name_str(static_cast<compute_mode_t>(mode)) == v) return static_cast<compute_mode_t>(mode);
  return dgemm;
}

// reference src/cublas.cu:60-86; call with st.mu held
handle_t device_handle(DeviceState &st) {
  if (st.handle == nullptr) {
    const malloc_mode_t mm = H::env_enabled("OZIMMU_MALLOC_ASYNC", false) ? malloc_async : malloc_sync;
    H::log_info("Initializing ozIMMU handle...");
    create(&st.handle, mm);
    H::log_info("Successfully initialized");
  }
  if (const char *t = std::getenv("OZIMMU_AUTO_AVG_MANTISSA_LOSS_THRESHOLD")) {
    char *end = nullptr;
    const double v = std::strtod(t, &end);
    if (end == t) throw std::runtime_error(std::string("ERROR: invalid OZIMMU_AUTO_AVG_MANTISSA_LOSS_THRESHOLD = ") + t);
    set_auto_mantissa_loss_threashold(st.handle, v);
  }
  return st.handle;
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
  cudaStream_t s = nullptr;
  std::string name;
  std::chrono::steady_clock::time_point t0;
  // make_name is only evaluated when profiling is on (it formats several numbers; every passthrough GEMM of a
  // preloaded application comes through here)
  template <class MakeName, class GetStream>
  CulipScope(GetStream &&get_stream, MakeName &&make_name) : on(H::env_enabled("OZIMMU_ENABLE_CULIP_PROFILING", false)) {
    if (!on) return;
    s = get_stream();
    name = make_name();
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

std::string shape_str(cublasOperation_t ta, cublasOperation_t tb, int m, int n, int k) {
  return std::string(op_str(ta)) + op_str(tb) + "-m" + std::to_string(m) + "-n" + std::to_string(n) + "-k" + std::to_string(k);
}

// true if this call goes through the Ozaki path (thresholds of the current device's handle)
bool take_call(compute_mode_t mode, int m, int n, int k, cudaDataType_t a, cudaDataType_t b, cudaDataType_t c) {
  try {
    DeviceState &st = device_state();
    std::lock_guard<std::mutex> lock(st.mu);
    return should_intercept(device_handle(st), mode, m, n, k, a, b, c);
  } catch (const std::exception &e) {
    H::log_error(e.what());
    return false;
  }
}

// one GEMM (batch == 1) or a strided batch through the Ozaki path, real or complex; strides in elements
cublasStatus_t ozaki_gemm(cublasHandle_t handle, compute_mode_t mode, cublasOperation_t ta, cublasOperation_t tb, int m,
                          int n, int k, const void *alpha, const void *A, int lda, long long strideA, const void *B,
                          int ldb, long long strideB, const void *beta, void *C, int ldc, long long strideC, int batch,
                          element_kind_t kind) {
  try {
    DeviceState &st = device_state();
    std::lock_guard<std::mutex> lock(st.mu);
    handle_t h = device_handle(st);
    cudaStream_t s = stream_of(handle);
    set_cuda_stream(h, s);
    // cuBLAS device pointer mode: the scalars stay on the device (reference src/gemm.cu:405 dereferences them on the host)
    set_scalar_pointer_mode(h, !host_pointer_mode(handle));
    CulipScope scope([&] { return s; }, [&] {
      return std::string(kind == mtk::ozimmu::real ? "D" : "Z") + get_compute_mode_name_str(mode) +
             (batch > 1 ? "-batched" + std::to_string(batch) : std::string()) + "-" + shape_str(ta, tb, m, n, k);
    });
    // reference src/cublas.cu:50-56: everything that is not OP_N is treated as OP_T (real data)
    const operation_t oa = ta == CUBLAS_OP_N ? op_n : op_t, ob = tb == CUBLAS_OP_N ? op_n : op_t;
    const int err = batch == 1 ? gemm(h, oa, ob, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, mode, kind)
                               : gemm_strided_batched(h, oa, ob, m, n, k, alpha, A, lda, strideA, B, ldb, strideB, beta, C,
                                                      ldc, strideC, static_cast<std::size_t>(batch), mode, kind);
    set_scalar_pointer_mode(h, false);
    return err ? CUBLAS_STATUS_INVALID_VALUE : CUBLAS_STATUS_SUCCESS;
  } catch (const std::exception &e) {
    H::log_error(e.what());
    return CUBLAS_STATUS_INTERNAL_ERROR;
  }
}

// a strided batch: one grouped launch (the reference loops, src/cublas.cu:380-406) unless entries of C overlap or walk
// backwards -- those cannot run concurrently and go entry by entry, as the reference does
cublasStatus_t ozaki_batch(cublasHandle_t handle, compute_mode_t mode, cublasOperation_t ta, cublasOperation_t tb, int m,
                           int n, int k, const void *alpha, const void *A, int lda, long long strideA, const void *B,
                           int ldb, long long strideB, const void *beta, void *C, int ldc, long long strideC, int batch,
                           element_kind_t kind) {
  const std::size_t esz = kind == mtk::ozimmu::real ? sizeof(double) : 2 * sizeof(double);
  const unsigned long long span = static_cast<unsigned long long>(ldc) * (n > 0 ? n - 1 : 0) + m;
  if (batch == 1 || strideC < 0 || static_cast<unsigned long long>(strideC) < span) {
    for (int i = 0; i < batch; i++) {
      const cublasStatus_t st =
          ozaki_gemm(handle, mode, ta, tb, m, n, k, alpha, static_cast<const char *>(A) + strideA * i * static_cast<long long>(esz),
                     lda, 0, static_cast<const char *>(B) + strideB * i * static_cast<long long>(esz), ldb, 0, beta,
                     static_cast<char *>(C) + strideC * i * static_cast<long long>(esz), ldc, 0, 1, kind);
      if (st != CUBLAS_STATUS_SUCCESS) return st;
    }
    return CUBLAS_STATUS_SUCCESS;
  }
  return ozaki_gemm(handle, mode, ta, tb, m, n, k, alpha, A, lda, strideA, B, ldb, strideB, beta, C, ldc, strideC, batch, kind);
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
    try {
      // pre-size this device's workspace for one 1024^3 fp64_int8_9 product (reference src/cublas.cu:12-16)
      DeviceState &ds = device_state();
      std::lock_guard<std::mutex> lock(ds.mu);
      reallocate_working_memory(device_handle(ds), gemm_list_t{{op_n, op_n, 1024, 1024, 1024, mtk::ozimmu::real, fp64_int8_9}});
    } catch (const std::exception &e) {
      H::log_error(e.what());
    }
  }
  return st;
}

// reference src/cublas.cu:117-131.  The reference tears its global handle down on ANY
// cublasDestroy; here the per-device handles live until the process ends (another cuBLAS handle may still be in use).
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
  if (mode != dgemm && (is_real || is_cplx) && take_call(mode, m, n, k, Atype, Btype, Ctype))
    return ozaki_gemm(handle, mode, transa, transb, m, n, k, alpha, A, lda, 0, B, ldb, 0, beta, C, ldc, 0, 1,
                      is_real ? mtk::ozimmu::real : complx);
  auto fn = real_fn<GemmExFn>("cublasGemmEx");
  if (fn == nullptr) return CUBLAS_STATUS_NOT_INITIALIZED;
  CulipScope scope([&] { return stream_of(handle); }, [&] { return "cublasGemmEx-" + shape_str(transa, transb, m, n, k); });
  return fn(handle, transa, transb, m, n, k, alpha, A, Atype, lda, B, Btype, ldb, beta, C, Ctype, ldc, computeType, algo);
}

// reference src/cublas.cu:280-295
cublasStatus_t cublasDgemm_v2(cublasHandle_t handle, cublasOperation_t transa, cublasOperation_t transb, int m, int n,
                              int k, const double *alpha, const double *A, int lda, const double *B, int ldb,
                              const double *beta, double *C, int ldc) {
  const compute_mode_t mode = env_compute_mode();
  if (mode != dgemm && take_call(mode, m, n, k, CUDA_R_64F, CUDA_R_64F, CUDA_R_64F))
    return ozaki_gemm(handle, mode, transa, transb, m, n, k, alpha, A, lda, 0, B, ldb, 0, beta, C, ldc, 0, 1,
                      mtk::ozimmu::real);
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
  if (mode != dgemm && no_conj(transa, transb) && take_call(mode, m, n, k, CUDA_C_64F, CUDA_C_64F, CUDA_C_64F))
    return ozaki_gemm(handle, mode, transa, transb, m, n, k, alpha, A, lda, 0, B, ldb, 0, beta, C, ldc, 0, 1, complx);
  auto fn = real_fn<cublasStatus_t (*)(cublasHandle_t, cublasOperation_t, cublasOperation_t, int, int, int,
                                    const cuDoubleComplex *, const cuDoubleComplex *, int, const cuDoubleComplex *, int,
                                    const cuDoubleComplex *, cuDoubleComplex *, int)>("cublasZgemm_v2");
  if (fn == nullptr) return CUBLAS_STATUS_NOT_INITIALIZED;
  return fn(handle, transa, transb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
}

// reference src/cublas.cu:315-472 (one Ozaki GEMM per batch entry, :380-406): here one grouped launch, real or complex
cublasStatus_t cublasGemmStridedBatchedEx(cublasHandle_t handle, cublasOperation_t transa, cublasOperation_t transb, int m,
                                          int n, int k, const void *alpha, const void *A, cudaDataType_t Atype, int lda,
                                          long long strideA, const void *B, cudaDataType_t Btype, int ldb,
                                          long long strideB, const void *beta, void *C, cudaDataType_t Ctype, int ldc,
                                          long long strideC, int batchCount, cublasComputeType_t computeType,
                                          cublasGemmAlgo_t algo) {
  const compute_mode_t mode = env_compute_mode();
  const bool is_real = Atype == CUDA_R_64F && Btype == CUDA_R_64F && Ctype == CUDA_R_64F;
  const bool is_cplx = Atype == CUDA_C_64F && Btype == CUDA_C_64F && Ctype == CUDA_C_64F && no_conj(transa, transb);
  if (mode != dgemm && (is_real || is_cplx) && batchCount > 0 && take_call(mode, m, n, k, Atype, Btype, Ctype))
    return ozaki_batch(handle, mode, transa, transb, m, n, k, alpha, A, lda, strideA, B, ldb, strideB, beta, C, ldc, strideC,
                       batchCount, is_real ? mtk::ozimmu::real : complx);
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
  if (mode != dgemm && batchCount > 0 && take_call(mode, m, n, k, CUDA_R_64F, CUDA_R_64F, CUDA_R_64F))
    return ozaki_batch(handle, mode, transa, transb, m, n, k, alpha, A, lda, strideA, B, ldb, strideB, beta, C, ldc, strideC,
                       batchCount, mtk::ozimmu::real);
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
  if (mode != dgemm && no_conj(transa, transb) && batchCount > 0 &&
      take_call(mode, m, n, k, CUDA_C_64F, CUDA_C_64F, CUDA_C_64F))
    return ozaki_batch(handle, mode, transa, transb, m, n, k, alpha, A, lda, strideA, B, ldb, strideB, beta, C, ldc, strideC,
                       batchCount, complx);
  auto fn = real_fn<cublasStatus_t (*)(cublasHandle_t, cublasOperation_t, cublasOperation_t, int, int, int,
                                    const cuDoubleComplex *, const cuDoubleComplex *, int, long long,
                                    const cuDoubleComplex *, int, long long, const cuDoubleComplex *, cuDoubleComplex *,
                                    int, long long, int)>("cublasZgemmStridedBatched");
  if (fn == nullptr) return CUBLAS_STATUS_NOT_INITIALIZED;
  return fn(handle, transa, transb, m, n, k, alpha, A, lda, strideA, B, ldb, strideB, beta, C, ldc, strideC, batchCount);
}

}  // extern "C"
