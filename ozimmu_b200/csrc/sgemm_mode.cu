// sgemm_mode.cu -- compute mode `sgemm`: the FP64 (or complex FP64) GEMM is carried out in FP32 by cuBLAS.
// Follows reference src/cublas_helper.cu:20-134 (convert_dtype_kernel, dgemm_f32): A, B (and C when beta != 0)
// are converted to tightly packed FP32 copies in the workspace, cublasSgemm / cublasCgemm runs on them, and the
// FP32 result is widened back into C.  Not an Ozaki path; here for users of the reference's
// OZIMMU_COMPUTE_MODE=sgemm.  HBM-bound conversions: one thread per 2 elements, coalesced along columns.
#include <cuComplex.h>

#include "host.hpp"
#include "oz_common.cuh"
#include "ozimmu_b200.h"

using namespace mtk::ozimmu;
namespace H = oz::host;

namespace {

// rows x cols column-major, element (r, c) at src[c * lds + r] -> dst[c * ldd + r]; rows counts scalars (a
// complex matrix is passed as 2 * rows interleaved scalars per column)
template <class Dst, class Src>
__global__ void __launch_bounds__(256)
convert_matrix_kernel(Dst *__restrict__ dst, const size_t ldd, const Src *__restrict__ src, const size_t lds,
                      const size_t rows, const size_t cols) {
  const size_t c = blockIdx.y;
  for (size_t r = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x; r < rows; r += static_cast<size_t>(gridDim.x) * 256)
    dst[c * ldd + r] = static_cast<Dst>(src[c * lds + r]);
}

template <class Dst, class Src>
void convert_matrix(Dst *dst, size_t ldd, const Src *src, size_t lds, size_t rows, size_t cols, cudaStream_t s) {
  if (rows == 0 || cols == 0) return;
  const unsigned gx = static_cast<unsigned>(std::min<size_t>((rows + 255) / 256, 64));
  for (size_t c0 = 0; c0 < cols; c0 += 65535) {
    const unsigned gy = static_cast<unsigned>(std::min<size_t>(cols - c0, 65535));
    convert_matrix_kernel<Dst, Src><<<dim3(gx, gy), 256, 0, s>>>(dst + c0 * ldd, ldd, src + c0 * lds, lds, rows, gy);
    oz::count_launch(1);
  }
  OZ_CUDA_CHECK(cudaGetLastError());
}

}  // namespace

// reference src/cublas_helper.cu:84-134 dgemm_f32<T>, T = double (element_kind real) / cuDoubleComplex
void oz::host::gemm_in_f32(handle_t h, cublasHandle_t cublas, operation_t op_a, operation_t op_b, std::size_t m,
                           std::size_t n, std::size_t k, const double *alpha, const double *a, std::size_t lda,
                           const double *b, std::size_t ldb, const double *beta, double *c, std::size_t ldc,
                           element_kind_t kind) {
  if (m == 0 || n == 0) return;
  const std::size_t es = kind == real ? 1 : 2;  // scalars per element
  const std::size_t a_rows = op_a == op_n ? m : k, a_cols = op_a == op_n ? k : m;
  const std::size_t b_rows = op_b == op_n ? k : n, b_cols = op_b == op_n ? n : k;
  const std::size_t bytes = (m * k + k * n + m * n) * es * sizeof(float);
  reallocate_working_memory(h, bytes);
  cudaStream_t s = h->cuda_stream;
  if (h->has_pending && h->last_stream != s) OZ_CUDA_CHECK(cudaStreamWaitEvent(s, h->ev_done, 0));
  auto *af = static_cast<float *>(h->working_memory_ptr);
  float *bf = af + m * k * es, *cf = bf + k * n * es;
  convert_matrix(af, a_rows * es, a, lda * es, a_rows * es, a_cols, s);
  convert_matrix(bf, b_rows * es, b, ldb * es, b_rows * es, b_cols, s);
  const bool beta_zero = beta[0] == 0 && (kind == real || beta[1] == 0);
  if (!beta_zero) convert_matrix(cf, m * es, c, ldc * es, m * es, n, s);
  const cublasOperation_t ta = op_a == op_n ? CUBLAS_OP_N : CUBLAS_OP_T, tb = op_b == op_n ? CUBLAS_OP_N : CUBLAS_OP_T;
  cublasStatus_t st;
  if (kind == real) {
    using Fn = cublasStatus_t (*)(cublasHandle_t, cublasOperation_t, cublasOperation_t, int, int, int, const float *,
                                  const float *, int, const float *, int, const float *, float *, int);
    auto fn = reinterpret_cast<Fn>(H::real_cublas_symbol("cublasSgemm_v2"));
    if (fn == nullptr) throw std::runtime_error("ozIMMU: cublasSgemm is not available for the sgemm mode");
    const float al = static_cast<float>(alpha[0]), be = static_cast<float>(beta[0]);
    st = fn(cublas, ta, tb, static_cast<int>(m), static_cast<int>(n), static_cast<int>(k), &al, af,
            static_cast<int>(a_rows), bf, static_cast<int>(b_rows), &be, cf, static_cast<int>(m));
  } else {
    using Fn = cublasStatus_t (*)(cublasHandle_t, cublasOperation_t, cublasOperation_t, int, int, int, const cuComplex *,
                                  const cuComplex *, int, const cuComplex *, int, const cuComplex *, cuComplex *, int);
    auto fn = reinterpret_cast<Fn>(H::real_cublas_symbol("cublasCgemm_v2"));
    if (fn == nullptr) throw std::runtime_error("ozIMMU: cublasCgemm is not available for the sgemm mode");
    const cuComplex al = make_cuComplex(static_cast<float>(alpha[0]), static_cast<float>(alpha[1]));
    const cuComplex be = make_cuComplex(static_cast<float>(beta[0]), static_cast<float>(beta[1]));
    st = fn(cublas, ta, tb, static_cast<int>(m), static_cast<int>(n), static_cast<int>(k), &al,
            reinterpret_cast<const cuComplex *>(af), static_cast<int>(a_rows), reinterpret_cast<const cuComplex *>(bf),
            static_cast<int>(b_rows), &be, reinterpret_cast<cuComplex *>(cf), static_cast<int>(m));
  }
  if (st != CUBLAS_STATUS_SUCCESS)
    throw std::runtime_error("ozIMMU: FP32 GEMM of the sgemm mode failed with cuBLAS status " + std::to_string(st));
  convert_matrix(c, ldc * es, cf, m * es, m * es, n, s);
  OZ_CUDA_CHECK(cudaEventRecord(h->ev_done, s));
  h->has_pending = true;
  h->last_stream = s;
}
