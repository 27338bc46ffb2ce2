// host_e2e.cu -- ozimmu_gemm_host: the same DGEMM with HOST operands (the end-to-end entry a
// caller without device buffers uses; bench.py times it as "e2e").
//
// The reference has no such entry: an application does cudaMemcpy(A), cudaMemcpy(B),
// cublasDgemm (intercepted -> reference src/cublas.cu:280-295 -> src/gemm.cu:344-410),
// cudaMemcpy(C), strictly one after the other.  Here the three legs are pipelined over column
// panels of op(B)/C on three streams:
//
//   h2d stream : B panel 0 | A | B panel 1 | B panel 2 | ...
//   compute    :               split(A) split(B0) fused(0) | split(B1) fused(1) | ...
//   d2h stream :                                  C panel 0          | C panel 1 | ...
//
// Results are bit-identical to the one-shot path: a column panel of C depends only on A and the
// same columns of op(B), and the split scales B per column (reference src/split.cu:277-282).
#include <algorithm>
#include <cstdlib>

#include "host.hpp"
#include "oz_common.cuh"
#include "ozimmu_b200.h"

using namespace mtk::ozimmu;
namespace H = oz::host;

namespace {

void ensure_stage(void **ptr, std::size_t *have, std::size_t need) {
  if (need <= *have) return;
  if (*ptr) {
    OZ_CUDA_CHECK(cudaDeviceSynchronize());
    OZ_CUDA_CHECK(cudaFree(*ptr));
    *ptr = nullptr;
    *have = 0;
  }
  OZ_CUDA_CHECK(cudaMalloc(ptr, need));
  *have = need;
}

void ensure_e2e_streams(handle_t h) {
  if (h->h2d_stream) return;
  OZ_CUDA_CHECK(cudaStreamCreateWithFlags(&h->h2d_stream, cudaStreamNonBlocking));
  OZ_CUDA_CHECK(cudaStreamCreateWithFlags(&h->d2h_stream, cudaStreamNonBlocking));
  OZ_CUDA_CHECK(cudaStreamCreateWithFlags(&h->compute_stream, cudaStreamNonBlocking));
  OZ_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_a_in, cudaEventDisableTiming));
  for (int i = 0; i < handle::kMaxPanels; i++) {
    OZ_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_panel_in[i], cudaEventDisableTiming));
    OZ_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_panel_out[i], cudaEventDisableTiming));
  }
}

// column-major ld x cols matrix with `rows` valid rows per column (BLAS guarantees only
// ld*(cols-1)+rows elements)
void copy_matrix(double *dst, const double *src, std::size_t ld, std::size_t rows, std::size_t cols,
                 cudaMemcpyKind kind, cudaStream_t st) {
  if (rows == 0 || cols == 0) return;
  if (ld == rows) {
    OZ_CUDA_CHECK(cudaMemcpyAsync(dst, src, sizeof(double) * ld * cols, kind, st));
  } else {
    OZ_CUDA_CHECK(cudaMemcpy2DAsync(dst, sizeof(double) * ld, src, sizeof(double) * ld, sizeof(double) * rows, cols,
                                    kind, st));
  }
}

int gemm_host_impl(handle_t h, operation_t op_a, operation_t op_b, std::size_t m, std::size_t n, std::size_t k,
                   double alpha, const double *a, std::size_t lda, const double *b, std::size_t ldb, double beta,
                   double *c, std::size_t ldc, compute_mode_t mode) {
  if ((op_a == op_n ? m : k) > lda || (op_b == op_n ? k : n) > ldb || m > ldc) {
    H::log_error("ozimmu_gemm_host: leading dimension smaller than the matrix");
    return 1;
  }
  if (m == 0 || n == 0) return 0;
  ensure_e2e_streams(h);
  H::ensure_streams(h);
  const std::size_t a_rows = (op_a == op_n) ? m : k, a_cols = (op_a == op_n) ? k : m;
  const std::size_t b_rows = (op_b == op_n) ? k : n, b_cols = (op_b == op_n) ? n : k;
  ensure_stage(&h->stage_a, &h->stage_a_bytes, sizeof(double) * lda * std::max<std::size_t>(a_cols, 1));
  ensure_stage(&h->stage_b, &h->stage_b_bytes, sizeof(double) * ldb * std::max<std::size_t>(b_cols, 1));
  ensure_stage(&h->stage_c, &h->stage_c_bytes, sizeof(double) * ldc * n);
  auto *da = static_cast<double *>(h->stage_a);
  auto *db = static_cast<double *>(h->stage_b);
  auto *dc = static_cast<double *>(h->stage_c);
  cudaStream_t sc = h->compute_stream, sin = h->h2d_stream, sout = h->d2h_stream;

  const bool pipelined = H::is_int8_mode(mode) && k > 0 && !h->profiler.enabled;
  if (!pipelined) {
    // one-shot: copy in, run the device entry, copy out
    copy_matrix(da, a, lda, a_rows, a_cols, cudaMemcpyHostToDevice, sc);
    copy_matrix(db, b, ldb, b_rows, b_cols, cudaMemcpyHostToDevice, sc);
    if (beta != 0) copy_matrix(dc, c, ldc, m, n, cudaMemcpyHostToDevice, sc);
    cudaStream_t saved = h->cuda_stream;
    h->cuda_stream = sc;
    int rc = 0;
    try {
      rc = gemm(h, op_a, op_b, m, n, k, &alpha, da, lda, db, ldb, &beta, dc, ldc, mode, real);
    } catch (...) {
      h->cuda_stream = saved;
      throw;
    }
    h->cuda_stream = saved;
    if (rc) return rc;
    copy_matrix(c, dc, ldc, m, n, cudaMemcpyDeviceToHost, sc);
    OZ_CUDA_CHECK(cudaStreamSynchronize(sc));
    return 0;
  }

  const unsigned s = H::num_split_of(mode);
  const unsigned bits = ozk_bits_per_int8(static_cast<std::uint32_t>(k));
  // column panels of C: multiples of 256 columns (the kernel's tile width), at most kMaxPanels, roughly
  // `target` wide (OZIMMU_B200_E2E_PANEL, default 1024)
  static const std::size_t target = [] {
    const char *e = std::getenv("OZIMMU_B200_E2E_PANEL");
    const long v = e ? std::atol(e) : 0;
    return static_cast<std::size_t>(v >= 256 ? v : 1024);
  }();
  std::size_t panels = std::min<std::size_t>(handle::kMaxPanels, std::max<std::size_t>(1, n / target));
  std::size_t pw = ((n + panels - 1) / panels + 255) / 256 * 256;
  panels = (n + pw - 1) / pw;

  // workspace: A slices for all of A, B slices for one panel
  const H::WorkspaceLayout w = H::workspace_layout(m, pw, k, s);
  reallocate_working_memory(h, w.total);
  if (h->has_pending) OZ_CUDA_CHECK(cudaStreamWaitEvent(sc, h->ev_done, 0));
  char *ws = static_cast<char *>(h->working_memory_ptr);
  double *amax = reinterpret_cast<double *>(ws + w.off_amax);
  double *bmax = reinterpret_cast<double *>(ws + w.off_bmax);
  auto *scr_a = reinterpret_cast<std::uint32_t *>(ws + w.off_scr_a);
  auto *scr_b = reinterpret_cast<std::uint32_t *>(ws + w.off_scr_b);
  auto *a_sl = reinterpret_cast<std::int8_t *>(ws + w.off_a_slices);
  auto *b_sl = reinterpret_cast<std::int8_t *>(ws + w.off_b_slices);

  auto copy_b_panel = [&](std::size_t p) {
    const std::size_t j0 = p * pw, nj = std::min(pw, n - j0);
    if (op_b == op_n) {  // k x n column-major: a column panel is contiguous
      copy_matrix(db + j0 * ldb, b + j0 * ldb, ldb, k, nj, cudaMemcpyHostToDevice, sin);
    } else {             // n x k column-major: rows j0..j0+nj of every column
      OZ_CUDA_CHECK(cudaMemcpy2DAsync(db + j0, sizeof(double) * ldb, b + j0, sizeof(double) * ldb,
                                      sizeof(double) * nj, k, cudaMemcpyHostToDevice, sin));
    }
    if (beta != 0) copy_matrix(dc + j0 * ldc, c + j0 * ldc, ldc, m, nj, cudaMemcpyHostToDevice, sin);
    OZ_CUDA_CHECK(cudaEventRecord(h->ev_panel_in[p], sin));
  };

  copy_b_panel(0);
  copy_matrix(da, a, lda, a_rows, a_cols, cudaMemcpyHostToDevice, sin);
  OZ_CUDA_CHECK(cudaEventRecord(h->ev_a_in, sin));
  for (std::size_t p = 1; p < panels; p++) copy_b_panel(p);

  OZ_CUDA_CHECK(cudaStreamWaitEvent(sc, h->ev_a_in, 0));
  OZ_KERNEL_CHECK(ozk_split_int8(a_sl, w.pitch, amax, scr_a, m, k, da, lda, op_a == op_n, s, bits, sc));
  for (std::size_t p = 0; p < panels; p++) {
    const std::size_t j0 = p * pw, nj = std::min(pw, n - j0);
    OZ_CUDA_CHECK(cudaStreamWaitEvent(sc, h->ev_panel_in[p], 0));
    const double *bp = (op_b == op_n) ? db + j0 * ldb : db + j0;
    OZ_KERNEL_CHECK(ozk_split_int8(b_sl, w.pitch, bmax, scr_b, nj, k, bp, ldb, op_b != op_n, s, bits, sc));
    OZ_KERNEL_CHECK(ozk_gemm_i8_fused(m, nj, k, a_sl, b_sl, w.pitch, amax, bmax, s, bits, alpha, beta,
                                      dc + j0 * ldc, ldc, sc));
    OZ_CUDA_CHECK(cudaEventRecord(h->ev_panel_out[p], sc));
    OZ_CUDA_CHECK(cudaStreamWaitEvent(sout, h->ev_panel_out[p], 0));
    copy_matrix(c + j0 * ldc, dc + j0 * ldc, ldc, m, nj, cudaMemcpyDeviceToHost, sout);
  }
  OZ_CUDA_CHECK(cudaEventRecord(h->ev_done, sc));
  h->has_pending = true;
  h->last_stream = sc;
  OZ_CUDA_CHECK(cudaStreamSynchronize(sout));
  OZ_CUDA_CHECK(cudaStreamSynchronize(sc));
  return 0;
}

}  // namespace

extern "C" int ozimmu_gemm_host(ozimmu_handle_t handle, int op_a, int op_b, size_t m, size_t n, size_t k,
                                const double *alpha, const double *a, size_t lda, const double *b, size_t ldb,
                                const double *beta, double *c, size_t ldc, int compute_mode) {
  if (handle == nullptr || alpha == nullptr || beta == nullptr || compute_mode < 0 ||
      compute_mode > OZIMMU_FP64_INT8_AUTO)
    return 1;
  try {
    return gemm_host_impl(reinterpret_cast<handle_t>(handle), static_cast<operation_t>(op_a != 0),
                          static_cast<operation_t>(op_b != 0), m, n, k, *alpha, a, lda, b, ldb, *beta, c, ldc,
                          static_cast<compute_mode_t>(compute_mode));
  } catch (const std::exception &e) {
    H::log_error(e.what());
    return -1;
  }
}
