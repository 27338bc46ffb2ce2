// ref_shim.cu -- C-ABI doorway into the UNMODIFIED reference ozIMMU (test infrastructure).
//
// oracle/Makefile compiles this file together with the reference's own sources, taken
// where they lie under /root/reference (never copied into this repo), into
// oracle/_ref/libozref.so.  Only `ozref_*` symbols are exported (oracle/ref_exports.map),
// so the reference's cuBLAS interposers stay private to that library and cannot shadow
// anything in the test process.
//
// Used by: tests/ (bit-for-bit parity of the product against the reference on the GPU),
// tests/golden/make_golden.py (golden vectors that pin oracle/oz_oracle.c) and
// bench.py --impl reference.  Never loaded by the product.
#include <cstdint>
#include <cstdio>
#include <exception>

#include "split.hpp"   // reference src/split.hpp: split_int8<T>, get_mantissa_loss_total<T>
#include "cublas_helper.hpp"   // reference src/cublas_helper.hpp: dgemm_f32<T> (compute mode `sgemm`)
#include <ozimmu/ozimmu.hpp>

#define OZREF_API extern "C" __attribute__((visibility("default")))

namespace {
template <class F> int guarded(F &&f) {
  try {
    return f();
  } catch (const std::exception &e) {
    std::fprintf(stderr, "[ozref] exception: %s\n", e.what());
    return -1;
  }
}
} // namespace

// include/ozimmu/ozimmu.hpp:47-49
OZREF_API int ozref_create(void **handle) {
  return guarded([&] { return mtk::ozimmu::create(reinterpret_cast<mtk::ozimmu::handle_t *>(handle)); });
}
OZREF_API int ozref_destroy(void *handle) {
  return guarded([&] { return mtk::ozimmu::destroy(static_cast<mtk::ozimmu::handle_t>(handle)); });
}
OZREF_API int ozref_set_stream(void *handle, void *stream) {
  return guarded([&] {
    mtk::ozimmu::set_cuda_stream(static_cast<mtk::ozimmu::handle_t>(handle),
                                 static_cast<cudaStream_t>(stream));
    return 0;
  });
}

// include/ozimmu/ozimmu.hpp:76-83; op: 0 = N, 1 = T; mode = compute_mode_t value; kind 0 = real
OZREF_API int ozref_gemm(void *handle, int op_a, int op_b, size_t m, size_t n, size_t k,
                         const void *alpha, const void *a, size_t lda, const void *b, size_t ldb,
                         const void *beta, void *c, size_t ldc, int mode, int kind) {
  return guarded([&] {
    return mtk::ozimmu::gemm(static_cast<mtk::ozimmu::handle_t>(handle),
                             static_cast<mtk::ozimmu::operation_t>(op_a),
                             static_cast<mtk::ozimmu::operation_t>(op_b), m, n, k, alpha, a, lda, b,
                             ldb, beta, c, ldc, static_cast<mtk::ozimmu::compute_mode_t>(mode),
                             static_cast<mtk::ozimmu::element_kind_t>(kind));
  });
}

// src/split.cu:266-283.  matrix: 0 = A (m x n view = rows x len), 1 = B (reference swaps).
OZREF_API int ozref_split_int8(int8_t *out, uint32_t ldo, double *max_exp, size_t m, size_t n,
                               const double *in, size_t ld, int op, int matrix, unsigned num_split,
                               unsigned bits_per_int8, void *stream) {
  return guarded([&] {
    mtk::ozimmu::split_int8<double>(out, ldo, max_exp, m, n, in, ld,
                                    static_cast<mtk::ozimmu::operation_t>(op),
                                    static_cast<mtk::ozimmu::detail::matrix_t>(matrix), num_split,
                                    bits_per_int8, static_cast<cudaStream_t>(stream));
    return 0;
  });
}

// include/ozimmu/ozimmu.hpp:85-94.  Returns the compute_mode_t value (or -1 on exception).
// counters8 (host, optional) receives the 8 device counters the reference actually owns
// (fp64_int8_3..10; src/handle.hpp:22 -- SURVEY App. B.1).
OZREF_API int ozref_auto_mode_select(void *handle, int op_a, int op_b, size_t m, size_t n, size_t k,
                                     const void *a, size_t lda, const void *b, size_t ldb, int kind,
                                     double threshold, unsigned long long *counters8) {
  return guarded([&] {
    auto h = static_cast<mtk::ozimmu::handle_t>(handle);
    const auto mode = mtk::ozimmu::auto_mode_select(
        h, static_cast<mtk::ozimmu::operation_t>(op_a), static_cast<mtk::ozimmu::operation_t>(op_b), m,
        n, k, a, lda, b, ldb, static_cast<mtk::ozimmu::element_kind_t>(kind), threshold);
    if (counters8) {
      cudaMemcpy(counters8, h->d_mantissa_loss_counter_ptr,
                 sizeof(unsigned long long) * mtk::ozimmu::handle::mantissa_loss_counter_length,
                 cudaMemcpyDefault);
    }
    return static_cast<int>(mode);
  });
}

OZREF_API unsigned ozref_bits_per_int8(unsigned k) { return mtk::ozimmu::get_bits_per_int8(k); }

// include/ozimmu/ozimmu.hpp:69-74
OZREF_API size_t ozref_reallocate(void *handle, int op_a, int op_b, size_t m, size_t n, size_t k,
                                  int kind, int mode) {
  size_t r = 0;
  guarded([&] {
    r = mtk::ozimmu::reallocate_working_memory(
        static_cast<mtk::ozimmu::handle_t>(handle),
        mtk::ozimmu::gemm_list_t{{static_cast<mtk::ozimmu::operation_t>(op_a),
                                  static_cast<mtk::ozimmu::operation_t>(op_b), m, n, k,
                                  static_cast<mtk::ozimmu::element_kind_t>(kind),
                                  static_cast<mtk::ozimmu::compute_mode_t>(mode)}});
    return 0;
  });
  return r;
}

// Per-stage profiler of the reference (include/ozimmu/ozimmu.hpp:52-56)
OZREF_API void ozref_profiling(void *handle, int enable) {
  auto h = static_cast<mtk::ozimmu::handle_t>(handle);
  if (enable) mtk::ozimmu::enable_profiling(h); else mtk::ozimmu::disable_profiling(h);
}
OZREF_API void ozref_print_profile(void *handle, const char *tag) {
  auto h = static_cast<mtk::ozimmu::handle_t>(handle);
  mtk::ozimmu::print_profiler_result(h, tag, true);
  mtk::ozimmu::clear_profiler_result(h);
}

// Compute mode `sgemm`: the reference only reaches it from its cuBLAS interposers (src/cublas.cu:169-186), which
// call dgemm_f32<double> (src/cublas_helper.cu:84-134); mtk::ozimmu::gemm(..., sgemm) throws "Not implemented".
OZREF_API int ozref_dgemm_f32(void *handle, int op_a, int op_b, size_t m, size_t n, size_t k, double alpha,
                              const double *a, size_t lda, const double *b, size_t ldb, double beta, double *c,
                              size_t ldc) {
  return guarded([&] {
    const cublasStatus_t st = mtk::ozimmu::dgemm_f32<double>(
        static_cast<mtk::ozimmu::handle_t>(handle), op_a ? CUBLAS_OP_T : CUBLAS_OP_N, op_b ? CUBLAS_OP_T : CUBLAS_OP_N, m,
        n, k, alpha, a, lda, b, ldb, beta, c, ldc);
    return st == CUBLAS_STATUS_SUCCESS ? 0 : 1;
  });
}
