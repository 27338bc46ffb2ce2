// ozimmu.hpp -- C++ interface of the B200-native Ozaki-scheme DGEMM library.
//
// Source-compatible with the public header of enp1s0/ozIMMU (reference
// include/ozimmu/ozimmu.hpp:8-102): same namespace, enumerator values, type aliases and
// function signatures, so a program written against the reference (e.g. its test driver,
// reference test/main_test.cu:170-172,242-251,257) rebuilds against this library unchanged.
// The plain-C spelling of the same interface lives in include/ozimmu_b200.h.
//
// Behavioural notes (deviations are fixes of reference defects, see DESIGN.md "Deviations"):
//   - gemm() is asynchronous on the handle's stream: no device-wide synchronisation.
//   - get_auto_mantissa_loss_threashold is actually defined (reference src/handle.cu:272
//     defines it outside the namespace, leaving the declared symbol undefined).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <tuple>
#include <vector>

#include <cuda_runtime_api.h>

namespace mtk {
namespace ozimmu {

struct handle;
using handle_t = handle *;

enum operation_t { op_n, op_t };

// sgemm = 0, dgemm = 1, fp64_int8_S = S - 1 (S = 3..18), fp64_int8_auto = 18
#define OZIMMU_B200_INT8_MODES(X) \
  X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16) X(17) X(18)
enum compute_mode_t {
  sgemm,
  dgemm,
#define OZIMMU_B200_ENUM(S) fp64_int8_##S,
  OZIMMU_B200_INT8_MODES(OZIMMU_B200_ENUM)
#undef OZIMMU_B200_ENUM
  fp64_int8_auto,
};

enum data_t { fp64, fp32, fp16, int8, original, none };
enum malloc_mode_t { malloc_sync, malloc_async };
enum element_kind_t { real, complx };

// --- lifetime / stream ---------------------------------------------------------------------
int create(handle_t *handle, const malloc_mode_t mm = malloc_sync);
int destroy(handle_t handle);
void set_cuda_stream(handle_t handle, const cudaStream_t cuda_stream);

// --- per-stage profiler ----------------------------------------------------------------------
void enable_profiling(handle_t handle);
void disable_profiling(handle_t handle);
void print_profiler_result(handle_t handle, const std::string tag, const bool csv = false);
void clear_profiler_result(handle_t handle);

// --- auto mode threshold (spelling as in the reference) -------------------------------------
void set_auto_mantissa_loss_threashold(handle_t handle, const double threshold);
double get_auto_mantissa_loss_threashold(handle_t handle);

// --- workspace -------------------------------------------------------------------------------
// (op_A, op_B, m, n, k, element kind, compute mode)
using gemm_params_t =
    std::tuple<operation_t, operation_t, std::size_t, std::size_t, std::size_t, element_kind_t, compute_mode_t>;
using gemm_list_t = std::vector<gemm_params_t>;

// Both return the new workspace size in bytes if it had to grow, otherwise 0.
std::size_t reallocate_working_memory(handle_t handle, const gemm_list_t gemm_list);
std::size_t reallocate_working_memory(handle_t handle, const std::size_t size_in_byte);

// --- the GEMM ----------------------------------------------------------------------------------
// C = alpha * op(A) * op(B) + beta * C, BLAS column-major.  alpha/beta: host pointers;
// a_ptr/b_ptr/c_ptr: device pointers.  Returns 0, or 1 for an invalid argument.
int gemm(handle_t handle, const operation_t op_A, const operation_t op_B, const std::size_t m,
         const std::size_t n, const std::size_t k, const void *alpha, const void *const a_ptr,
         const std::size_t lda, const void *const b_ptr, const std::size_t ldb, const void *beta,
         void *const c_ptr, std::size_t ldc, const compute_mode_t compute_mode,
         const element_kind_t element_kind);

// Extension (not in the reference header, whose strided-batched interposers loop over gemm(),
// reference src/cublas.cu:380-406): a strided batch of real DGEMMs, C_e = alpha*op(A_e)*op(B_e) + beta*C_e with
// X_e = X + e*stride_x (strides in elements), multiplied by one grouped launch.  Each entry is bit-identical to
// a separate gemm() call.  Returns 0, or 1 for an invalid argument.
int gemm_strided_batched(handle_t handle, const operation_t op_A, const operation_t op_B, const std::size_t m,
                         const std::size_t n, const std::size_t k, const double *alpha, const double *const a_ptr,
                         const std::size_t lda, const long long stride_a, const double *const b_ptr,
                         const std::size_t ldb, const long long stride_b, const double *beta, double *const c_ptr,
                         const std::size_t ldc, const long long stride_c, const std::size_t batch_count,
                         const compute_mode_t compute_mode);

// The same for real or complex data: alpha / beta as for gemm(), strides in (complex) elements.  fp64_int8_auto picks
// the split count per entry from one counter pass over the whole batch.
int gemm_strided_batched(handle_t handle, const operation_t op_A, const operation_t op_B, const std::size_t m,
                         const std::size_t n, const std::size_t k, const void *alpha, const void *const a_ptr,
                         const std::size_t lda, const long long stride_a, const void *const b_ptr,
                         const std::size_t ldb, const long long stride_b, const void *beta, void *const c_ptr,
                         const std::size_t ldc, const long long stride_c, const std::size_t batch_count,
                         const compute_mode_t compute_mode, const element_kind_t element_kind);

// Extension: alpha / beta of the following gemm() / gemm_strided_batched() calls are DEVICE pointers (cuBLAS device
// pointer mode).  The fp64_int8_S modes read them on the device in stream order; the other modes fetch them with one
// blocking read.  The reference always dereferences them on the host (src/gemm.cu:405).
void set_scalar_pointer_mode(handle_t handle, const bool on_device);

// Extension: the same real GEMM when B arrives column panel by column panel (e.g. as a broadcast from another
// GPU delivers it): panel p = columns [col_edges[p], col_edges[p+1]) of op(B) and C (inner edges multiples of 256,
// col_edges[0] = 0, col_edges[num_panels] = n, at most 16 panels) may be read once ready[p] -- an event the caller
// records on any stream -- has fired.  Bit-identical to gemm().  Asynchronous on the handle's stream.
int gemm_streamed_b(handle_t handle, const operation_t op_A, const operation_t op_B, const std::size_t m,
                    const std::size_t n, const std::size_t k, const double *alpha, const double *const a_ptr,
                    const std::size_t lda, const double *const b_ptr, const std::size_t ldb, const double *beta,
                    double *const c_ptr, const std::size_t ldc, const compute_mode_t compute_mode,
                    const std::size_t num_panels, const std::size_t *col_edges, const cudaEvent_t *ready);

compute_mode_t auto_mode_select(handle_t handle, const operation_t op_A, const operation_t op_B,
                                const std::size_t m, const std::size_t n, const std::size_t k,
                                const void *const a_ptr, const std::size_t lda,
                                const void *const b_ptr, const std::size_t ldb,
                                const element_kind_t element_kind,
                                const double mantissa_loss_threshold);

// --- helpers -----------------------------------------------------------------------------------
std::string get_compute_mode_name_str(const compute_mode_t mode);
data_t get_output_type(const compute_mode_t mode);
std::size_t get_data_size_in_byte(const data_t d);
std::uint32_t get_bits_per_int8(const std::uint32_t k);

}  // namespace ozimmu
}  // namespace mtk
