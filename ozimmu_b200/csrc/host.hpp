// host.hpp -- host-side state of the library: handle, workspace, profiler, real-cuBLAS lookup.
// Functional equivalent of reference src/handle.hpp:6-31, src/utils.hpp:77-141.
#pragma once
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include <cublas_v2.h>  // types only: every cuBLAS entry point is resolved with dlsym at run time
#include <cuda_runtime.h>
#include <ozimmu/ozimmu.hpp>

namespace oz {
namespace host {

// ---- logging (reference src/utils.hpp:77-115: OZIMMU_INFO / OZIMMU_ERROR) ----------------------
bool env_enabled(const char *name, bool default_value);
void log_info(const std::string &msg);
void log_error(const std::string &msg);
std::string env_or(const char *name, const std::string &fallback);

// CUDA failure -> std::runtime_error, as the reference's CUTF_CHECK_ERROR does
// (reference src/cutf/include/cutf/cuda.hpp:11-21).
void cuda_check(cudaError_t e, const char *what, const char *file, int line);
#define OZ_CUDA_CHECK(expr) ::oz::host::cuda_check((expr), #expr, __FILE__, __LINE__)
// Kernel-ABI launchers return a cudaError_t value as int.
#define OZ_KERNEL_CHECK(expr) ::oz::host::cuda_check(static_cast<cudaError_t>(expr), #expr, __FILE__, __LINE__)

// ---- the real cuBLAS (reference src/utils.hpp:117-141: dlsym(RTLD_NEXT, name)) -----------------
// RTLD_NEXT first (LD_PRELOAD case), then a libcublas already mapped into the process, then an
// explicit dlopen (ctypes / direct-link case, where this library is not ahead of cuBLAS in the
// search order).  nullptr if unavailable.
void *real_cublas_symbol(const char *name);

// ---- per-stage profiler (reference: cutf::debug::time_breakdown::profiler in the handle) ------
struct StageProfiler {
  struct Entry {
    std::uint64_t count = 0;
    double seconds = 0;
  };
  bool enabled = false;
  std::map<std::string, Entry> entries;
  std::map<std::string, std::chrono::steady_clock::time_point> open;
  void start(const std::string &name, cudaStream_t s);
  void stop(const std::string &name, cudaStream_t s);
  void print(const std::string &tag, bool csv) const;
  void clear() { entries.clear(); open.clear(); }
};

// Workspace carve-up for one real fp64_int8_S GEMM.  The reference's layout
// (src/gemm.cu:359-379) also holds an m*n FP64 accumulator and an m*n int32 product buffer;
// the fused kernel needs neither.
struct WorkspaceLayout {
  std::size_t pitch;       // bytes of K per slice row (k rounded up to 128)
  std::size_t a_plane, b_plane;  // bytes of one operand plane (all slices; blocked layout, rows padded to 256)
  std::size_t off_amax, off_bmax, off_scr_a, off_scr_b, off_a_slices, off_b_slices;
  std::size_t total;
};
// planes = 1 (real) or 2 (complex: real and imaginary planes of each operand, back to back)
WorkspaceLayout workspace_layout(std::size_t m, std::size_t n, std::size_t k, unsigned num_split, unsigned planes = 1);

inline bool is_int8_mode(mtk::ozimmu::compute_mode_t mode) {
  return mode >= mtk::ozimmu::fp64_int8_3 && mode <= mtk::ozimmu::fp64_int8_18;
}
inline unsigned num_split_of(mtk::ozimmu::compute_mode_t mode) {
  return static_cast<unsigned>(mode) - static_cast<unsigned>(mtk::ozimmu::fp64_int8_3) + 3u;
}
inline mtk::ozimmu::compute_mode_t mode_of_num_split(unsigned s) {
  return static_cast<mtk::ozimmu::compute_mode_t>(static_cast<unsigned>(mtk::ozimmu::fp64_int8_3) + s - 3u);
}

// lazily created aux stream + ordering events of a handle
void ensure_streams(mtk::ozimmu::handle *h);
// lazily created copy / split / product streams and per-block events (host_e2e.cu)
void ensure_pipeline_streams(mtk::ozimmu::handle *h);

// compute mode `sgemm` (reference src/cublas_helper.cu:84-134): the GEMM in FP32 through cuBLAS (sgemm_mode.cu).
// `cublas` = a real cuBLAS handle bound to the handle's stream; alpha / beta: 1 (real) or 2 (complex) doubles.
void gemm_in_f32(mtk::ozimmu::handle *h, cublasHandle_t cublas, mtk::ozimmu::operation_t op_a,
                 mtk::ozimmu::operation_t op_b, std::size_t m, std::size_t n, std::size_t k, const double *alpha,
                 const double *a, std::size_t lda, const double *b, std::size_t ldb, const double *beta, double *c,
                 std::size_t ldc, mtk::ozimmu::element_kind_t kind);

// reference src/config.cu:85-92: the ordered (A_id, B_id) list of one fp64_int8_S product sweep
std::vector<std::pair<int, int>> pair_list(unsigned num_split);

}  // namespace host
}  // namespace oz

// reference src/handle.hpp:6-31
struct mtk::ozimmu::handle {
  cublasHandle_t cublas_handle = nullptr;  // private real-cuBLAS handle, created lazily (passthrough)
  cudaStream_t cuda_stream = nullptr;

  void *working_memory_ptr = nullptr;
  std::size_t current_working_memory_size = 0;

  oz::host::StageProfiler profiler;
  malloc_mode_t malloc_mode = malloc_sync;

  // auto mode: 16 counters, fp64_int8_3..18 (the reference sizes this 8)
  enum { mantissa_loss_counter_length = 16 };
  unsigned long long *d_mantissa_loss_counter_ptr = nullptr;
  unsigned long long *h_mantissa_loss_counter_ptr = nullptr;  // pinned
  unsigned long long last_loss_counters[mantissa_loss_counter_length] = {};
  double avg_mantissa_loss_threshold = 0;

  std::uint32_t intercept_threshold_m = 1024;
  std::uint32_t intercept_threshold_n = 1024;
  std::uint32_t intercept_threshold_k = 1024;

  // split(A) || split(B) overlap and cross-stream ordering of the shared workspace
  cudaStream_t aux_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_done = nullptr;
  cudaStream_t last_stream = nullptr;
  bool has_pending = false;

  // staging for ozimmu_gemm_host (grow-only device buffers + copy streams/events)
  void *stage_a = nullptr, *stage_b = nullptr, *stage_c = nullptr;
  std::size_t stage_a_bytes = 0, stage_b_bytes = 0, stage_c_bytes = 0;
  // compute_stream: the splits (and the one-shot path); product_stream[]: the fused launches, round-robin, so
  // that a launch back-fills the SMs the previous one leaves idle in its last round of tiles
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr, compute_stream = nullptr;
  static constexpr int kMaxBlocks = 16;        // row blocks of op(A) / column blocks of op(B)
  static constexpr int kProductStreams = 8;    // streams available; the pipelines rotate over the first 3 by default
  cudaStream_t product_stream[kProductStreams] = {};
  cudaEvent_t ev_block_in[2][kMaxBlocks] = {}, ev_block_split[2][kMaxBlocks] = {};  // [0] = A, [1] = B
  std::vector<cudaEvent_t> ev_rect_out;                                             // one per fused launch, grown on demand
  cudaEvent_t ev_product_tail[kProductStreams] = {};
  // alpha / beta of gemm() are device pointers (set by the interposers when the application's cuBLAS handle is in
  // CUBLAS_POINTER_MODE_DEVICE; the reference dereferences them on the host regardless, src/gemm.cu:405)
  bool scalars_on_device = false;
  int device = 0;
};
