// api.cu -- the public interface: namespace mtk::ozimmu (source-compatible with reference
// include/ozimmu/ozimmu.hpp:47-100) and its plain-C spelling (include/ozimmu_b200.h).
//
// Host orchestration of one DGEMM (replaces reference src/gemm.cu:344-410 gemm_int8<double> and
// :524-653 mtk::ozimmu::gemm, src/handle.cu, src/split.cu:454-518 auto_mode_select):
//
//   stream:      [wait prev]  split(A) -------------+
//   aux stream:  [fork]       split(B) --[join]------> fused tcgen05 product+accumulate+finalize
//
// No device-wide synchronisation, no per-call allocation once the workspace has grown, two
// (row-contiguous operand) or three launches per operand-free call instead of the reference's
// 2*P(s)+4.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "host.hpp"
#include "oz_common.cuh"
#include "ozimmu_b200.h"

using namespace mtk::ozimmu;
namespace H = oz::host;

void oz::host::ensure_streams(mtk::ozimmu::handle *h) {
  if (h->aux_stream) return;
  OZ_CUDA_CHECK(cudaGetDevice(&h->device));
  OZ_CUDA_CHECK(cudaStreamCreateWithFlags(&h->aux_stream, cudaStreamNonBlocking));
  OZ_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  OZ_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
  OZ_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming));
}

namespace {
using oz::host::ensure_streams;

// reference src/utils.hpp:143-156 check_gemm_shape
int check_shape(operation_t op, std::size_t rows, std::size_t cols, std::size_t ld, const char *name) {
  const std::size_t need = (op == op_n) ? rows : cols;
  if (need > ld) {
    H::log_error(std::string("The leading dimension of ") + name + " (" + std::to_string(ld) +
                 ") must be larger or equal to the number of " + (op == op_n ? "rows" : "cols") + " (" +
                 std::to_string(need) + ")");
    return 1;
  }
  return 0;
}

// reference src/utils.hpp:158-168 check_address_alignment
int check_alignment(const void *p, std::size_t size, const char *name) {
  if (reinterpret_cast<std::uintptr_t>(p) % size) {
    H::log_error(std::string("Invalid address alignment for matrix ") + name);
    return 1;
  }
  return 0;
}

cublasHandle_t private_cublas(handle_t h) {
  if (h->cublas_handle == nullptr) {
    using Fn = cublasStatus_t (*)(cublasHandle_t *);
    auto fn = reinterpret_cast<Fn>(H::real_cublas_symbol("cublasCreate_v2"));
    if (fn == nullptr || fn(&h->cublas_handle) != CUBLAS_STATUS_SUCCESS)
      throw std::runtime_error("ozIMMU: cuBLAS is not available for the dgemm passthrough");
  }
  using SetStream = cublasStatus_t (*)(cublasHandle_t, cudaStream_t);
  auto set_stream = reinterpret_cast<SetStream>(H::real_cublas_symbol("cublasSetStream_v2"));
  if (set_stream) set_stream(h->cublas_handle, h->cuda_stream);
  return h->cublas_handle;
}

// Order this call after the previous user of the shared workspace if that one ran on another stream.  Work recorded
// into a CUDA graph is not "pending" in that sense: an event recorded during capture cannot be waited on outside the
// graph (and vice versa), and whoever replays the graph orders the replays against other users of the handle, as with
// any captured library call that owns scratch memory.
bool is_capturing(cudaStream_t s) {
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(s, &st) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return st != cudaStreamCaptureStatusNone;
}
void wait_previous(handle_t h, cudaStream_t s) {
  if (h->has_pending && h->last_stream != s && !is_capturing(s)) OZ_CUDA_CHECK(cudaStreamWaitEvent(s, h->ev_done, 0));
}
void mark_done(handle_t h, cudaStream_t s) {
  if (is_capturing(s)) {
    h->has_pending = false;
    return;
  }
  OZ_CUDA_CHECK(cudaEventRecord(h->ev_done, s));
  h->has_pending = true;
  h->last_stream = s;
}

// alpha / beta of one call: host values, or device pointers (cuBLAS device pointer mode; read by the kernels in
// stream order, never on the host)
struct Scalars {
  double alpha[2] = {0, 0}, beta[2] = {0, 0};
  const double *alpha_dev = nullptr, *beta_dev = nullptr;
};

// Complex GEMM: fold the four plane products inside the fused launch (every tile runs 4 x P(s) products), or compute
// them as 4 x as many real work items into scratch and combine afterwards?  The second form fills the CTA pairs in
// finer grains; it pays when it saves more rounds of tiles than its extra pass over C costs.  Round time from
// profiles/r2_sweep_tile_width.txt (1.28 ms per round of 256-wide tiles at P = 45, k = 8192), the combine pass at
// ~100 bytes per element and ~4 TB/s.  OZIMMU_B200_ZGEMM_PLANES_FIRST=0/1 forces the choice.
bool complex_planes_first(std::size_t m, std::size_t n, std::size_t pitch, unsigned num_split) {
  const std::string forced = H::env_or("OZIMMU_B200_ZGEMM_PLANES_FIRST", "");
  if (forced == "0") return false;
  if (forced == "1") return true;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const std::size_t pairs = std::max(1, sms / 2);
  const std::size_t tiles = ((m + 255) / 256) * ((n + 255) / 256);
  const double round_ms = 1.28 * (static_cast<double>(num_split * (num_split + 1) / 2) * static_cast<double>(pitch)) / (45.0 * 8192.0);
  const double fused = 4.0 * static_cast<double>((tiles + pairs - 1) / pairs) * round_ms;
  const double split = 2.0 * static_cast<double>((2 * tiles + pairs - 1) / pairs) * round_ms +
                       100.0 * static_cast<double>(m) * static_cast<double>(n) / 4.0e9;
  return split < 0.97 * fused;
}

// The Ozaki path of every fp64_int8_S GEMM -- real (reference src/gemm.cu:344-410 gemm_int8<double>) or complex
// (:412-521 gemm_int8<cuDoubleComplex>: the real and imaginary planes of each operand are split independently and four
// real plane products are folded into C), a single product or a strided batch (the reference's interposers run one
// GEMM per entry, src/cublas.cu:380-406):
//   * split: ONE launch per operand plane for the whole chunk of entries, A on the caller's stream, B on the aux stream
//     (an operand shared by all entries, stride 0, is split once);
//   * products: ONE grouped launch whose tile queue spans all entries; a complex tile runs its four plane products
//     back to back, so a complex GEMM is one launch as well (the reference: 4 x (P(s) cuBLAS GEMMs + P(s) + 1 kernels)).
// Entries are processed in chunks that keep the workspace below OZIMMU_B200_BATCH_WORKSPACE_MB (default 8192).
// Strides count elements (complex elements for complex data).
void gemm_int8(handle_t h, operation_t op_a, operation_t op_b, std::size_t m, std::size_t n, std::size_t k,
               const Scalars &sc, const double *a, std::size_t lda, long long stride_a, const double *b, std::size_t ldb,
               long long stride_b, double *c, std::size_t ldc, long long stride_c, std::size_t batch, unsigned num_split,
               element_kind_t kind) {
  if (m == 0 || n == 0 || batch == 0) return;
  const unsigned es = kind == real ? 1u : 2u;   // doubles per element = planes per operand
  cudaStream_t s = h->cuda_stream;
  if (k == 0) {
    for (std::size_t e = 0; e < batch; e++)
      OZ_KERNEL_CHECK(ozk_scale_c_ex(m, n, sc.beta, sc.beta_dev, es == 2, c + static_cast<long long>(e) * stride_c * es,
                                     ldc, s));
    return;
  }
  const unsigned bits = ozk_bits_per_int8(static_cast<std::uint32_t>(k));
  const H::WorkspaceLayout w = H::workspace_layout(m, n, k, num_split, es);
  const std::size_t limit = std::stoull(H::env_or("OZIMMU_B200_BATCH_WORKSPACE_MB", "8192")) << 20;
  const std::size_t chunk = std::max<std::size_t>(1, std::min<std::size_t>({batch, limit / w.total, 65535}));
  const bool planes_first = es == 2 && batch == 1 && complex_planes_first(m, n, w.pitch, num_split);
  reallocate_working_memory(h, w.total * chunk + (planes_first ? 4 * sizeof(double) * m * n : 0));
  ensure_streams(h);
  char *ws = static_cast<char *>(h->working_memory_ptr);
  wait_previous(h, s);
  // "rows" of the split are rows of op(A) and columns of op(B) (reference src/split.cu:266-283):
  //   op_n A (m x k, col-major)  -> row r strided by lda  -> col_major
  //   op_n B (k x n, col-major)  -> column j contiguous   -> !col_major
  const int a_col_major = (op_a == op_n), b_col_major = (op_b != op_n);
  const bool overlap = !h->profiler.enabled;
  cudaStream_t sb = overlap ? h->aux_stream : s;
  for (std::size_t e0 = 0; e0 < batch; e0 += chunk) {
    const std::size_t ne = std::min(chunk, batch - e0);
    if (overlap) {
      OZ_CUDA_CHECK(cudaEventRecord(h->ev_fork, s));   // also: the previous chunk's products are done with the slices
      OZ_CUDA_CHECK(cudaStreamWaitEvent(sb, h->ev_fork, 0));
    }
    // one split launch per operand plane for the whole chunk (entry e's workspace is ws + e * w.total)
    auto split_chunk = [&](const double *x, long long stride, std::size_t ld, std::size_t rows, int col_major,
                           std::size_t off_slices, std::size_t plane_bytes, std::size_t off_max, std::size_t off_scr,
                           cudaStream_t st) {
      for (unsigned part = 0; part < es; part++) {
        auto *out = reinterpret_cast<std::int8_t *>(ws + off_slices + part * plane_bytes);
        auto *mx = reinterpret_cast<double *>(ws + off_max) + part * rows;
        auto *scr = reinterpret_cast<std::uint32_t *>(ws + off_scr) + part * rows;
        if (stride >= 0) {
          // stride 0: one operand shared by every entry -- split once, the grouped launch reads it with stride 0
          const std::size_t count = stride == 0 ? 1 : ne;
          OZ_KERNEL_CHECK(ozk_split_int8_batched_strided(
              out, w.total, w.pitch, mx, w.total / sizeof(double), scr, w.total / sizeof(std::uint32_t), rows, k,
              x + static_cast<long long>(e0) * stride * es + part, ld, static_cast<std::size_t>(stride) * es, col_major,
              num_split, bits, es, count, st));
          continue;
        }
        for (std::size_t e = 0; e < ne; e++)   // backward-strided input: entry by entry
          OZ_KERNEL_CHECK(ozk_split_int8_strided(out + e * w.total, w.pitch, mx + e * (w.total / sizeof(double)),
                                                 scr + e * (w.total / sizeof(std::uint32_t)), rows, k,
                                                 x + static_cast<long long>(e0 + e) * stride * es + part, ld, col_major,
                                                 num_split, bits, es, st));
      }
    };
    h->profiler.start("split_A", s);
    split_chunk(a, stride_a, lda, m, a_col_major, w.off_a_slices, w.a_plane, w.off_amax, w.off_scr_a, s);
    h->profiler.stop("split_A", s);
    h->profiler.start("split_B", sb);
    split_chunk(b, stride_b, ldb, n, b_col_major, w.off_b_slices, w.b_plane, w.off_bmax, w.off_scr_b, sb);
    h->profiler.stop("split_B", sb);
    if (overlap) {
      OZ_CUDA_CHECK(cudaEventRecord(h->ev_join, sb));
      OZ_CUDA_CHECK(cudaStreamWaitEvent(s, h->ev_join, 0));
    }
    // A single complex GEMM whose tile count quantises badly: its four plane products as 4 x as many REAL work items
    // (two grouped launches of two entries each: the planes of A against one plane of B) into scratch, then one
    // elementwise pass that folds them into C in the reference's order -- e.g. 4096^3: 2 x 7 rounds of tiles instead of
    // 4 x 4, at the price of ~100 bytes of scratch traffic per element of C.  Same bits either way.
    if (es == 2 && batch == 1 && complex_planes_first(m, n, w.pitch, num_split)) {
      double *x4 = reinterpret_cast<double *>(ws + w.total);
      h->profiler.start("int8tc_accumulate_fused", s);
      for (unsigned bp = 0; bp < 2; bp++) {
        ozk_fused_args_t fa{};
        fa.m = m, fa.n = n, fa.k = k, fa.pitch = w.pitch;
        fa.a_slices = reinterpret_cast<const std::int8_t *>(ws + w.off_a_slices);
        fa.b_slices = reinterpret_cast<const std::int8_t *>(ws + w.off_b_slices + bp * w.b_plane);
        fa.amax = reinterpret_cast<const double *>(ws + w.off_amax);
        fa.bmax = reinterpret_cast<const double *>(ws + w.off_bmax) + bp * n;
        fa.num_split = num_split, fa.bits_per_int8 = bits;
        fa.alpha[0] = 1.0;
        fa.c = x4 + static_cast<std::size_t>(2 * bp) * m * n;
        fa.ldc = m;
        fa.batch = 2;   // entry = plane of A
        fa.a_batch_bytes = w.a_plane, fa.b_batch_bytes = 0, fa.amax_batch = m, fa.bmax_batch = 0, fa.c_batch = m * n;
        OZ_KERNEL_CHECK(ozk_gemm_i8_fused_ex(&fa, s));
      }
      OZ_KERNEL_CHECK(ozk_zgemm_combine(m, n, x4, sc.alpha, sc.beta, sc.alpha_dev, sc.beta_dev, c, ldc, s));
      h->profiler.stop("int8tc_accumulate_fused", s);
      continue;
    }
    ozk_fused_args_t fa{};
    fa.m = m, fa.n = n, fa.k = k, fa.pitch = w.pitch;
    fa.a_slices = reinterpret_cast<const std::int8_t *>(ws + w.off_a_slices);
    fa.b_slices = reinterpret_cast<const std::int8_t *>(ws + w.off_b_slices);
    fa.amax = reinterpret_cast<const double *>(ws + w.off_amax);
    fa.bmax = reinterpret_cast<const double *>(ws + w.off_bmax);
    fa.num_split = num_split, fa.bits_per_int8 = bits;
    fa.complex_c = es == 2;
    fa.a_plane_bytes = w.a_plane, fa.b_plane_bytes = w.b_plane, fa.amax_plane = m, fa.bmax_plane = n;
    fa.alpha[0] = sc.alpha[0], fa.alpha[1] = sc.alpha[1], fa.beta[0] = sc.beta[0], fa.beta[1] = sc.beta[1];
    fa.alpha_dev = sc.alpha_dev, fa.beta_dev = sc.beta_dev;
    fa.c = c + static_cast<long long>(e0) * stride_c * es;
    fa.ldc = ldc;
    fa.batch = ne;
    fa.a_batch_bytes = stride_a == 0 ? 0 : w.total, fa.b_batch_bytes = stride_b == 0 ? 0 : w.total;
    fa.amax_batch = stride_a == 0 ? 0 : w.total / sizeof(double), fa.bmax_batch = stride_b == 0 ? 0 : w.total / sizeof(double);
    fa.c_batch = static_cast<std::size_t>(stride_c);
    h->profiler.start("int8tc_accumulate_fused", s);
    OZ_KERNEL_CHECK(ozk_gemm_i8_fused_ex(&fa, s));
    h->profiler.stop("int8tc_accumulate_fused", s);
  }
  mark_done(h, s);
}

// host copies of device-resident scalars (a blocking read; only the paths that need the values on the host -- auto
// mode, dgemm / sgemm through the private cuBLAS handle -- use it)
Scalars fetch_scalars(handle_t h, const Scalars &sc, element_kind_t kind) {
  if (sc.alpha_dev == nullptr) return sc;
  Scalars out;
  const std::size_t bytes = sizeof(double) * (kind == real ? 1 : 2);
  OZ_CUDA_CHECK(cudaMemcpyAsync(out.alpha, sc.alpha_dev, bytes, cudaMemcpyDeviceToHost, h->cuda_stream));
  OZ_CUDA_CHECK(cudaMemcpyAsync(out.beta, sc.beta_dev, bytes, cudaMemcpyDeviceToHost, h->cuda_stream));
  OZ_CUDA_CHECK(cudaStreamSynchronize(h->cuda_stream));
  return out;
}

Scalars read_scalars(handle_t h, const void *alpha, const void *beta, element_kind_t kind) {
  Scalars sc;
  if (h->scalars_on_device) {
    sc.alpha_dev = static_cast<const double *>(alpha);
    sc.beta_dev = static_cast<const double *>(beta);
    return sc;
  }
  sc.alpha[0] = static_cast<const double *>(alpha)[0];
  sc.beta[0] = static_cast<const double *>(beta)[0];
  if (kind != real) {
    sc.alpha[1] = static_cast<const double *>(alpha)[1];
    sc.beta[1] = static_cast<const double *>(beta)[1];
  }
  return sc;
}

template <class F>
int guarded(F &&f) {
  try {
    return f();
  } catch (const std::exception &e) {
    H::log_error(e.what());
    return -1;
  }
}

}  // namespace

// ===============================================================================================
// namespace mtk::ozimmu
// ===============================================================================================
int mtk::ozimmu::create(handle_t *handle, const malloc_mode_t mm) {
  H::log_info("Initializing ozIMMU handle");
  auto h = (*handle = new mtk::ozimmu::handle);
  h->malloc_mode = mm;
  OZ_CUDA_CHECK(cudaMalloc(&h->d_mantissa_loss_counter_ptr,
                           sizeof(unsigned long long) * handle::mantissa_loss_counter_length));
  OZ_CUDA_CHECK(cudaMallocHost(&h->h_mantissa_loss_counter_ptr,
                               sizeof(unsigned long long) * handle::mantissa_loss_counter_length));
  // reference src/handle.cu:25-30 (defaults 1024; README says 128 -- SURVEY App. B.3)
  h->intercept_threshold_m = std::stoul(H::env_or("OZIMMU_INTERCEPT_THRESHOLD_M", "1024"));
  h->intercept_threshold_n = std::stoul(H::env_or("OZIMMU_INTERCEPT_THRESHOLD_N", "1024"));
  h->intercept_threshold_k = std::stoul(H::env_or("OZIMMU_INTERCEPT_THRESHOLD_K", "1024"));
  return 0;
}

int mtk::ozimmu::destroy(handle_t h) {
  if (h == nullptr) return 0;
  H::log_info("Destroying ozIMMU handle");
  cudaDeviceSynchronize();
  if (h->cublas_handle) {
    using Fn = cublasStatus_t (*)(cublasHandle_t);
    if (auto fn = reinterpret_cast<Fn>(H::real_cublas_symbol("cublasDestroy_v2"))) fn(h->cublas_handle);
  }
  cudaFree(h->working_memory_ptr);
  cudaFree(h->d_mantissa_loss_counter_ptr);
  cudaFreeHost(h->h_mantissa_loss_counter_ptr);
  cudaFree(h->stage_a);
  cudaFree(h->stage_b);
  cudaFree(h->stage_c);
  for (cudaStream_t s : {h->aux_stream, h->h2d_stream, h->d2h_stream, h->compute_stream})
    if (s) cudaStreamDestroy(s);
  for (cudaStream_t s : h->product_stream)
    if (s) cudaStreamDestroy(s);
  for (cudaEvent_t e : {h->ev_fork, h->ev_join, h->ev_done})
    if (e) cudaEventDestroy(e);
  for (auto &row : h->ev_block_in)
    for (cudaEvent_t e : row)
      if (e) cudaEventDestroy(e);
  for (auto &row : h->ev_block_split)
    for (cudaEvent_t e : row)
      if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : h->ev_rect_out)
    if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : h->ev_product_tail)
    if (e) cudaEventDestroy(e);
  delete h;
  return 0;
}

void mtk::ozimmu::set_cuda_stream(handle_t handle, const cudaStream_t cuda_stream) {
  handle->cuda_stream = cuda_stream;
}

void mtk::ozimmu::enable_profiling(handle_t handle) { handle->profiler.enabled = true; }
void mtk::ozimmu::disable_profiling(handle_t handle) { handle->profiler.enabled = false; }
void mtk::ozimmu::print_profiler_result(handle_t handle, const std::string tag, const bool csv) {
  handle->profiler.print(tag, csv);
}
void mtk::ozimmu::clear_profiler_result(handle_t handle) { handle->profiler.clear(); }

void mtk::ozimmu::set_auto_mantissa_loss_threashold(handle_t handle, const double threshold) {
  handle->avg_mantissa_loss_threshold = threshold;
}
double mtk::ozimmu::get_auto_mantissa_loss_threashold(handle_t handle) {
  return handle->avg_mantissa_loss_threshold;
}

// reference src/handle.cu:63-93 (grow-only)
std::size_t mtk::ozimmu::reallocate_working_memory(handle_t h, const std::size_t size_in_byte) {
  if (size_in_byte <= h->current_working_memory_size) return 0;
  H::log_info("Reallocated memory : " + std::to_string(size_in_byte) + " B");
  if (h->working_memory_ptr != nullptr) {
    if (h->malloc_mode == malloc_sync) {
      OZ_CUDA_CHECK(cudaDeviceSynchronize());
      OZ_CUDA_CHECK(cudaFree(h->working_memory_ptr));
    } else {
      // the old block may still be in use by work queued on the previous stream
      if (h->has_pending) OZ_CUDA_CHECK(cudaStreamWaitEvent(h->cuda_stream, h->ev_done, 0));
      OZ_CUDA_CHECK(cudaFreeAsync(h->working_memory_ptr, h->cuda_stream));
    }
    h->working_memory_ptr = nullptr;
  }
  h->current_working_memory_size = 0;
  if (h->malloc_mode == malloc_sync) {
    OZ_CUDA_CHECK(cudaMalloc(&h->working_memory_ptr, size_in_byte));
  } else {
    OZ_CUDA_CHECK(cudaMallocAsync(&h->working_memory_ptr, size_in_byte, h->cuda_stream));
  }
  h->current_working_memory_size = size_in_byte;
  return size_in_byte;
}

// reference src/handle.cu:95-144 -- sized for this library's layout (no FP64/int32 m*n buffers)
std::size_t mtk::ozimmu::reallocate_working_memory(handle_t h, const gemm_list_t gemm_list) {
  std::size_t need = 0;
  for (const auto &g : gemm_list) {
    const std::size_t m = std::get<2>(g), n = std::get<3>(g), k = std::get<4>(g);
    const element_kind_t kind = std::get<5>(g);
    const compute_mode_t mode = std::get<6>(g);
    unsigned s = 0;
    if (H::is_int8_mode(mode)) s = H::num_split_of(mode);
    else if (mode == fp64_int8_auto) s = 18;
    if (s == 0) continue;
    const std::size_t bytes = H::workspace_layout(m, n, k, s, kind == complx ? 2 : 1).total;
    need = std::max(need, bytes);
  }
  return reallocate_working_memory(h, need);
}

std::string mtk::ozimmu::get_compute_mode_name_str(const compute_mode_t mode) {
  if (mode == sgemm) return "sgemm";
  if (mode == dgemm) return "dgemm";
  if (mode == fp64_int8_auto) return "fp64_int8_auto";
  if (H::is_int8_mode(mode)) return "fp64_int8_" + std::to_string(H::num_split_of(mode));
  throw std::runtime_error("ozIMMU: unknown compute mode " + std::to_string(static_cast<int>(mode)));
}

// reference src/handle.cu:194-225
data_t mtk::ozimmu::get_output_type(const compute_mode_t mode) {
  if (mode == sgemm) return fp32;
  if (mode == dgemm || mode == fp64_int8_auto || H::is_int8_mode(mode)) return fp64;
  throw std::runtime_error("ozIMMU: unknown compute mode " + std::to_string(static_cast<int>(mode)));
}

// reference src/handle.cu:227-244
std::size_t mtk::ozimmu::get_data_size_in_byte(const data_t d) {
  switch (d) {
    case fp64: return 8;
    case fp32: return 4;
    case fp16: return 2;
    case int8: return 1;
    case none: return 0;
    default: throw std::runtime_error("ozIMMU: data type has no size");
  }
}

std::uint32_t mtk::ozimmu::get_bits_per_int8(const std::uint32_t k) { return ozk_bits_per_int8(k); }

// reference src/split.cu:383-445 get_mantissa_loss_total, for `batch` problems of the same shape at once: ONE counter
// pass per operand plane over all entries (the reference's strided-batched interposers run auto mode entry by entry
// with a blocking read each, src/cublas.cu:380-406), then one blocking read.  out: [batch][16].
namespace {
void mantissa_loss_totals(handle_t h, operation_t op_A, operation_t op_B, std::size_t m, std::size_t n, std::size_t k,
                          const double *a, std::size_t lda, long long stride_a, const double *b, std::size_t ldb,
                          long long stride_b, std::size_t batch, element_kind_t kind,
                          std::vector<unsigned long long> &out) {
  constexpr int N = handle::mantissa_loss_counter_length;
  out.assign(batch * N, 0);
  if (m == 0 || n == 0 || k == 0 || batch == 0) return;
  const unsigned es = kind == real ? 1u : 2u;  // complex: both planes, each against its own row scale
  const unsigned bits = ozk_bits_per_int8(static_cast<std::uint32_t>(k));
  ensure_streams(h);
  cudaStream_t s = h->cuda_stream;
  std::vector<unsigned long long> part(N);
  for (std::size_t e0 = 0; e0 < batch;) {
    // workspace of a chunk: [counters ne x 16 u64][row-max scratch ne x max(m, n) u32]
    const std::size_t ne = std::min<std::size_t>(batch - e0, 4096);
    const std::size_t scr_rows = std::max(m, n);
    const std::size_t cnt_bytes = sizeof(unsigned long long) * N * ne;
    reallocate_working_memory(h, cnt_bytes + sizeof(std::uint32_t) * scr_rows * ne);
    wait_previous(h, s);
    auto *cnt = static_cast<unsigned long long *>(h->working_memory_ptr);
    auto *scr = reinterpret_cast<std::uint32_t *>(static_cast<char *>(h->working_memory_ptr) + cnt_bytes);
    OZ_CUDA_CHECK(cudaMemsetAsync(cnt, 0, cnt_bytes, s));
    // a negative stride walks backwards: entry by entry (the batched launch takes unsigned strides)
    auto pass = [&](const double *x, long long stride, std::size_t ld, std::size_t rows, int col_major) {
      for (unsigned p = 0; p < es; p++) {
        if (stride >= 0) {
          OZ_KERNEL_CHECK(ozk_mantissa_loss_batched(cnt, scr, scr_rows, rows, k, x + static_cast<long long>(e0) * stride * es + p,
                                                    ld, static_cast<std::size_t>(stride) * es, col_major, bits, es, ne, s));
        } else {
          for (std::size_t e = 0; e < ne; e++)
            OZ_KERNEL_CHECK(ozk_mantissa_loss_strided(cnt + e * N, scr, rows, k,
                                                      x + static_cast<long long>(e0 + e) * stride * es + p, ld, col_major,
                                                      bits, es, s));
        }
      }
    };
    pass(a, stride_a, lda, m, op_A == op_n);
    pass(b, stride_b, ldb, n, op_B != op_n);
    OZ_CUDA_CHECK(cudaMemcpyAsync(out.data() + e0 * N, cnt, cnt_bytes, cudaMemcpyDeviceToHost, s));
    mark_done(h, s);
    OZ_CUDA_CHECK(cudaStreamSynchronize(s));
    e0 += ne;
  }
}

// reference src/split.cu:484-493: first split count whose average loss is within the threshold, else dgemm
compute_mode_t mode_for_loss(const unsigned long long *counters, std::size_t m, std::size_t n, std::size_t k,
                             double threshold) {
  const double denom = static_cast<double>(m * k + k * n);
  for (int i = 0; i < handle::mantissa_loss_counter_length; i++)
    if (static_cast<double>(counters[i]) / denom <= threshold) return H::mode_of_num_split(static_cast<unsigned>(i) + 3u);
  return dgemm;
}
}  // namespace

// reference src/split.cu:454-518
compute_mode_t mtk::ozimmu::auto_mode_select(handle_t h, const operation_t op_A, const operation_t op_B,
                                             const std::size_t m, const std::size_t n, const std::size_t k,
                                             const void *const a_ptr, const std::size_t lda,
                                             const void *const b_ptr, const std::size_t ldb,
                                             const element_kind_t element_kind,
                                             const double mantissa_loss_threshold) {
  constexpr int N = handle::mantissa_loss_counter_length;
  for (int i = 0; i < N; i++) h->last_loss_counters[i] = 0;
  if (m == 0 || n == 0 || k == 0) return fp64_int8_3;
  std::vector<unsigned long long> totals;
  mantissa_loss_totals(h, op_A, op_B, m, n, k, static_cast<const double *>(a_ptr), lda, 0,
                       static_cast<const double *>(b_ptr), ldb, 0, 1, element_kind, totals);
  for (int i = 0; i < N; i++) h->last_loss_counters[i] = totals[i];
  return mode_for_loss(h->last_loss_counters, m, n, k, mantissa_loss_threshold);
}

int mtk::ozimmu::gemm(handle_t h, const operation_t op_A, const operation_t op_B, const std::size_t m,
                      const std::size_t n, const std::size_t k, const void *alpha, const void *const a_ptr,
                      const std::size_t lda, const void *const b_ptr, const std::size_t ldb, const void *beta,
                      void *const c_ptr, std::size_t ldc, const compute_mode_t compute_mode,
                      const element_kind_t element_kind) {
  int arg_error = 0;
  arg_error |= check_shape(op_A, m, k, lda, "A");
  arg_error |= check_shape(op_B, k, n, ldb, "B");
  arg_error |= check_shape(op_n, m, n, ldc, "C");
  const std::size_t esz = element_kind == real ? sizeof(double) : 2 * sizeof(double);
  arg_error |= check_alignment(a_ptr, esz, "A");
  arg_error |= check_alignment(b_ptr, esz, "B");
  arg_error |= check_alignment(c_ptr, esz, "C");
  if (arg_error) return 1;

  const Scalars sc = read_scalars(h, alpha, beta, element_kind);
  if (H::is_int8_mode(compute_mode)) {
    gemm_int8(h, op_A, op_B, m, n, k, sc, static_cast<const double *>(a_ptr), lda, 0, static_cast<const double *>(b_ptr),
              ldb, 0, static_cast<double *>(c_ptr), ldc, 0, 1, H::num_split_of(compute_mode), element_kind);
    return 0;
  }
  if (h->scalars_on_device) {
    // every other mode needs the values on the host (auto mode reads its counters anyway; dgemm / sgemm go through the
    // private cuBLAS handle, which is in host pointer mode): one blocking read, then the host-scalar path
    const Scalars hs = fetch_scalars(h, sc, element_kind);
    h->scalars_on_device = false;
    int rc = 0;
    try {
      rc = gemm(h, op_A, op_B, m, n, k, hs.alpha, a_ptr, lda, b_ptr, ldb, hs.beta, c_ptr, ldc, compute_mode, element_kind);
    } catch (...) {
      h->scalars_on_device = true;
      throw;
    }
    h->scalars_on_device = true;
    return rc;
  }
  if (compute_mode == fp64_int8_auto) {
    const compute_mode_t chosen = auto_mode_select(h, op_A, op_B, m, n, k, a_ptr, lda, b_ptr, ldb, element_kind,
                                                   h->avg_mantissa_loss_threshold);
    H::log_info("AUTO selected mode = " + get_compute_mode_name_str(chosen) +
                ", threshold average mantissa loss = " + std::to_string(h->avg_mantissa_loss_threshold));
    return gemm(h, op_A, op_B, m, n, k, alpha, a_ptr, lda, b_ptr, ldb, beta, c_ptr, ldc, chosen, element_kind);
  }
  if (compute_mode == dgemm) {
    using Fn = cublasStatus_t (*)(cublasHandle_t, cublasOperation_t, cublasOperation_t, int, int, int, const void *,
                                  const void *, cudaDataType_t, int, const void *, cudaDataType_t, int, const void *,
                                  void *, cudaDataType_t, int, cublasComputeType_t, cublasGemmAlgo_t);
    auto fn = reinterpret_cast<Fn>(H::real_cublas_symbol("cublasGemmEx"));
    if (fn == nullptr) throw std::runtime_error("ozIMMU: cublasGemmEx is not available for the dgemm passthrough");
    const cudaDataType_t dt = element_kind == real ? CUDA_R_64F : CUDA_C_64F;
    const cublasStatus_t st =
        fn(private_cublas(h), op_A == op_n ? CUBLAS_OP_N : CUBLAS_OP_T, op_B == op_n ? CUBLAS_OP_N : CUBLAS_OP_T,
           static_cast<int>(m), static_cast<int>(n), static_cast<int>(k), alpha, a_ptr, dt, static_cast<int>(lda),
           b_ptr, dt, static_cast<int>(ldb), beta, c_ptr, dt, static_cast<int>(ldc), CUBLAS_COMPUTE_64F,
           CUBLAS_GEMM_DEFAULT);
    if (st != CUBLAS_STATUS_SUCCESS)
      throw std::runtime_error("ozIMMU: cublasGemmEx passthrough failed with status " + std::to_string(st));
    return 0;
  }
  if (compute_mode == sgemm) {
    ensure_streams(h);
    H::gemm_in_f32(h, private_cublas(h), op_A, op_B, m, n, k, static_cast<const double *>(alpha),
                   static_cast<const double *>(a_ptr), lda, static_cast<const double *>(b_ptr), ldb,
                   static_cast<const double *>(beta), static_cast<double *>(c_ptr), ldc, element_kind);
    return 0;
  }
  throw std::runtime_error("ozIMMU: compute mode " + get_compute_mode_name_str(compute_mode) +
                           " is not implemented");
}

// Not in the reference: B becomes valid column panel by column panel (the multi-GPU path broadcasts it that way,
// SURVEY 8e), and every panel of C is computed as soon as its columns of B are there.  split(A) and the products of
// the panels that have landed run while the rest of B is still on the wire; the launches rotate over the product
// streams so that each back-fills the SMs its predecessor leaves idle in its last round of tiles.
int mtk::ozimmu::gemm_streamed_b(handle_t h, const operation_t op_A, const operation_t op_B, const std::size_t m,
                                 const std::size_t n, const std::size_t k, const double *alpha, const double *const a_ptr,
                                 const std::size_t lda, const double *const b_ptr, const std::size_t ldb,
                                 const double *beta, double *const c_ptr, const std::size_t ldc,
                                 const compute_mode_t compute_mode, const std::size_t num_panels,
                                 const std::size_t *col_edges, const cudaEvent_t *ready) {
  int arg_error = 0;
  arg_error |= check_shape(op_A, m, k, lda, "A");
  arg_error |= check_shape(op_B, k, n, ldb, "B");
  arg_error |= check_shape(op_n, m, n, ldc, "C");
  arg_error |= check_alignment(a_ptr, sizeof(double), "A");
  arg_error |= check_alignment(b_ptr, sizeof(double), "B");
  arg_error |= check_alignment(c_ptr, sizeof(double), "C");
  if (num_panels == 0 || num_panels > static_cast<std::size_t>(handle::kMaxBlocks) || col_edges == nullptr ||
      ready == nullptr || col_edges[0] != 0 || col_edges[num_panels] != n) {
    H::log_error("gemm_streamed_b: 1.." + std::to_string(handle::kMaxBlocks) +
                 " panels, col_edges[0] = 0, col_edges[num_panels] = n");
    arg_error |= 1;
  } else {
    for (std::size_t p = 0; p < num_panels; p++)
      if (col_edges[p] > col_edges[p + 1] || (p + 1 < num_panels && col_edges[p + 1] % 256 != 0)) {
        H::log_error("gemm_streamed_b: inner panel edges must be ascending multiples of 256");
        arg_error |= 1;
        break;
      }
  }
  if (arg_error) return 1;
  cudaStream_t s = h->cuda_stream;
  if (!H::is_int8_mode(compute_mode) || k == 0 || m == 0 || n == 0 || h->profiler.enabled) {
    // auto mode looks at all of B first, dgemm is cuBLAS: wait for every panel, then the plain call
    for (std::size_t p = 0; p < num_panels; p++) OZ_CUDA_CHECK(cudaStreamWaitEvent(s, ready[p], 0));
    return gemm(h, op_A, op_B, m, n, k, alpha, a_ptr, lda, b_ptr, ldb, beta, c_ptr, ldc, compute_mode, real);
  }
  const unsigned num_split = H::num_split_of(compute_mode);
  const unsigned bits = ozk_bits_per_int8(static_cast<std::uint32_t>(k));
  const H::WorkspaceLayout w = H::workspace_layout(m, n, k, num_split);
  reallocate_working_memory(h, w.total);
  ensure_streams(h);
  H::ensure_pipeline_streams(h);
  char *ws = static_cast<char *>(h->working_memory_ptr);
  double *amax = reinterpret_cast<double *>(ws + w.off_amax);
  double *bmax = reinterpret_cast<double *>(ws + w.off_bmax);
  auto *scr_a = reinterpret_cast<std::uint32_t *>(ws + w.off_scr_a);
  auto *scr_b = reinterpret_cast<std::uint32_t *>(ws + w.off_scr_b);
  auto *a_sl = reinterpret_cast<std::int8_t *>(ws + w.off_a_slices);
  auto *b_sl = reinterpret_cast<std::int8_t *>(ws + w.off_b_slices);
  wait_previous(h, s);
  cudaStream_t sb = h->aux_stream;
  OZ_CUDA_CHECK(cudaEventRecord(h->ev_fork, s));
  OZ_CUDA_CHECK(cudaStreamWaitEvent(sb, h->ev_fork, 0));
  OZ_KERNEL_CHECK(ozk_split_int8(a_sl, w.pitch, amax, scr_a, m, k, a_ptr, lda, op_A == op_n, num_split, bits, s));
  cudaEvent_t ev_a = h->ev_block_split[0][0];
  OZ_CUDA_CHECK(cudaEventRecord(ev_a, s));  // also orders the products after everything queued on s before this call

  bool used[handle::kProductStreams] = {};
  // one CTA pair per tile (OZIMMU_B200_STREAMED_ONE_TILE, default 1): the panels' launches interleave tile by tile and
  // whatever carries the next panel (a NCCL broadcast kernel, the panel's split) gets SMs whenever a tile ends
  // A single panel (B arrives in one piece; only split(A) overlaps its transfer) is the ordinary persistent launch,
  // paced by the lockstep.
  const unsigned panel_flags =
      num_panels == 1 ? 0u
                      : OZK_FUSED_NO_LOCKSTEP |
                            (H::env_or("OZIMMU_B200_STREAMED_ONE_TILE", "1") != "0" ? OZK_FUSED_ONE_TILE_PER_PAIR : 0u);
  for (std::size_t p = 0; p < num_panels; p++) {
    const std::size_t j0 = col_edges[p], nj = col_edges[p + 1] - j0;
    if (nj == 0) continue;
    OZ_CUDA_CHECK(cudaStreamWaitEvent(sb, ready[p], 0));
    const double *src = (op_B == op_n) ? b_ptr + j0 * ldb : b_ptr + j0;
    OZ_KERNEL_CHECK(ozk_split_int8_block(b_sl, w.pitch, n, j0, bmax + j0, scr_b + j0, nj, k, src, ldb, op_B != op_n,
                                         num_split, bits, 1, sb));
    OZ_CUDA_CHECK(cudaEventRecord(h->ev_block_split[1][p], sb));
    const int r = static_cast<int>(p % 3);
    cudaStream_t sp = h->product_stream[r];
    used[r] = true;
    OZ_CUDA_CHECK(cudaStreamWaitEvent(sp, ev_a, 0));
    OZ_CUDA_CHECK(cudaStreamWaitEvent(sp, h->ev_block_split[1][p], 0));
    OZ_KERNEL_CHECK(ozk_gemm_i8_fused_block(m, nj, k, a_sl, m, 0, b_sl, n, j0, w.pitch, amax, bmax + j0, num_split, bits,
                                            *alpha, *beta, c_ptr + j0 * ldc, ldc, panel_flags, sp));
  }
  for (int r = 0; r < handle::kProductStreams; r++) {
    if (!used[r]) continue;
    OZ_CUDA_CHECK(cudaEventRecord(h->ev_product_tail[r], h->product_stream[r]));
    OZ_CUDA_CHECK(cudaStreamWaitEvent(s, h->ev_product_tail[r], 0));
  }
  mark_done(h, s);
  return 0;
}

// Not in the reference's header: its strided-batched interposers loop over gemm() (src/cublas.cu:380-406).
// Strides count elements (complex elements for element_kind == complx).
int mtk::ozimmu::gemm_strided_batched(handle_t h, const operation_t op_A, const operation_t op_B, const std::size_t m,
                                      const std::size_t n, const std::size_t k, const void *alpha,
                                      const void *const a_ptr, const std::size_t lda, const long long stride_a,
                                      const void *const b_ptr, const std::size_t ldb, const long long stride_b,
                                      const void *beta, void *const c_ptr, const std::size_t ldc,
                                      const long long stride_c, const std::size_t batch_count,
                                      const compute_mode_t compute_mode, const element_kind_t element_kind) {
  const std::size_t es = element_kind == real ? 1 : 2, esz = es * sizeof(double);
  int arg_error = 0;
  arg_error |= check_shape(op_A, m, k, lda, "A");
  arg_error |= check_shape(op_B, k, n, ldb, "B");
  arg_error |= check_shape(op_n, m, n, ldc, "C");
  arg_error |= check_alignment(a_ptr, esz, "A");
  arg_error |= check_alignment(b_ptr, esz, "B");
  arg_error |= check_alignment(c_ptr, esz, "C");
  // entries of C must not overlap (they are written concurrently); stride 0 is fine for the inputs
  if (batch_count > 1 && m > 0 && n > 0 &&
      static_cast<unsigned long long>(stride_c < 0 ? -stride_c : stride_c) < ldc * (n - 1) + m) {
    H::log_error("gemm_strided_batched: entries of C overlap (|stride_c| < ldc*(n-1)+m)");
    arg_error |= 1;
  }
  if (arg_error) return 1;
  const auto *a = static_cast<const double *>(a_ptr);
  const auto *b = static_cast<const double *>(b_ptr);
  auto *c = static_cast<double *>(c_ptr);
  // entries [e0, e0 + ne) in one grouped launch (int8 modes, forward-strided C), else entry by entry
  auto run = [&](std::size_t e0, std::size_t ne, compute_mode_t mode) -> int {
    if (H::is_int8_mode(mode) && stride_c >= 0) {
      gemm_int8(h, op_A, op_B, m, n, k, read_scalars(h, alpha, beta, element_kind),
                a + static_cast<long long>(e0) * stride_a * es, lda, stride_a, b + static_cast<long long>(e0) * stride_b * es,
                ldb, stride_b, c + static_cast<long long>(e0) * stride_c * es, ldc, stride_c, ne, H::num_split_of(mode),
                element_kind);
      return 0;
    }
    for (std::size_t e = e0; e < e0 + ne; e++) {
      const int rc = gemm(h, op_A, op_B, m, n, k, alpha, a + static_cast<long long>(e) * stride_a * es, lda,
                          b + static_cast<long long>(e) * stride_b * es, ldb, beta,
                          c + static_cast<long long>(e) * stride_c * es, ldc, mode, element_kind);
      if (rc) return rc;
    }
    return 0;
  };
  if (compute_mode != fp64_int8_auto || m == 0 || n == 0 || k == 0 || batch_count == 0)
    return run(0, batch_count, compute_mode == fp64_int8_auto ? fp64_int8_3 : compute_mode);
  // auto mode: the split count is chosen per entry (as the reference does by looping), but from ONE counter pass over
  // the whole batch and one blocking read; runs of consecutive entries with the same choice share a grouped launch
  std::vector<unsigned long long> totals;
  mantissa_loss_totals(h, op_A, op_B, m, n, k, a, lda, stride_a, b, ldb, stride_b, batch_count, element_kind, totals);
  constexpr int N = handle::mantissa_loss_counter_length;
  for (int i = 0; i < N; i++) h->last_loss_counters[i] = totals[(batch_count - 1) * N + i];
  std::vector<compute_mode_t> chosen(batch_count);
  for (std::size_t e = 0; e < batch_count; e++)
    chosen[e] = mode_for_loss(totals.data() + e * N, m, n, k, h->avg_mantissa_loss_threshold);
  for (std::size_t e0 = 0; e0 < batch_count;) {
    std::size_t e1 = e0 + 1;
    while (e1 < batch_count && chosen[e1] == chosen[e0]) e1++;
    H::log_info("AUTO selected mode = " + get_compute_mode_name_str(chosen[e0]) + " for batch entries " +
                std::to_string(e0) + ".." + std::to_string(e1 - 1));
    const int rc = run(e0, e1 - e0, chosen[e0]);
    if (rc) return rc;
    e0 = e1;
  }
  return 0;
}

int mtk::ozimmu::gemm_strided_batched(handle_t h, const operation_t op_A, const operation_t op_B, const std::size_t m,
                                      const std::size_t n, const std::size_t k, const double *alpha,
                                      const double *const a_ptr, const std::size_t lda, const long long stride_a,
                                      const double *const b_ptr, const std::size_t ldb, const long long stride_b,
                                      const double *beta, double *const c_ptr, const std::size_t ldc,
                                      const long long stride_c, const std::size_t batch_count,
                                      const compute_mode_t compute_mode) {
  return gemm_strided_batched(h, op_A, op_B, m, n, k, static_cast<const void *>(alpha), a_ptr, lda, stride_a, b_ptr, ldb,
                              stride_b, static_cast<const void *>(beta), c_ptr, ldc, stride_c, batch_count, compute_mode,
                              real);
}

void mtk::ozimmu::set_scalar_pointer_mode(handle_t handle, const bool on_device) { handle->scalars_on_device = on_device; }

// ===============================================================================================
// C spelling
// ===============================================================================================
extern "C" {

int ozimmu_create(ozimmu_handle_t *handle, int malloc_mode) {
  return guarded([&] {
    return create(reinterpret_cast<handle_t *>(handle), static_cast<malloc_mode_t>(malloc_mode));
  });
}
int ozimmu_destroy(ozimmu_handle_t handle) {
  return guarded([&] { return destroy(reinterpret_cast<handle_t>(handle)); });
}
int ozimmu_set_cuda_stream(ozimmu_handle_t handle, void *stream) {
  set_cuda_stream(reinterpret_cast<handle_t>(handle), static_cast<cudaStream_t>(stream));
  return 0;
}
int ozimmu_enable_profiling(ozimmu_handle_t handle) {
  enable_profiling(reinterpret_cast<handle_t>(handle));
  return 0;
}
int ozimmu_disable_profiling(ozimmu_handle_t handle) {
  disable_profiling(reinterpret_cast<handle_t>(handle));
  return 0;
}
int ozimmu_print_profiler_result(ozimmu_handle_t handle, const char *tag, int csv) {
  print_profiler_result(reinterpret_cast<handle_t>(handle), tag ? tag : "", csv != 0);
  return 0;
}
int ozimmu_clear_profiler_result(ozimmu_handle_t handle) {
  clear_profiler_result(reinterpret_cast<handle_t>(handle));
  return 0;
}
int ozimmu_set_auto_mantissa_loss_threshold(ozimmu_handle_t handle, double t) {
  set_auto_mantissa_loss_threashold(reinterpret_cast<handle_t>(handle), t);
  return 0;
}
double ozimmu_get_auto_mantissa_loss_threshold(ozimmu_handle_t handle) {
  return get_auto_mantissa_loss_threashold(reinterpret_cast<handle_t>(handle));
}

size_t ozimmu_reallocate_working_memory(ozimmu_handle_t handle, int op_a, int op_b, size_t m, size_t n, size_t k,
                                        int element_kind, int compute_mode) {
  size_t r = 0;
  guarded([&] {
    r = reallocate_working_memory(
        reinterpret_cast<handle_t>(handle),
        gemm_list_t{{static_cast<operation_t>(op_a), static_cast<operation_t>(op_b), m, n, k,
                     static_cast<element_kind_t>(element_kind), static_cast<compute_mode_t>(compute_mode)}});
    return 0;
  });
  return r;
}
size_t ozimmu_reallocate_working_memory_bytes(ozimmu_handle_t handle, size_t size_in_byte) {
  size_t r = 0;
  guarded([&] {
    r = reallocate_working_memory(reinterpret_cast<handle_t>(handle), size_in_byte);
    return 0;
  });
  return r;
}

int ozimmu_gemm(ozimmu_handle_t handle, int op_a, int op_b, size_t m, size_t n, size_t k, const void *alpha,
                const void *a, size_t lda, const void *b, size_t ldb, const void *beta, void *c, size_t ldc,
                int compute_mode, int element_kind) {
  if (handle == nullptr || alpha == nullptr || beta == nullptr || compute_mode < 0 ||
      compute_mode > OZIMMU_FP64_INT8_AUTO || (element_kind != OZIMMU_REAL && element_kind != OZIMMU_COMPLX))
    return 1;
  return guarded([&] {
    return gemm(reinterpret_cast<handle_t>(handle), static_cast<operation_t>(op_a != 0),
                static_cast<operation_t>(op_b != 0), m, n, k, alpha, a, lda, b, ldb, beta, c, ldc,
                static_cast<compute_mode_t>(compute_mode), static_cast<element_kind_t>(element_kind));
  });
}

int ozimmu_gemm_strided_batched(ozimmu_handle_t handle, int op_a, int op_b, size_t m, size_t n, size_t k,
                                const double *alpha, const double *a, size_t lda, long long stride_a,
                                const double *b, size_t ldb, long long stride_b, const double *beta, double *c,
                                size_t ldc, long long stride_c, size_t batch, int compute_mode) {
  if (handle == nullptr || alpha == nullptr || beta == nullptr || compute_mode < 0 ||
      compute_mode > OZIMMU_FP64_INT8_AUTO)
    return 1;
  return guarded([&] {
    return gemm_strided_batched(reinterpret_cast<handle_t>(handle), static_cast<operation_t>(op_a != 0),
                                static_cast<operation_t>(op_b != 0), m, n, k, alpha, a, lda, stride_a, b, ldb, stride_b,
                                beta, c, ldc, stride_c, batch, static_cast<compute_mode_t>(compute_mode));
  });
}

int ozimmu_gemm_strided_batched_ex(ozimmu_handle_t handle, int op_a, int op_b, size_t m, size_t n, size_t k,
                                   const void *alpha, const void *a, size_t lda, long long stride_a, const void *b,
                                   size_t ldb, long long stride_b, const void *beta, void *c, size_t ldc,
                                   long long stride_c, size_t batch, int compute_mode, int element_kind) {
  if (handle == nullptr || alpha == nullptr || beta == nullptr || compute_mode < 0 ||
      compute_mode > OZIMMU_FP64_INT8_AUTO || (element_kind != OZIMMU_REAL && element_kind != OZIMMU_COMPLX))
    return 1;
  return guarded([&] {
    return gemm_strided_batched(reinterpret_cast<handle_t>(handle), static_cast<operation_t>(op_a != 0),
                                static_cast<operation_t>(op_b != 0), m, n, k, alpha, a, lda, stride_a, b, ldb, stride_b,
                                beta, c, ldc, stride_c, batch, static_cast<compute_mode_t>(compute_mode),
                                static_cast<element_kind_t>(element_kind));
  });
}

int ozimmu_set_scalar_pointer_mode(ozimmu_handle_t handle, int on_device) {
  if (handle == nullptr) return 1;
  set_scalar_pointer_mode(reinterpret_cast<handle_t>(handle), on_device != 0);
  return 0;
}

int ozimmu_gemm_streamed_b(ozimmu_handle_t handle, int op_a, int op_b, size_t m, size_t n, size_t k,
                           const double *alpha, const double *a, size_t lda, const double *b, size_t ldb,
                           const double *beta, double *c, size_t ldc, int compute_mode, size_t num_panels,
                           const size_t *col_edges, void *const *ready_events) {
  if (handle == nullptr || alpha == nullptr || beta == nullptr || compute_mode < 0 ||
      compute_mode > OZIMMU_FP64_INT8_AUTO)
    return 1;
  return guarded([&] {
    return gemm_streamed_b(reinterpret_cast<handle_t>(handle), static_cast<operation_t>(op_a != 0),
                           static_cast<operation_t>(op_b != 0), m, n, k, alpha, a, lda, b, ldb, beta, c, ldc,
                           static_cast<compute_mode_t>(compute_mode), num_panels, col_edges,
                           reinterpret_cast<const cudaEvent_t *>(ready_events));
  });
}

int ozimmu_auto_mode_select(ozimmu_handle_t handle, int op_a, int op_b, size_t m, size_t n, size_t k,
                            const void *a, size_t lda, const void *b, size_t ldb, int element_kind,
                            double mantissa_loss_threshold, unsigned long long *counters16) {
  if (handle == nullptr) return -1;
  return guarded([&] {
    auto h = reinterpret_cast<handle_t>(handle);
    const compute_mode_t mode =
        auto_mode_select(h, static_cast<operation_t>(op_a != 0), static_cast<operation_t>(op_b != 0), m, n, k, a,
                         lda, b, ldb, static_cast<element_kind_t>(element_kind), mantissa_loss_threshold);
    if (counters16)
      std::memcpy(counters16, h->last_loss_counters, sizeof(unsigned long long) * handle::mantissa_loss_counter_length);
    return static_cast<int>(mode);
  });
}

const char *ozimmu_get_compute_mode_name_str(int compute_mode) {
  static thread_local std::string name;
  try {
    name = get_compute_mode_name_str(static_cast<compute_mode_t>(compute_mode));
  } catch (const std::exception &) {
    return nullptr;
  }
  return name.c_str();
}

uint32_t ozimmu_get_bits_per_int8(uint32_t k) { return ozk_bits_per_int8(k); }

}  // extern "C"
