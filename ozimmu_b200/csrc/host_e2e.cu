// host_e2e.cu -- ozimmu_gemm_host: the same DGEMM with HOST operands (the end-to-end entry a
// caller without device buffers uses; bench.py times it as "e2e").
//
// The reference has no such entry: an application does cudaMemcpy(A), cudaMemcpy(B),
// cublasDgemm (intercepted -> reference src/cublas.cu:280-295 -> src/gemm.cu:344-410),
// cudaMemcpy(C), strictly one after the other.  Here the call is PCIe-bound (8192^3: 1 GiB in ~ 20 ms
// against ~ 17 ms of products), so the work is cut into row blocks of op(A) and column blocks of op(B)
// that travel alternately, and every block of C is computed as soon as both of its operands are on the GPU:
//
//   h2d stream      : B0 | A0 | B1 | A1 | B2 | A2 | ...
//   compute stream  :      split(B0) split(A0) split(B1) split(A1) ...
//   product streams :                fused(A0 x B0) | fused(A0 x B1) | fused(A1 x [B0 B1]) | ...   (round-robin)
//   d2h stream      :                                 C(A0,B0) | C(A0,B1) | C(A1,[B0 B1]) | ...
//
// After a of A's blocks and b of B's have arrived a*b blocks of C are computable, so alternating the operands
// keeps the most work available per byte transferred; what remains after the last byte has landed is only the
// last block's row / column of C.  The fused launches rotate over several streams so that a launch back-fills
// the SMs the previous one leaves idle in its last (ragged) round of tiles.
//
// Results are bit-identical to the one-shot path: a block of C depends only on the same rows of op(A) and the
// same columns of op(B), and the split scales every row of A / column of B on its own (reference
// src/split.cu:193-242,277-282).
#include <algorithm>
#include <cstdlib>
#include <string>
#include <vector>

#include "host.hpp"
#include "oz_common.cuh"
#include "ozimmu_b200.h"
#include "sharded.hpp"

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

}  // namespace

// copy / split / product streams and events shared by the host-operand pipeline and gemm_streamed_b
void oz::host::ensure_pipeline_streams(mtk::ozimmu::handle *h) {
  if (h->h2d_stream) return;
  int lo = 0, hi = 0;
  OZ_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  OZ_CUDA_CHECK(cudaStreamCreateWithFlags(&h->h2d_stream, cudaStreamNonBlocking));
  OZ_CUDA_CHECK(cudaStreamCreateWithFlags(&h->d2h_stream, cudaStreamNonBlocking));
  // the splits are short and gate everything behind them: their CTAs go first when SMs free up
  OZ_CUDA_CHECK(cudaStreamCreateWithPriority(&h->compute_stream, cudaStreamNonBlocking, hi));
  for (auto &st : h->product_stream) OZ_CUDA_CHECK(cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, lo));
  auto make = [](cudaEvent_t &e) { OZ_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); };
  for (auto &row : h->ev_block_in)
    for (auto &e : row) make(e);
  for (auto &row : h->ev_block_split)
    for (auto &e : row) make(e);
  for (auto &e : h->ev_product_tail) make(e);
}

namespace {

// Block boundaries of one operand: blocks of `want` rows (rounded up to the kernel's 256-row tile, grown until at
// most kMaxBlocks are needed); with `taper` the last ~1.5 blocks are cut into halves, quarters, ... so that the
// blocks that arrive LAST -- nothing overlaps the products they enable -- carry little work.
std::vector<std::size_t> block_edges(std::size_t extent, std::size_t want, bool taper) {
  const std::size_t kmax = handle::kMaxBlocks;
  std::size_t e = (std::max<std::size_t>(want, 256) + 255) / 256 * 256;
  const std::size_t room = taper ? kmax - 2 : kmax;
  if ((extent + e - 1) / e > room) e = ((extent + room - 1) / room + 255) / 256 * 256;
  std::vector<std::size_t> edges{0};
  std::size_t at = 0;
  if (taper && e >= 512) {
    while (extent - at > e + e / 2) edges.push_back(at += e);
    // the rest (between e/2 and 1.5 e): halve until 256
    while (extent - at > 256 && edges.size() < kmax) {
      const std::size_t rest = extent - at;
      const std::size_t piece = std::max<std::size_t>(256, (rest / 2 + 255) / 256 * 256);
      if (piece >= rest) break;
      edges.push_back(at += piece);
    }
  } else {
    while (extent - at > e) edges.push_back(at += e);
  }
  edges.push_back(extent);
  return edges;
}

std::size_t env_size(const char *name, std::size_t fallback) {
  const char *e = std::getenv(name);
  if (e == nullptr || *e == 0) return fallback;
  const long v = std::atol(e);
  return v < 0 ? fallback : static_cast<std::size_t>(v);
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

// OZIMMU_B200_E2E_TRACE=1: time stamps (CUDA events with timing) of every stage of one host-operand call -- block
// landed, block split, rectangle multiplied, rectangle copied out -- printed as one line each, relative to the first
// copy.  A development aid: where does the time between the last byte in and the last byte out go?
struct E2eTrace {
  bool on = false;
  cudaEvent_t t0 = nullptr;
  std::vector<std::pair<std::string, cudaEvent_t>> marks;
  void begin(cudaStream_t s) {
    if (!on) return;
    OZ_CUDA_CHECK(cudaEventCreate(&t0));
    OZ_CUDA_CHECK(cudaEventRecord(t0, s));
  }
  void mark(const std::string &what, cudaStream_t s) {
    if (!on) return;
    cudaEvent_t e = nullptr;
    OZ_CUDA_CHECK(cudaEventCreate(&e));
    OZ_CUDA_CHECK(cudaEventRecord(e, s));
    marks.emplace_back(what, e);
  }
  void print() {
    if (!on) return;
    for (auto &m : marks) {
      float ms = 0;
      cudaEventElapsedTime(&ms, t0, m.second);
      std::printf("[ozIMMU e2e trace] %8.3f ms  %s\n", ms, m.first.c_str());
      cudaEventDestroy(m.second);
    }
    cudaEventDestroy(t0);
    std::fflush(stdout);
  }
};

// comm == nullptr (or a communicator of one rank): the single-GPU entry.  Otherwise this rank's row block of the
// sharded product (m = its rows of op(A) and C): B lives in the HOST memory of rank `src` only (b is ignored elsewhere);
// the owner uploads it block by block between its own blocks of A and broadcasts every block over NVLink as soon as it
// has landed, the other ranks receive the blocks on the communicator's stream while their own blocks of A arrive over
// their own PCIe links.  Each rank then runs exactly the single-GPU block pipeline, with "block of B has arrived"
// signalled by the broadcast instead of the H2D copy.
int gemm_host_impl(handle_t h, operation_t op_a, operation_t op_b, std::size_t m, std::size_t n, std::size_t k,
                   double alpha, const double *a, std::size_t lda, const double *b, std::size_t ldb, double beta,
                   double *c, std::size_t ldc, compute_mode_t mode, H::Comm *comm = nullptr, int src = 0) {
  const bool sharded = comm != nullptr && comm->size > 1;
  const bool owner = !sharded || comm->rank == src;   // B's host copy is here
  if ((op_a == op_n ? m : k) > lda || (op_b == op_n ? k : n) > ldb || m > ldc || (sharded && (src < 0 || src >= comm->size))) {
    H::log_error("ozimmu_gemm_host: leading dimension smaller than the matrix (or bad owner rank)");
    return 1;
  }
  if (n == 0 || (m == 0 && !sharded)) return 0;
  H::ensure_pipeline_streams(h);
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
    // one-shot: copy in (sharded: B from its owner, then one broadcast), run the device entry, copy out
    copy_matrix(da, a, lda, a_rows, a_cols, cudaMemcpyHostToDevice, sc);
    if (owner) copy_matrix(db, b, ldb, b_rows, b_cols, cudaMemcpyHostToDevice, sc);
    if (sharded) H::comm_broadcast_f64(comm, db, b_cols == 0 ? 0 : ldb * (b_cols - 1) + b_rows, src, sc);
    if (m == 0) {
      OZ_CUDA_CHECK(cudaStreamSynchronize(sc));
      return 0;
    }
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
  // Block edges (multiples of 256, the kernel's tile): OZIMMU_B200_E2E_PANEL = columns of op(B) / C per block,
  // OZIMMU_B200_E2E_ROWBLOCK = rows of op(A) / C per block; 0 = the whole operand in one piece (ROWBLOCK=0:
  // column panels only).  Read per call.  768 measured best at 8192^3
  // (profiles/r1_e2e_block_sweep.txt).
  const std::size_t want_cols = env_size("OZIMMU_B200_E2E_PANEL", 768);
  const std::size_t want_rows = env_size("OZIMMU_B200_E2E_ROWBLOCK", 768);
  // OZIMMU_B200_E2E_TAPER=1 halves the last blocks (less work behind the last byte); measured +-0 at the default
  // edge of 768, +0.5 ms better at 1024 (profiles/r1_e2e_block_sweep.txt): off by default.
  const bool taper = env_size("OZIMMU_B200_E2E_TAPER", 0) != 0;
  // sharded: a block of B is one broadcast, so it must be contiguous -- column panels of an op_n B; an op_t B (its
  // column panels are row ranges of the stored matrix) travels in one piece
  const bool b_whole = want_cols == 0 || (sharded && op_b != op_n);
  const std::vector<std::size_t> be = block_edges(n, b_whole ? n : want_cols, taper && !b_whole);
  const std::vector<std::size_t> ae = m == 0 ? std::vector<std::size_t>{0, 0}
                                             : block_edges(m, want_rows == 0 ? m : want_rows, taper && want_rows != 0);
  const std::size_t nbb = be.size() - 1, nab = m == 0 ? 0 : ae.size() - 1;

  const H::WorkspaceLayout w = H::workspace_layout(m, n, k, s);
  reallocate_working_memory(h, w.total);
  if (h->has_pending) OZ_CUDA_CHECK(cudaStreamWaitEvent(sc, h->ev_done, 0));
  char *ws = static_cast<char *>(h->working_memory_ptr);
  double *amax = reinterpret_cast<double *>(ws + w.off_amax);
  double *bmax = reinterpret_cast<double *>(ws + w.off_bmax);
  auto *scr_a = reinterpret_cast<std::uint32_t *>(ws + w.off_scr_a);
  auto *scr_b = reinterpret_cast<std::uint32_t *>(ws + w.off_scr_b);
  auto *a_sl = reinterpret_cast<std::int8_t *>(ws + w.off_a_slices);
  auto *b_sl = reinterpret_cast<std::int8_t *>(ws + w.off_b_slices);

  E2eTrace trace;
  trace.on = env_size("OZIMMU_B200_E2E_TRACE", 0) != 0;
  trace.begin(sin);

  auto copy_a_block = [&](std::size_t i) {
    const std::size_t i0 = ae[i], mi = ae[i + 1] - i0;
    if (op_a == op_n) {  // m x k column-major: rows i0..i0+mi of every column
      copy_matrix(da + i0, a + i0, lda, mi, k, cudaMemcpyHostToDevice, sin);
    } else {             // k x m column-major: a row block of op(A) is a contiguous column panel
      copy_matrix(da + i0 * lda, a + i0 * lda, lda, k, mi, cudaMemcpyHostToDevice, sin);
    }
    OZ_CUDA_CHECK(cudaEventRecord(h->ev_block_in[0][i], sin));
    trace.mark("A" + std::to_string(i) + " landed", sin);
  };
  // the element range of db that holds block j of op(B) when blocks are broadcast (contiguous by construction)
  auto b_block_range = [&](std::size_t j, std::size_t &first, std::size_t &count) {
    const std::size_t j0 = be[j], nj = be[j + 1] - j0;
    if (op_b == op_n) {
      first = j0 * ldb;
      count = (j + 1 == nbb) ? (nj - 1) * ldb + k : nj * ldb;
    } else {   // one block: the whole stored n x k matrix
      first = 0;
      count = ldb * (k - 1) + n;
    }
  };
  auto copy_b_block = [&](std::size_t j) {
    const std::size_t j0 = be[j], nj = be[j + 1] - j0;
    if (owner) {
      if (op_b == op_n) {  // k x n column-major: a column panel is contiguous
        copy_matrix(db + j0 * ldb, b + j0 * ldb, ldb, k, nj, cudaMemcpyHostToDevice, sin);
      } else {             // n x k column-major: rows j0..j0+nj of every column
        copy_matrix(db + j0, b + j0, ldb, nj, k, cudaMemcpyHostToDevice, sin);
      }
    }
    if (beta != 0 && m != 0) copy_matrix(dc + j0 * ldc, c + j0 * ldc, ldc, m, nj, cudaMemcpyHostToDevice, sin);
    OZ_CUDA_CHECK(cudaEventRecord(h->ev_block_in[1][j], sin));
    trace.mark("B" + std::to_string(j) + " landed (H2D)", sin);
    if (sharded) {
      // the owner forwards the block as soon as it is on its GPU; everybody else's copy arrives here.  All ranks
      // issue the broadcasts in block order on their communicator stream.
      std::size_t first, count;
      b_block_range(j, first, count);
      OZ_CUDA_CHECK(cudaStreamWaitEvent(comm->stream, h->ev_block_in[1][j], 0));  // owner: H2D done; others: C panel (beta)
      H::comm_broadcast_f64(comm, db + first, count, src, comm->stream);
      if (!owner) OZ_CUDA_CHECK(cudaEventRecord(h->ev_block_in[1][j], comm->stream));
    }
  };
  auto split_a_block = [&](std::size_t i) {
    const std::size_t i0 = ae[i], mi = ae[i + 1] - i0;
    OZ_CUDA_CHECK(cudaStreamWaitEvent(sc, h->ev_block_in[0][i], 0));
    const double *src = (op_a == op_n) ? da + i0 : da + i0 * lda;
    OZ_KERNEL_CHECK(ozk_split_int8_block(a_sl, w.pitch, m, i0, amax + i0, scr_a + i0, mi, k, src, lda, op_a == op_n, s,
                                         bits, 1, sc));
    OZ_CUDA_CHECK(cudaEventRecord(h->ev_block_split[0][i], sc));
    trace.mark("A" + std::to_string(i) + " split", sc);
  };
  auto split_b_block = [&](std::size_t j) {
    const std::size_t j0 = be[j], nj = be[j + 1] - j0;
    OZ_CUDA_CHECK(cudaStreamWaitEvent(sc, h->ev_block_in[1][j], 0));
    const double *src = (op_b == op_n) ? db + j0 * ldb : db + j0;
    OZ_KERNEL_CHECK(ozk_split_int8_block(b_sl, w.pitch, n, j0, bmax + j0, scr_b + j0, nj, k, src, ldb, op_b != op_n, s,
                                         bits, 1, sc));
    OZ_CUDA_CHECK(cudaEventRecord(h->ev_block_split[1][j], sc));
    trace.mark("B" + std::to_string(j) + " split", sc);
  };
  // C[i0 : i0+mi, j0 : j0+nj] once `ready` (the split that completed its operands; the compute stream runs the
  // splits in arrival order, so it implies every earlier one) has fired.  OZIMMU_B200_E2E_ONE_TILE (default 1): the
  // launches are non-persistent, one CTA pair per tile, so the hardware CTA scheduler interleaves the tiles of
  // consecutive launches and the high-priority block splits get SMs whenever a tile ends.
  // OZIMMU_B200_E2E_RECT_TILES=t (default 0 = off): a rectangle of more than t tiles is cut along its long side into
  // launches of at most t tiles, each with its own copy-out, so that the D2H of a rectangle starts before its last
  // tile is done.
  const bool one_tile = env_size("OZIMMU_B200_E2E_ONE_TILE", 1) != 0;
  const std::size_t rect_tiles = env_size("OZIMMU_B200_E2E_RECT_TILES", 0);
  const unsigned nstreams = static_cast<unsigned>(
      std::min<std::size_t>(std::max<std::size_t>(env_size("OZIMMU_B200_E2E_STREAMS", 3), 1), handle::kProductStreams));
  const unsigned fused_flags = OZK_FUSED_NO_LOCKSTEP | (one_tile ? OZK_FUSED_ONE_TILE_PER_PAIR : 0u);
  unsigned rects = 0;
  bool used[handle::kProductStreams] = {};
  auto product_piece = [&](std::size_t i0, std::size_t mi, std::size_t j0, std::size_t nj, cudaEvent_t ready) {
    const unsigned r = rects % nstreams;
    cudaStream_t sp = h->product_stream[r];
    used[r] = true;
    while (h->ev_rect_out.size() <= rects) {
      cudaEvent_t e = nullptr;
      OZ_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      h->ev_rect_out.push_back(e);
    }
    OZ_CUDA_CHECK(cudaStreamWaitEvent(sp, ready, 0));
    double *dblk = dc + j0 * ldc + i0;
    OZ_KERNEL_CHECK(ozk_gemm_i8_fused_block(mi, nj, k, a_sl, m, i0, b_sl, n, j0, w.pitch, amax + i0, bmax + j0, s, bits,
                                            alpha, beta, dblk, ldc, fused_flags, sp));
    OZ_CUDA_CHECK(cudaEventRecord(h->ev_rect_out[rects], sp));
    const std::string name = "C[" + std::to_string(i0) + "+" + std::to_string(mi) + ", " + std::to_string(j0) + "+" + std::to_string(nj) + "]";
    trace.mark(name + " multiplied", sp);
    OZ_CUDA_CHECK(cudaStreamWaitEvent(sout, h->ev_rect_out[rects], 0));
    copy_matrix(c + j0 * ldc + i0, dblk, ldc, mi, nj, cudaMemcpyDeviceToHost, sout);
    trace.mark(name + " copied out", sout);
    rects++;
  };
  // `pieces` > 1 (the last rectangles of the call): cut along the long side into that many launches, so that the
  // copy-out of the first pieces overlaps the products of the later ones and only one piece's D2H trails the last tile
  auto product_rect = [&](std::size_t i0, std::size_t mi, std::size_t j0, std::size_t nj, cudaEvent_t ready,
                          std::size_t pieces = 1) {
    if (mi == 0 || nj == 0) return;
    const std::size_t tm = (mi + 255) / 256, tn = (nj + 255) / 256;
    std::size_t limit = rect_tiles;
    if (pieces > 1) {
      const std::size_t per_piece = std::max<std::size_t>(1, (tm * tn + pieces - 1) / pieces);
      limit = limit == 0 ? per_piece : std::min(limit, per_piece);
    }
    if (limit == 0 || tm * tn <= limit) return product_piece(i0, mi, j0, nj, ready);
    if (tn >= tm) {   // wide: pieces of whole tile columns
      const std::size_t step = std::max<std::size_t>(1, limit / tm) * 256;
      for (std::size_t j = 0; j < nj; j += step) product_piece(i0, mi, j0 + j, std::min(step, nj - j), ready);
    } else {          // tall: pieces of whole tile rows
      const std::size_t step = std::max<std::size_t>(1, limit / tn) * 256;
      for (std::size_t i = 0; i < mi; i += step) product_piece(i0 + i, std::min(step, mi - i), j0, nj, ready);
    }
  };

  // arrival order: A0, B0, A1, B1, ... (the longer operand's remaining blocks follow at the end); queue every
  // copy first so the H2D engine never waits for the host.  A first, so that the LAST rectangle of C is a column block
  // (op(A) complete x the last block of op(B)): contiguous in a column-major C, it leaves at the full D2H rate, where a
  // row block of 768 x 8-byte segments reaches 36 GB/s while the GPU is busy (profiles/r2_ubench_pcie_2d.txt).
  // OZIMMU_B200_E2E_A_FIRST=0 restores B first: 25.65 -> 25.25 ms at 8192^3 (profiles/r2_e2e_tail_sweep.txt).
  // OZIMMU_B200_E2E_TAIL_PIECES=p cuts the last two rectangles into p launches each, so that only a fraction of their
  // copy-out trails the last tile; it pays with block edges of 1024 (p = 4: 25.15 ms) but not at the default edge of 768
  // (p = 1 / 2 / 4: 25.3 / 25.3 / 25.6 ms): off (1) by default.
  // Sharded: the owner uploads ALL of B first and forwards every block as it lands, then its own A.  While B is in
  // flight no rank has much to multiply yet, so NCCL's broadcast kernels find free SMs at once on every GPU; with B
  // interleaved behind A on the owner, each of its broadcasts had to wait for SMs on eight busy GPUs and the step grew
  // with the rank count (8 GPUs: 72 ms, profiles/r2_bench_8gpu_call8.json).  The other ranks receive B at the rate
  // their own blocks of A arrive over their own PCIe links: blocks alternate as on a single GPU.
  const bool a_first = !sharded && env_size("OZIMMU_B200_E2E_A_FIRST", 1) != 0;
  const std::size_t tail_pieces = std::max<std::size_t>(1, env_size("OZIMMU_B200_E2E_TAIL_PIECES", 1));
  struct Arrival { int which; std::size_t idx; };
  std::vector<Arrival> order;
  if (sharded && owner) {
    for (std::size_t ib = 0; ib < nbb; ib++) order.push_back({1, ib});
    for (std::size_t ia = 0; ia < nab; ia++) order.push_back({0, ia});
  } else {
    for (std::size_t ia = 0, ib = 0; ia < nab || ib < nbb;) {
      const bool take_b = a_first ? (ib < nbb && (ib < ia || ia >= nab)) : (ib < nbb && (ib <= ia || ia >= nab));
      if (take_b) order.push_back({1, ib++});
      else order.push_back({0, ia++});
    }
  }

  if (sharded) {
    // the communicator stream joins this call: the previous call's last broadcast is already ordered before it
    OZ_CUDA_CHECK(cudaEventRecord(comm->ev_begin, sc));
    OZ_CUDA_CHECK(cudaStreamWaitEvent(comm->stream, comm->ev_begin, 0));
  }
  for (const Arrival &x : order) x.which ? copy_b_block(x.idx) : copy_a_block(x.idx);

  std::size_t have_a = 0, have_b = 0;  // blocks split so far
  for (std::size_t o = 0; o < order.size(); o++) {
    const Arrival &x = order[o];
    const std::size_t pieces = o + 2 >= order.size() ? tail_pieces : 1;
    if (x.which) {
      split_b_block(x.idx);
      product_rect(0, ae[have_a], be[x.idx], be[x.idx + 1] - be[x.idx], h->ev_block_split[1][x.idx], pieces);
      have_b++;
    } else {
      split_a_block(x.idx);
      product_rect(ae[x.idx], ae[x.idx + 1] - ae[x.idx], 0, be[have_b], h->ev_block_split[0][x.idx], pieces);
      have_a++;
    }
  }
  // the workspace is free again once every product stream has drained
  for (int r = 0; r < handle::kProductStreams; r++) {
    if (!used[r]) continue;
    OZ_CUDA_CHECK(cudaEventRecord(h->ev_product_tail[r], h->product_stream[r]));
    OZ_CUDA_CHECK(cudaStreamWaitEvent(sc, h->ev_product_tail[r], 0));
  }
  if (sharded) {
    OZ_CUDA_CHECK(cudaEventRecord(comm->ev_end, comm->stream));   // the owner's sends read db until here
    OZ_CUDA_CHECK(cudaStreamWaitEvent(sc, comm->ev_end, 0));
  }
  OZ_CUDA_CHECK(cudaEventRecord(h->ev_done, sc));
  h->has_pending = true;
  h->last_stream = sc;
  OZ_CUDA_CHECK(cudaStreamSynchronize(sout));
  OZ_CUDA_CHECK(cudaStreamSynchronize(sc));
  trace.print();
  return 0;
}

}  // namespace

// Row block of a sharded DGEMM with HOST operands (see gemm_host_impl): a_block / c_block are this rank's rows, b is
// read on rank src_rank only.  Collective over the communicator; returns when this rank's block of C is complete.
extern "C" int ozimmu_gemm_sharded_host(ozimmu_handle_t handle, ozimmu_comm_t comm, int op_a, int op_b, size_t m_local,
                                        size_t n, size_t k, const double *alpha, const double *a_block, size_t lda,
                                        const double *b, size_t ldb, const double *beta, double *c_block, size_t ldc,
                                        int compute_mode, int src_rank) {
  if (handle == nullptr || alpha == nullptr || beta == nullptr || compute_mode < 0 ||
      compute_mode > OZIMMU_FP64_INT8_AUTO)
    return 1;
  try {
    return gemm_host_impl(reinterpret_cast<handle_t>(handle), static_cast<operation_t>(op_a != 0),
                          static_cast<operation_t>(op_b != 0), m_local, n, k, *alpha, a_block, lda, b, ldb, *beta, c_block,
                          ldc, static_cast<compute_mode_t>(compute_mode), reinterpret_cast<H::Comm *>(comm), src_rank);
  } catch (const std::exception &e) {
    cudaDeviceSynchronize();
    H::log_error(e.what());
    return -1;
  }
}

// Diagnostic: the block boundaries ozimmu_gemm_host uses for one operand (CPU-testable host logic).
extern "C" size_t ozimmu_host_block_edges(size_t extent, size_t want, int taper, size_t *edges, size_t capacity) {
  if (extent == 0) return 0;
  const std::vector<std::size_t> e = block_edges(extent, want == 0 ? extent : want, taper != 0 && want != 0);
  for (std::size_t i = 0; i < e.size() && i < capacity; i++) edges[i] = e[i];
  return e.size();
}

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
    // copies that reference the caller's host buffers may still be queued on the internal streams: drain them before
    // the caller gets its buffers back
    cudaDeviceSynchronize();
    H::log_error(e.what());
    return -1;
  }
}
