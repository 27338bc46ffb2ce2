// sharded.cpp -- the multi-GPU DGEMM behind the C-ABI: communicator management (NCCL, resolved at run time) and
// the device-operand entry ozimmu_gemm_sharded.  The host-operand entry (ozimmu_gemm_sharded_host) shares the block
// pipeline of ozimmu_gemm_host and lives in host_e2e.cu.
//
// Not in the reference (it has no multi-GPU path); SURVEY 8e / BASELINE.json config 4: row blocks of A and C per GPU,
// one broadcast of B from its owner, no reduction (K is never split) => every rank's block is bit-identical to the
// single-GPU result.  Default (max_panels <= 1): ONE broadcast on the communicator's own high-priority stream with
// split(A) running meanwhile, then split(B) and the ordinary persistent product launch.  max_panels > 1: B travels in
// column panels and each panel of C is computed as soon as its columns have landed (gemm_streamed_b, one CTA pair per
// tile at low stream priority so that NCCL's kernels get SMs whenever a tile ends) -- measured slower on 2, 4 and 8
// GPUs, kept as the caller's choice.
#include <dlfcn.h>
#include <link.h>

#include <cstring>
#include <mutex>
#include <string>

#include <nccl.h>  // types and enums only: every NCCL entry point is resolved with dlsym at run time

#include "host.hpp"
#include "ozimmu_b200.h"
#include "sharded.hpp"

using namespace mtk::ozimmu;
namespace H = oz::host;

namespace {

// ---- NCCL, looked up at run time: the copy the process has already loaded (e.g. PyTorch's bundled one -- two NCCL
// instances in one process would each open their own transports), else libnccl.so.2 from the loader path -------------
struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommCount)(const ncclComm_t, int *) = nullptr;
  ncclResult_t (*CommUserRank)(const ncclComm_t, int *) = nullptr;
  ncclResult_t (*CommCuDevice)(const ncclComm_t, int *) = nullptr;
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

int find_loaded_nccl(struct dl_phdr_info *info, size_t, void *data) {
  if (info->dlpi_name && std::strstr(info->dlpi_name, "libnccl.so")) {
    *static_cast<std::string *>(data) = info->dlpi_name;
    return 1;
  }
  return 0;
}

const NcclApi &nccl() {
  static const NcclApi api = [] {
    NcclApi a;
    std::string loaded;
    dl_iterate_phdr(find_loaded_nccl, &loaded);
    void *lib = nullptr;
    if (const char *forced = std::getenv("OZIMMU_B200_NCCL_LIB")) lib = dlopen(forced, RTLD_NOW | RTLD_LOCAL);
    if (!lib && !loaded.empty()) lib = dlopen(loaded.c_str(), RTLD_NOW | RTLD_LOCAL);
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
      if (lib) break;
      lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
    }
    if (!lib) return a;
    auto sym = [&](const char *n) { return dlsym(lib, n); };
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
    a.CommCount = reinterpret_cast<decltype(a.CommCount)>(sym("ncclCommCount"));
    a.CommUserRank = reinterpret_cast<decltype(a.CommUserRank)>(sym("ncclCommUserRank"));
    a.CommCuDevice = reinterpret_cast<decltype(a.CommCuDevice)>(sym("ncclCommCuDevice"));
    a.Broadcast = reinterpret_cast<decltype(a.Broadcast)>(sym("ncclBroadcast"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.CommCount && a.CommUserRank && a.Broadcast;
    return a;
  }();
  return api;
}

void nccl_check(ncclResult_t r, const char *what) {
  if (r == ncclSuccess) return;
  const char *msg = nccl().GetErrorString ? nccl().GetErrorString(r) : "?";
  throw std::runtime_error(std::string("ozIMMU: ") + what + " failed: " + msg);
}

void require_nccl() {
  if (!nccl().ok) throw std::runtime_error("ozIMMU: NCCL (libnccl.so.2) is not available in this process");
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

H::Comm *finish_comm(ncclComm_t nc, bool owned) {
  auto *c = new H::Comm;
  c->nccl = nc;
  c->owned = owned;
  nccl_check(nccl().CommCount(nc, &c->size), "ncclCommCount");
  nccl_check(nccl().CommUserRank(nc, &c->rank), "ncclCommUserRank");
  OZ_CUDA_CHECK(cudaGetDevice(&c->device));
  int lo = 0, hi = 0;
  OZ_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  // the collectives feed everything else: their kernels go first when SMs free up
  OZ_CUDA_CHECK(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, hi));
  OZ_CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_begin, cudaEventDisableTiming));
  OZ_CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_end, cudaEventDisableTiming));
  return c;
}

}  // namespace

void oz::host::comm_broadcast_f64(Comm *c, double *buf, std::size_t count, int root, cudaStream_t s) {
  if (count == 0) return;
  nccl_check(nccl().Broadcast(buf, buf, count, ncclDouble, root, static_cast<ncclComm_t>(c->nccl), s), "ncclBroadcast");
}

cudaEvent_t oz::host::comm_panel_event(Comm *c, std::size_t p) {
  while (c->ev_panel.size() <= p) {
    cudaEvent_t e = nullptr;
    OZ_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c->ev_panel.push_back(e);
  }
  return c->ev_panel[p];
}

namespace {

// Column panels of op(B) for the broadcast pipeline: equal widths, multiples of 256 (the kernel's tile and the block-wise
// split's granularity), at least min_width wide, at most max_panels of them.
std::vector<std::size_t> panel_edges(std::size_t n, std::size_t max_panels, std::size_t min_width) {
  std::size_t panels = std::max<std::size_t>(1, std::min(max_panels, n / std::max<std::size_t>(1, min_width)));
  panels = std::min<std::size_t>(panels, handle::kMaxBlocks);
  std::size_t w = (n + panels - 1) / panels;
  w = (w + 255) / 256 * 256;
  std::vector<std::size_t> edges{0};
  for (std::size_t j = w; j < n; j += w) edges.push_back(j);
  edges.push_back(n);
  return edges;
}

int gemm_sharded_impl(handle_t h, H::Comm *c, operation_t op_a, operation_t op_b, std::size_t m_local, std::size_t n,
                      std::size_t k, const double *alpha, const double *a, std::size_t lda, double *b, std::size_t ldb,
                      const double *beta, double *cc, std::size_t ldc, compute_mode_t mode, int src, unsigned max_panels) {
  if (c == nullptr || c->size == 1)
    return gemm(h, op_a, op_b, m_local, n, k, alpha, a, lda, b, ldb, beta, cc, ldc, mode, real);
  if (src < 0 || src >= c->size) {
    H::log_error("gemm_sharded: the owner rank of B is outside the communicator");
    return 1;
  }
  if ((op_b == op_n ? k : n) > ldb) {
    H::log_error("gemm_sharded: ldb smaller than the rows of B");
    return 1;
  }
  cudaStream_t s = h->cuda_stream;
  // B (the owner's content, the other ranks' receive buffer) is in the caller's stream order up to here
  OZ_CUDA_CHECK(cudaEventRecord(c->ev_begin, s));
  OZ_CUDA_CHECK(cudaStreamWaitEvent(c->stream, c->ev_begin, 0));
  const std::size_t b_cols = (op_b == op_n) ? n : k, b_rows = (op_b == op_n) ? k : n;
  const std::size_t total = b_cols == 0 ? 0 : ldb * (b_cols - 1) + b_rows;
  const std::size_t min_panel = 1024;
  const bool pipelined = max_panels > 1 && op_b == op_n && H::is_int8_mode(mode) && k > 0 && n >= 2 * min_panel &&
                         !h->profiler.enabled;
  if (!pipelined) {
    // one broadcast (op_t B: a column panel of op(B) is not contiguous; auto mode looks at all of B; max_panels <= 1)
    H::comm_broadcast_f64(c, b, total, src, c->stream);
    OZ_CUDA_CHECK(cudaEventRecord(c->ev_end, c->stream));
    int rc = 0;
    if (m_local != 0 && H::is_int8_mode(mode) && k > 0 && n > 0 && !h->profiler.enabled) {
      // B as ONE panel guarded by the broadcast's event: split(A) runs while B is on the wire, then the ordinary
      // single product launch
      const std::size_t edges[2] = {0, n};
      const cudaEvent_t ready = (c->rank == src) ? c->ev_begin : c->ev_end;
      rc = gemm_streamed_b(h, op_a, op_b, m_local, n, k, alpha, a, lda, b, ldb, beta, cc, ldc, mode, 1, edges, &ready);
      OZ_CUDA_CHECK(cudaStreamWaitEvent(s, c->ev_end, 0));
      return rc;
    }
    OZ_CUDA_CHECK(cudaStreamWaitEvent(s, c->ev_end, 0));
    if (m_local == 0) return 0;
    return gemm(h, op_a, op_b, m_local, n, k, alpha, a, lda, b, ldb, beta, cc, ldc, mode, real);
  }
  const std::vector<std::size_t> edges = panel_edges(n, max_panels, min_panel);
  const std::size_t np = edges.size() - 1;
  std::vector<cudaEvent_t> ready(np);
  for (std::size_t p = 0; p < np; p++) {
    const std::size_t j0 = edges[p], nj = edges[p + 1] - j0;
    const std::size_t count = (p + 1 == np) ? (nj - 1) * ldb + k : nj * ldb;
    H::comm_broadcast_f64(c, b + j0 * ldb, count, src, c->stream);
    // the owner already holds every panel: its products need not wait for its own sends
    ready[p] = (c->rank == src) ? c->ev_begin : H::comm_panel_event(c, p);
    if (c->rank != src) OZ_CUDA_CHECK(cudaEventRecord(ready[p], c->stream));
  }
  OZ_CUDA_CHECK(cudaEventRecord(c->ev_end, c->stream));
  int rc = 0;
  if (m_local != 0)
    rc = gemm_streamed_b(h, op_a, op_b, m_local, n, k, alpha, a, lda, b, ldb, beta, cc, ldc, mode, np, edges.data(),
                         ready.data());
  // B must stay untouched until the owner's sends are done / is complete on the receivers when the call's work is
  OZ_CUDA_CHECK(cudaStreamWaitEvent(s, c->ev_end, 0));
  return rc;
}

}  // namespace

extern "C" {

int ozimmu_comm_unique_id(void *id128) {
  if (id128 == nullptr) return 1;
  return guarded([&] {
    require_nccl();
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    nccl_check(nccl().GetUniqueId(&id), "ncclGetUniqueId");
    std::memcpy(id128, &id, sizeof(id));
    return 0;
  });
}

int ozimmu_comm_create(ozimmu_comm_t *comm, int nranks, int rank, const void *id128) {
  if (comm == nullptr || id128 == nullptr || nranks < 1 || rank < 0 || rank >= nranks) return 1;
  return guarded([&] {
    require_nccl();
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    ncclComm_t nc = nullptr;
    nccl_check(nccl().CommInitRank(&nc, nranks, id, rank), "ncclCommInitRank");
    *comm = reinterpret_cast<ozimmu_comm_t>(finish_comm(nc, true));
    return 0;
  });
}

int ozimmu_comm_adopt(ozimmu_comm_t *comm, void *nccl_comm) {
  if (comm == nullptr || nccl_comm == nullptr) return 1;
  return guarded([&] {
    require_nccl();
    *comm = reinterpret_cast<ozimmu_comm_t>(finish_comm(static_cast<ncclComm_t>(nccl_comm), false));
    return 0;
  });
}

int ozimmu_comm_destroy(ozimmu_comm_t comm) {
  if (comm == nullptr) return 0;
  return guarded([&] {
    auto *c = reinterpret_cast<H::Comm *>(comm);
    cudaStreamSynchronize(c->stream);
    if (c->owned && c->nccl) nccl().CommDestroy(static_cast<ncclComm_t>(c->nccl));
    cudaStreamDestroy(c->stream);
    cudaEventDestroy(c->ev_begin);
    cudaEventDestroy(c->ev_end);
    for (cudaEvent_t e : c->ev_panel) cudaEventDestroy(e);
    delete c;
    return 0;
  });
}

int ozimmu_comm_rank(ozimmu_comm_t comm) { return comm ? reinterpret_cast<H::Comm *>(comm)->rank : 0; }
int ozimmu_comm_size(ozimmu_comm_t comm) { return comm ? reinterpret_cast<H::Comm *>(comm)->size : 1; }

size_t ozimmu_sharded_panel_edges(size_t n, size_t max_panels, size_t *edges, size_t capacity) {
  if (n == 0) return 0;
  const std::vector<std::size_t> e = panel_edges(n, max_panels, 1024);
  for (std::size_t i = 0; i < e.size() && i < capacity; i++) edges[i] = e[i];
  return e.size();
}

void ozimmu_row_block(size_t m, int nranks, int rank, size_t *row0, size_t *rows) {
  const std::size_t bm = nranks > 0 ? (m + static_cast<std::size_t>(nranks) - 1) / static_cast<std::size_t>(nranks) : m;
  const std::size_t r0 = std::min(m, static_cast<std::size_t>(rank < 0 ? 0 : rank) * bm);
  if (row0) *row0 = r0;
  if (rows) *rows = std::min(bm, m - r0);
}

int ozimmu_gemm_sharded(ozimmu_handle_t handle, ozimmu_comm_t comm, int op_a, int op_b, size_t m_local, size_t n, size_t k,
                        const double *alpha, const double *a_block, size_t lda, double *b, size_t ldb, const double *beta,
                        double *c_block, size_t ldc, int compute_mode, int src_rank, unsigned max_panels) {
  if (handle == nullptr || alpha == nullptr || beta == nullptr || compute_mode < 0 || compute_mode > OZIMMU_FP64_INT8_AUTO)
    return 1;
  return guarded([&] {
    return gemm_sharded_impl(reinterpret_cast<handle_t>(handle), reinterpret_cast<H::Comm *>(comm),
                             static_cast<operation_t>(op_a != 0), static_cast<operation_t>(op_b != 0), m_local, n, k, alpha,
                             a_block, lda, b, ldb, beta, c_block, ldc, static_cast<compute_mode_t>(compute_mode), src_rank,
                             max_panels);
  });
}

}  // extern "C"
