// host_util.cu -- env/log helpers, real-cuBLAS symbol lookup, stage profiler, workspace layout.
// Functional equivalents of reference src/utils.hpp:77-141 (env, log, dlsym), the
// cutf time_breakdown profiler the reference keeps in its handle (src/handle.hpp:16) and
// the workspace arithmetic of src/handle.cu:95-144 / src/gemm.cu:359-379.
#include <dlfcn.h>
#include <link.h>

#include <algorithm>
#include <cstring>
#include <mutex>

#include "host.hpp"
#include "oz_common.cuh"
#include "ozimmu_b200.h"

namespace oz {

unsigned long long g_launch_count = 0;

namespace host {

// reference src/utils.hpp:88-96: set and != "0"  -> on; unset -> default
bool env_enabled(const char *name, bool default_value) {
  const char *v = std::getenv(name);
  if (v == nullptr) return default_value;
  return std::strcmp(v, "0") != 0;
}

std::string env_or(const char *name, const std::string &fallback) {
  const char *v = std::getenv(name);
  return v ? std::string(v) : fallback;
}

void log_info(const std::string &msg) {
  if (!env_enabled("OZIMMU_INFO", false)) return;
  std::fprintf(stdout, "[ozIMMU LOG] %s\n", msg.c_str());
  std::fflush(stdout);
}

void log_error(const std::string &msg) {
  if (!env_enabled("OZIMMU_ERROR", true)) return;
  std::fprintf(stdout, "[ozIMMU ERROR] %s\n", msg.c_str());
  std::fflush(stdout);
}

void cuda_check(cudaError_t e, const char *what, const char *file, int line) {
  if (e == cudaSuccess) return;
  throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e) + " [" + what + "] at " +
                           file + ":" + std::to_string(line));
}

// ---- real cuBLAS ------------------------------------------------------------------------------
namespace {
int find_loaded_cublas(struct dl_phdr_info *info, size_t, void *data) {
  if (info->dlpi_name && std::strstr(info->dlpi_name, "libcublas.so")) {
    *static_cast<std::string *>(data) = info->dlpi_name;
    return 1;
  }
  return 0;
}

void *cublas_library_handle() {
  static void *lib = [] {
    std::string loaded;
    dl_iterate_phdr(find_loaded_cublas, &loaded);
    void *h = nullptr;
    if (!loaded.empty()) h = dlopen(loaded.c_str(), RTLD_NOW | RTLD_LOCAL);
    for (const char *name : {"libcublas.so.12", "libcublas.so"}) {
      if (h) break;
      h = dlopen(name, RTLD_NOW | RTLD_LOCAL);
    }
    return h;
  }();
  return lib;
}
}  // namespace

void *real_cublas_symbol(const char *name) {
  static std::mutex mu;
  static std::map<std::string, void *> cache;
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(name);
  if (it != cache.end()) return it->second;
  void *fn = dlsym(RTLD_NEXT, name);
  if (fn == nullptr) {
    if (void *lib = cublas_library_handle()) fn = dlsym(lib, name);
  }
  // never hand back one of our own interposers
  Dl_info self{}, found{};
  if (fn && dladdr(reinterpret_cast<void *>(&real_cublas_symbol), &self) && dladdr(fn, &found) &&
      self.dli_fbase == found.dli_fbase) {
    fn = nullptr;
    if (void *lib = cublas_library_handle()) fn = dlsym(lib, name);
    if (fn && dladdr(fn, &found) && self.dli_fbase == found.dli_fbase) fn = nullptr;
  }
  if (fn == nullptr) log_error(std::string("Failed to resolve the cuBLAS function ") + name);
  cache[name] = fn;
  return fn;
}

// ---- profiler ---------------------------------------------------------------------------------
void StageProfiler::start(const std::string &name, cudaStream_t s) {
  if (!enabled) return;
  cudaStreamSynchronize(s);
  open[name] = std::chrono::steady_clock::now();
}

void StageProfiler::stop(const std::string &name, cudaStream_t s) {
  if (!enabled) return;
  cudaStreamSynchronize(s);
  const auto t1 = std::chrono::steady_clock::now();
  auto it = open.find(name);
  if (it == open.end()) return;
  Entry &e = entries[name];
  e.count++;
  e.seconds += std::chrono::duration<double>(t1 - it->second).count();
  open.erase(it);
}

void StageProfiler::print(const std::string &tag, bool csv) const {
  double total = 0;
  for (const auto &kv : entries) total += kv.second.seconds;
  if (csv) {
    std::printf("tag,stage,count,total_s,mean_s,share\n");
    for (const auto &kv : entries)
      std::printf("%s,%s,%llu,%.6e,%.6e,%.4f\n", tag.c_str(), kv.first.c_str(),
                  static_cast<unsigned long long>(kv.second.count), kv.second.seconds,
                  kv.second.seconds / std::max<std::uint64_t>(1, kv.second.count),
                  total > 0 ? kv.second.seconds / total : 0.0);
  } else {
    std::printf("# ozIMMU profiler [%s]\n", tag.c_str());
    for (const auto &kv : entries)
      std::printf("  %-20s n=%-6llu total=%10.3f ms  mean=%10.3f us  %5.1f%%\n", kv.first.c_str(),
                  static_cast<unsigned long long>(kv.second.count), kv.second.seconds * 1e3,
                  kv.second.seconds * 1e6 / std::max<std::uint64_t>(1, kv.second.count),
                  total > 0 ? 100.0 * kv.second.seconds / total : 0.0);
  }
  std::fflush(stdout);
}

// ---- workspace --------------------------------------------------------------------------------
namespace {
std::size_t align_up(std::size_t v, std::size_t a) { return (v + a - 1) / a * a; }
}  // namespace

WorkspaceLayout workspace_layout(std::size_t m, std::size_t n, std::size_t k, unsigned num_split, unsigned planes) {
  WorkspaceLayout w{};
  w.pitch = slice_pitch(k);
  w.a_plane = slices_bytes(m, k, num_split);
  w.b_plane = slices_bytes(n, k, num_split);
  std::size_t off = 0;
  w.off_amax = off;      off = align_up(off + sizeof(double) * m * planes, 256);
  w.off_bmax = off;      off = align_up(off + sizeof(double) * n * planes, 256);
  w.off_scr_a = off;     off = align_up(off + sizeof(std::uint32_t) * m * planes, 256);
  w.off_scr_b = off;     off = align_up(off + sizeof(std::uint32_t) * n * planes, 256);
  w.off_a_slices = off;  off = align_up(off + w.a_plane * planes, 1024);
  w.off_b_slices = off;  off = align_up(off + w.b_plane * planes, 1024);
  w.total = off;
  return w;
}

std::vector<std::pair<int, int>> pair_list(unsigned num_split) {
  std::vector<std::pair<int, int>> out;
  const int s = static_cast<int>(num_split);
  for (int sum = 2; sum <= s + 1; sum++)
    for (int a = 1; a < sum; a++) out.emplace_back(a, sum - a);
  return out;
}

}  // namespace host
}  // namespace oz

extern "C" unsigned long long ozimmu_launch_count(void) { return oz::g_launch_count; }
