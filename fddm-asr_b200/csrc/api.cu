// api.cu -- library-level entry points: version, thread-local error string, launch accounting.
#include <stdarg.h>
#include <stdio.h>

#include <string.h>

#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "common.cuh"

namespace fddm {
namespace {
thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};   // process-wide: autograd runs the backward launches on its own thread
std::atomic<int> g_sm_reserve{0};     // SMs the persistent row kernels leave free (fddm_set_sm_reserve)
}  // namespace

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int num_sms() {
  // per-device cache (the library is used by one process per GPU, but stay correct regardless)
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// ---- NVTX ranges + optional per-kernel CUDA-event timing ----------------------------------------
namespace {
struct TimedLaunch { const char* name; cudaEvent_t e0, e1; };
std::atomic<int> g_profiling{0};
std::mutex g_prof_mu;
std::vector<TimedLaunch> g_prof;
}  // namespace

ApiRange::ApiRange(const char* name) { nvtxRangePushA(name); }
ApiRange::~ApiRange() { nvtxRangePop(); }

KernelScope::KernelScope(const char* name, cudaStream_t stream) : name_(name), stream_(stream), e0_(nullptr) {
  nvtxRangePushA(name);
  if (g_profiling.load(std::memory_order_relaxed) == 0) return;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) return;
  if (cudaEventCreate(&e0_) != cudaSuccess) { e0_ = nullptr; return; }
  cudaEventRecord(e0_, stream);
}
KernelScope::~KernelScope() {
  if (e0_ != nullptr) {
    cudaEvent_t e1 = nullptr;
    if (cudaEventCreate(&e1) == cudaSuccess) {
      cudaEventRecord(e1, stream_);
      std::lock_guard<std::mutex> lk(g_prof_mu);
      g_prof.push_back({name_, e0_, e1});
    } else {
      cudaEventDestroy(e0_);
    }
  }
  nvtxRangePop();
}

int row_kernel_sms() {
  const int n = num_sms() - g_sm_reserve.load(std::memory_order_relaxed);
  return n > 8 ? n : 8;
}

}  // namespace fddm

extern "C" {
int fddm_set_sm_reserve(int n) {
  if (n < 0 || n > 128) {
    fddm::set_error("set_sm_reserve: n must be in [0, 128]");
    return FDDM_EINVAL;
  }
  fddm::g_sm_reserve.store(n, std::memory_order_relaxed);
  return FDDM_OK;
}

int fddm_profile_enable(int on) {
  using namespace fddm;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (on) {
    for (auto& t : g_prof) { cudaEventDestroy(t.e0); cudaEventDestroy(t.e1); }
    g_prof.clear();
  }
  g_profiling.store(on ? 1 : 0, std::memory_order_relaxed);
  return FDDM_OK;
}

int64_t fddm_profile_read(char* buf, int64_t cap) {
  using namespace fddm;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  std::map<std::string, std::pair<int64_t, double>> agg;
  for (auto& t : g_prof) {
    float ms = 0.0f;
    if (cudaEventSynchronize(t.e1) != cudaSuccess || cudaEventElapsedTime(&ms, t.e0, t.e1) != cudaSuccess) continue;
    auto& a = agg[t.name];
    a.first += 1;
    a.second += static_cast<double>(ms);
  }
  std::string out;
  char line[256];
  for (auto& kv : agg) {
    snprintf(line, sizeof(line), "%s\t%lld\t%.6f\n", kv.first.c_str(), static_cast<long long>(kv.second.first),
             kv.second.second);
    out += line;
  }
  const int64_t need = static_cast<int64_t>(out.size()) + 1;
  if (buf != nullptr && cap >= need) memcpy(buf, out.c_str(), static_cast<size_t>(need));
  return need;
}

int fddm_version(void) { return FDDM_ABI_VERSION; }
const char* fddm_last_error(void) { return fddm::g_err; }
int64_t fddm_launch_count(void) { return fddm::g_launches.load(std::memory_order_relaxed); }
}
