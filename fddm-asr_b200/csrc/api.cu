// api.cu -- library-level entry points: version, thread-local error string, launch accounting.
#include <stdarg.h>
#include <stdio.h>

#include <atomic>

#include "common.cuh"

namespace fddm {
namespace {
thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};   // process-wide: autograd runs the backward launches on its own thread
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

}  // namespace fddm

extern "C" {
int fddm_version(void) { return FDDM_ABI_VERSION; }
const char* fddm_last_error(void) { return fddm::g_err; }
int64_t fddm_launch_count(void) { return fddm::g_launches.load(std::memory_order_relaxed); }
}
