// abi.cu -- library-level entry points and the error / launch-count plumbing.
#include <atomic>
#include <mutex>

#include "common.cuh"

namespace sed {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

int sm_count() {
  static std::mutex mu;
  static int cache[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  std::lock_guard<std::mutex> lk(mu);
  if (cache[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cache[dev] = n;
  }
  return cache[dev];
}

}  // namespace sed

extern "C" {

const char* sed_last_error_string(void) { return sed::g_err; }

int sed_abi_version(void) { return 1; }

unsigned long long sed_launch_count(void) {
  return sed::g_launches.load(std::memory_order_relaxed);
}

int sed_device_sm_count(int* out_sms) {
  SED_REQUIRE(out_sms != nullptr, "sed_device_sm_count: null output");
  int dev = 0;
  SED_CUDA(cudaGetDevice(&dev));
  SED_CUDA(cudaDeviceGetAttribute(out_sms, cudaDevAttrMultiProcessorCount, dev));
  return 0;
}

}  // extern "C"
