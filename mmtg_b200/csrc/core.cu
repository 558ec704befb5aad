// Library-wide plumbing: thread-local error text, SM count cache, launch counter.
#include <atomic>
#include <stdarg.h>
#include <string.h>

#include "../../include/mmtg_b200.h"
#include "common.cuh"

namespace mmtg {

static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_last_error() { return g_err; }

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace mmtg

extern "C" const char* mmtg_last_error(void) { return mmtg::get_last_error(); }
extern "C" int mmtg_abi_version(void) { return MMTG_ABI_VERSION; }
extern "C" int64_t mmtg_launch_count(void) { return (int64_t)mmtg::g_launches.load(); }
