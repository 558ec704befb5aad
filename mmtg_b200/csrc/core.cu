// Library-wide plumbing: thread-local error text, SM count cache, launch counter.
#include <atomic>
#include <mutex>
#include <vector>
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

int current_device_index() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  return dev;
}

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

// ---- optional per-launch timing (bench.py roofline): CUDA events on the launching stream ----
struct ProfSlot {
  cudaEvent_t e0, e1;
  int cls;
  double flops, bytes;
};
static bool g_prof_on = false;
static std::vector<ProfSlot> g_slots;
static std::mutex g_prof_mu;

int prof_begin(int cls, double flops, double bytes, cudaStream_t st) {
  if (!g_prof_on) return -1;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfSlot s;
  s.cls = cls; s.flops = flops; s.bytes = bytes;
  if (cudaEventCreate(&s.e0) != cudaSuccess || cudaEventCreate(&s.e1) != cudaSuccess) return -1;
  cudaEventRecord(s.e0, st);
  g_slots.push_back(s);
  return (int)g_slots.size() - 1;
}
void prof_end(int idx, cudaStream_t st) {
  if (idx < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (idx < (int)g_slots.size()) cudaEventRecord(g_slots[idx].e1, st);
}

}  // namespace mmtg

extern "C" const char* mmtg_last_error(void) { return mmtg::get_last_error(); }
extern "C" int mmtg_abi_version(void) { return MMTG_ABI_VERSION; }
extern "C" int64_t mmtg_launch_count(void) { return (int64_t)mmtg::g_launches.load(); }

// Writes one CSV line per recorded launch: class, ms, flops, bytes (diagnostics for profiles/).
extern "C" int mmtg_prof_dump(const char* path) {
  std::lock_guard<std::mutex> lk(mmtg::g_prof_mu);
  FILE* f = fopen(path, "w");
  if (!f) return -1;
  fprintf(f, "class,ms,flops,bytes\n");
  for (auto& s : mmtg::g_slots) {
    float e = 0.f;
    if (cudaEventSynchronize(s.e1) != cudaSuccess) continue;
    if (cudaEventElapsedTime(&e, s.e0, s.e1) != cudaSuccess) continue;
    fprintf(f, "%d,%.6f,%.0f,%.0f\n", s.cls, e, s.flops, s.bytes);
  }
  fclose(f);
  return 0;
}
extern "C" void mmtg_prof_enable(int32_t on) { mmtg::g_prof_on = on != 0; }
extern "C" void mmtg_prof_reset(void) {
  std::lock_guard<std::mutex> lk(mmtg::g_prof_mu);
  for (auto& s : mmtg::g_slots) {
    cudaEventDestroy(s.e0);
    cudaEventDestroy(s.e1);
  }
  mmtg::g_slots.clear();
}
// Sums (after synchronising the events) the device time, algorithmic FLOPs and bytes of every
// recorded launch of class `cls` (0 = tcgen05 GEMM, 1 = attention, 2 = row/reduction kernels).
extern "C" int mmtg_prof_collect(int32_t cls, double* ms, double* flops, double* bytes, int64_t* count) {
  std::lock_guard<std::mutex> lk(mmtg::g_prof_mu);
  double t = 0, f = 0, b = 0;
  int64_t n = 0;
  for (auto& s : mmtg::g_slots) {
    if (s.cls != cls) continue;
    if (cudaEventSynchronize(s.e1) != cudaSuccess) continue;
    float e = 0.f;
    if (cudaEventElapsedTime(&e, s.e0, s.e1) != cudaSuccess) continue;
    t += e; f += s.flops; b += s.bytes; ++n;
  }
  if (ms) *ms = t;
  if (flops) *flops = f;
  if (bytes) *bytes = b;
  if (count) *count = n;
  return 0;
}
