// mmtg_b200 — shared device/host helpers for the sm_100a kernels.
// PTX wrappers for mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (UMMA / TMEM) and
// a few math helpers. No torch types; everything below is plain CUDA.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace mmtg {

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------------------------------------
// error plumbing (host)
// ---------------------------------------------------------------------------------------------
void set_last_error(const char* fmt, ...);
const char* get_last_error();

#define MMTG_CHECK_ARG(cond, ...)                  \
  do {                                             \
    if (!(cond)) {                                 \
      ::mmtg::set_last_error(__VA_ARGS__);         \
      return -1;                                   \
    }                                              \
  } while (0)

#define MMTG_CUDA_OK(expr)                                                              \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      ::mmtg::set_last_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr,               \
                             cudaGetErrorString(_e));                                   \
      return (int)_e;                                                                   \
    }                                                                                   \
  } while (0)

#define MMTG_LAUNCH_OK()                                                                \
  do {                                                                                  \
    cudaError_t _e = cudaGetLastError();                                                \
    if (_e != cudaSuccess) {                                                            \
      ::mmtg::set_last_error("%s:%d launch -> %s", __FILE__, __LINE__,                  \
                             cudaGetErrorString(_e));                                   \
      return (int)_e;                                                                   \
    }                                                                                   \
  } while (0)

#define MMTG_TRY(expr)          \
  do {                          \
    int _rc = (expr);           \
    if (_rc != 0) return _rc;   \
  } while (0)

int num_sms();  // cached SM count of the current device
int current_device_index();  // cudaGetDevice() clamped to [0, 63]
// One-time CUDA state that lives PER DEVICE (function attributes, __constant__ uploads): the flag
// is indexed by the current device, so a process that touches a second GPU initialises it too.
#define MMTG_PER_DEVICE_FLAG(name)          \
  static bool name##_by_dev[64] = {false};  \
  bool& name = name##_by_dev[::mmtg::current_device_index()]

// Optional launch timing (see core.cu). classes: 0 GEMM, 1 attention, 2 row/reduction kernels.
int prof_begin(int cls, double flops, double bytes, cudaStream_t st);
void prof_end(int idx, cudaStream_t st);
struct ProfScope {
  int idx;
  cudaStream_t st;
  ProfScope(int cls, double flops, double bytes, cudaStream_t s) : idx(prof_begin(cls, flops, bytes, s)), st(s) {}
  ~ProfScope() { prof_end(idx, st); }
};

__host__ __device__ static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline long long cdivll(long long a, long long b) { return (a + b - 1) / b; }

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a protocol bug traps instead of hanging the GPU box.
#ifndef MMTG_SPIN_LIMIT
#define MMTG_SPIN_LIMIT (1u << 26)
#endif
// TAG makes each wait site recognisable in SASS / ncu source views (distinct spin-limit immediate).
template <uint32_t TAG = 0>
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (TAG >= 15 ? (MMTG_SPIN_LIMIT >> 6) : MMTG_SPIN_LIMIT) + TAG) {
      printf("mmtg: mbarrier wait %u timed out (block %d thread %d)\n", TAG, blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

// ---- TMA ----
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// multicast variant: the box lands at the same smem offset in every CTA of `mask`, and each
// destination CTA's mbarrier (same offset) receives the complete_tx for its copy.
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* map, uint64_t* bar,
                                               int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---- tcgen05 / TMEM ----
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate. One thread issues.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
          smem_u32(bar))
      : "memory");
}
// Same, but the arrival is delivered to the barrier at this offset in every CTA of `mask`.
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
// ---- cta_group::2 (CTA pair) variants: one MMA spans the two SMs of a cluster ----
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// D[tmem of both CTAs] (+)= A (128 rows from each CTA's smem) * B (N/2 rows from each CTA's smem).
// Issued by ONE thread of the leader CTA; M = 256 in the instruction descriptor.
__device__ __forceinline__ void umma_bf16_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                              uint32_t idesc, uint32_t accumulate) {
  const uint32_t z = 0;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(z)
      : "memory");
}
__device__ __forceinline__ void umma_commit2_mc(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
// TMA load into THIS CTA's smem whose completion is signalled on the LEADER CTA's mbarrier
// (same offset; bit 24 of the shared-window address selects the CTA within the pair).
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* map, uint64_t* bar,
                                                int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
// arrive on the barrier at the same offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, %1;\n"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(rank)
      : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets row (lane base + i).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
        "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// store 32 fp32 columns per lane back to TMEM (accumulator rescale in attention)
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),
        "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]),
        "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory matrix descriptor, SWIZZLE_128B, sm_100 version field = 1.
// Bit layout follows the PTX ISA "matrix descriptor" table: start address [0,14) >>4,
// leading byte offset [16,30) >>4, stride byte offset [32,46) >>4, version [46,48),
// swizzle mode [61,64) (2 = 128B).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr, uint32_t lbo_bytes,
                                                    uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
// Instruction descriptor for kind::f16 with bf16 A/B and fp32 D.
// c_format [4,6)=1 (F32), a_format [7,10)=1 (BF16), b_format [10,13)=1, a_major bit 15,
// b_major bit 16 (0 = K-major, 1 = MN-major), N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ __forceinline__ uint32_t umma_idesc_bf16(int M, int N, int a_mn_major,
                                                             int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= 1u << 7;
  d |= 1u << 10;
  d |= (uint32_t)(a_mn_major & 1) << 15;
  d |= (uint32_t)(b_mn_major & 1) << 16;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}

// ---- math ----
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// MUFU tanh (max rel. error 2^-11): used where the result is rounded to bf16 anyway.
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// gelu_new(x) = 0.5 x (1 + tanh(k0 (x + k1 x^3))), written as FMA chains (the epilogues that
// use these were measured ALU-bound: 20 FP instructions + 2 MUFU per element before).
__device__ __forceinline__ float gelu_new_fast(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  const float x2 = x * x;
  const float t = tanh_fast(x * fmaf(x2, k0 * k1, k0));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}
__device__ __forceinline__ float dgelu_new_fast(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  const float x2 = x * x;
  const float t = tanh_fast(x * fmaf(x2, k0 * k1, k0));
  const float r = (0.5f * x) * fmaf(x2, 3.f * k0 * k1, k0);
  return fmaf(fmaf(-t, t, 1.f), r, fmaf(0.5f, t, 0.5f));
}
// both from one tanh: 3 FMUL + 6 FFMA + 1 MUFU
__device__ __forceinline__ void gelu_new_both(float x, float& g, float& d) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  const float x2 = x * x;
  const float t = tanh_fast(x * fmaf(x2, k0 * k1, k0));
  const float hx = 0.5f * x;
  const float r = hx * fmaf(x2, 3.f * k0 * k1, k0);
  d = fmaf(fmaf(-t, t, 1.f), r, fmaf(0.5f, t, 0.5f));
  g = fmaf(hx, t, hx);
}
// gelu_new (tanh approximation), HF activations.py NewGELUActivation.
__device__ __forceinline__ float gelu_new_f(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float t = tanhf(k0 * (x + k1 * x * x * x));
  return 0.5f * x * (1.f + t);
}
__device__ __forceinline__ float dgelu_new_f(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float x2 = x * x;
  float t = tanhf(k0 * (x + k1 * x * x2));
  float dt = (1.f - t * t) * k0 * (1.f + 3.f * k1 * x2);
  return 0.5f * (1.f + t) + 0.5f * x * dt;
}

// ---- dropout: counter-based mask, a pure function of (step seed, site, element index) ----
// The forward and backward kernels of a site regenerate the same mask instead of storing it.
// One 32-bit hash decides TWO neighbouring elements (16 random bits each); keyed at two points
// of the mixing function so the streams of different sites are not shifted copies of each other.
struct DropKey {
  uint32_t k1, k2, thresh;  // thresh = round(p * 65536); keep iff bits >= thresh
  float inv_keep;           // 1 / (1 - p)
};
__device__ __forceinline__ DropKey drop_key(const unsigned long long* seed, uint32_t site, float p) {
  DropKey k;
  const unsigned long long s = seed ? *seed : 0ull;
  uint32_t x = (uint32_t)s ^ ((uint32_t)(s >> 32) * 0x9E3779B9u) ^ (site * 0x85EBCA6Bu + 0xC2B2AE35u);
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  k.k1 = x;
  k.k2 = x * 0x9E3779B9u + 0x7F4A7C15u;
  k.thresh = (uint32_t)(p * 65536.f + 0.5f);
  k.inv_keep = 1.f / (1.f - p);
  return k;
}
// random bits of elements 2*pair (low half) and 2*pair + 1 (high half)
__device__ __forceinline__ uint32_t drop_bits(const DropKey& k, uint32_t pair) {
  uint32_t x = pair ^ k.k1;
  x ^= x >> 16; x *= 0x7feb352du;
  x ^= (x >> 15) ^ k.k2; x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}
__device__ __forceinline__ bool drop_keep_lo(const DropKey& k, uint32_t bits) { return (bits & 0xffffu) >= k.thresh; }
__device__ __forceinline__ bool drop_keep_hi(const DropKey& k, uint32_t bits) { return (bits >> 16) >= k.thresh; }
#endif  // __CUDACC__

}  // namespace mmtg
