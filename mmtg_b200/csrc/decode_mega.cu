// Persistent decode megakernel: all GPT-2 blocks + ln_f + lm_head of ONE decode position in a
// single launch. Generation at batch <= 64 is latency-bound, not bandwidth-bound: the per-op path
// (mmtg_decode_step) pays ~90 launches per position. Here one CTA per SM stays resident and a
// grid-wide barrier separates five phases per block:
//   A: qkv_acc += bf16(h) · W'_attn            side jobs: LN1 row stats of h, u_acc = 0, h2 = h + b_proj
//   B: cached attention, one warp per (row, head): q/k/v = rstd (acc - mean cs) + b', append K/V,
//      softmax(q K^T) V on tensor cores over TMA-staged 64-key boxes
//   C: h2 += att · W_proj                      side job: qkv_acc = 0
//   D: u_acc += bf16(h2) · W'_fc               side jobs: LN2 row stats of h2, h_next = h2 + b_proj2
//   E: h_next += gelu(rstd (u_acc - mean cs) + b') · W_proj2
// and finally F: logits = rstd (bf16(h) · wte'^T - mean cs) + b'.
// LayerNorm is FOLDED into the following linear layer (mmtg_decode_fold_weights): W' = g ⊙ W,
// cs = column sums of W', b' = bias + beta · W, so LN(x) W + bias = rstd (x W' - mean cs) + b'.
// That takes the row statistics off the critical path: they are computed as a side job of the
// phase that streams the weights and applied by the consumer of the accumulator.
// Every GEMM phase is cut into 144 (64-column, k-slice) units so each CTA streams one slice of
// the weights; the first unit's weights are prefetched (cp.async) BEFORE the preceding grid
// barrier, as are the cached K/V boxes of the attention phase. Split-K partials meet in fp32
// global accumulators through red.global.add (summation order is not fixed: logits can differ
// in the last bits between runs). mma.sync m16n8k16 on 64x64 128B-swizzled smem tiles.
#include <string.h>

#include "../../include/mmtg_b200.h"
#include "ops.h"
#include "mma_tiles.cuh"

namespace mmtg {

void count_launch(int n = 1);
int make_tmap_bf16_2d(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld,
                      uint32_t box0, uint32_t box1);

namespace {

constexpr int MG_THREADS = 256;
constexpr int OFF_SW = 12 * 8192;              // sA: 12 activation tiles (lm_head: all of K = 768)
constexpr int OFF_STATS = OFF_SW + 16 * 8192;  // sW: 16 weight tiles
constexpr int OFF_BARS = OFF_STATS + 512;
constexpr int MG_SMEM = OFF_BARS + 128;
constexpr int ATT_WARPS = 6;  // attention: 2 x (8 KB K + 8 KB V) per warp = 192 KB

struct MegaParams {
  const float* P;
  const bf16* W;
  mmtg_param_offsets off;  // by value (~5 KB of kernel parameters)
  CUtensorMap tm_k, tm_v;  // [NL*B*NH*Lmax, 64] bf16, 64x64 boxes, 128B swizzle
  int B, E, NH, NL, V, Pl, Lmax;
  MegaBufs w;
  const int* j_ptr;
  float* logits;
  unsigned long long* trace;  // optional: CTA 0 stamps globaltimer after every phase
};

// Grid barrier split into arrive / wait: work that does not depend on the other CTAs' results of
// the finished phase (prefetch issue, zeroing / pre-initialising accumulators of LATER phases)
// runs between the two, inside the barrier's latency. Whatever that window writes is published
// by the CTA's NEXT arrive, so it may only be consumed two phases later.
__device__ __forceinline__ void barrier_arrive(unsigned int* counter) {
  __syncthreads();
  if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
}
__device__ __forceinline__ void barrier_wait(unsigned int* counter, unsigned int target) {
  if (threadIdx.x == 0) {
    unsigned int spins = 0;
    while (true) {
      unsigned int v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      if (v >= target) break;
      if (++spins > (1u << 24)) {
        printf("mmtg: decode grid barrier timed out (block %d, target %u, saw %u)\n", blockIdx.x, target, v);
        __trap();
      }
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void substamp(unsigned long long* sub, int k) {
  if (sub && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    sub[k] = t;
  }
}

// side job: (mean, rstd) of row `row` of a [*, 768] fp32 matrix by one warp
__device__ __forceinline__ void warp_row_stats(const float* __restrict__ src, int row, int E, float* __restrict__ out) {
  const int l = lane_id();
  const float4* xr = reinterpret_cast<const float4*>(src + (long long)row * E);
  float4 v[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) v[i] = __ldcg(xr + l + i * 32);  // written by other CTAs in this launch: L2, never .nc / L1
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 6; ++i) s += v[i].x + v[i].y + v[i].z + v[i].w;
  const float mean = warp_sum(s) / (float)E;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += a * a + b * b + c * c + d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)E + 1e-5f);
  if (l == 0) {
    out[row * 2] = mean;
    out[row * 2 + 1] = rstd;
  }
}

enum { A_BF16 = 0, A_RAW = 1, A_GELU = 2 };

// ---- weights: [K, N] (Conv1D) slice of KT 64x64 tiles, cp.async, one commit group ----
template <int KT>
__device__ __forceinline__ void issue_w_kn(const bf16* __restrict__ Wm, long long ldw, int n0, int k0, bf16* slot) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int t = 0; t < KT; ++t) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + i * MG_THREADS;
      const int r = idx >> 3, c = idx & 7;
      cp_async16(reinterpret_cast<uint8_t*>(slot + t * 4096) + r * 128 + ((c ^ (r & 7)) << 4),
                 Wm + (long long)(k0 + t * 64 + r) * ldw + n0 + c * 8, true);
    }
  }
  cp_async_commit();
}

// ---- activations: transform while staging (thread -> row tid/4, 16 columns per k tile) ----
//   A_BF16: A is a bf16 matrix; A_RAW: fp32 -> bf16;
//   A_GELU: gelu_new(rstd (src - mean cs[k]) + b[k]) with (mean, rstd) of the row from `stats`.
// Rows >= B are zero.
template <int MODE, int KT>
__device__ __forceinline__ void stage_a(const void* __restrict__ A, long long lda, const float* __restrict__ cs,
                                        const float* __restrict__ bvec, const float* __restrict__ stats, int B,
                                        int k0, bf16* sA) {
  const int tid = threadIdx.x;
  const int r = tid >> 2, cq = tid & 3, c0 = cq * 2;
  uint8_t* ta = reinterpret_cast<uint8_t*>(sA) + r * 128;
  const uint32_t o0 = ((c0) ^ (r & 7)) << 4, o1 = ((c0 + 1) ^ (r & 7)) << 4;
  if (r >= B) {
#pragma unroll
    for (int t = 0; t < KT; ++t) {
      *reinterpret_cast<uint4*>(ta + t * 8192 + o0) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(ta + t * 8192 + o1) = make_uint4(0, 0, 0, 0);
    }
    return;
  }
  if (MODE == A_BF16) {
    uint4 u[KT][2];
#pragma unroll
    for (int t = 0; t < KT; ++t) {
      const uint4* s4 = reinterpret_cast<const uint4*>((const bf16*)A + (long long)r * lda + k0 + t * 64 + cq * 16);
      u[t][0] = __ldcg(s4);
      u[t][1] = __ldcg(s4 + 1);
    }
#pragma unroll
    for (int t = 0; t < KT; ++t) {
      *reinterpret_cast<uint4*>(ta + t * 8192 + o0) = u[t][0];
      *reinterpret_cast<uint4*>(ta + t * 8192 + o1) = u[t][1];
    }
    return;
  }
  float4 v[KT][4];
#pragma unroll
  for (int t = 0; t < KT; ++t) {
    const float4* s4 = reinterpret_cast<const float4*>((const float*)A + (long long)r * lda + k0 + t * 64 + cq * 16);
#pragma unroll
    for (int i = 0; i < 4; ++i) v[t][i] = __ldcg(s4 + i);
  }
  float mean = 0.f, rstd = 0.f;
  if (MODE == A_GELU) {
    mean = __ldcg(stats + r * 2);
    rstd = __ldcg(stats + r * 2 + 1);
  }
#pragma unroll
  for (int t = 0; t < KT; ++t) {
    const int kc = k0 + t * 64 + cq * 16;
    uint32_t pk[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 x = v[t][i];
      if (MODE == A_GELU) {
        const float4 bb = __ldg(reinterpret_cast<const float4*>(bvec + kc) + i);
        const float4 cc = __ldg(reinterpret_cast<const float4*>(cs + kc) + i);
        x.x = gelu_new_fast(rstd * (x.x - mean * cc.x) + bb.x);
        x.y = gelu_new_fast(rstd * (x.y - mean * cc.y) + bb.y);
        x.z = gelu_new_fast(rstd * (x.z - mean * cc.z) + bb.z);
        x.w = gelu_new_fast(rstd * (x.w - mean * cc.w) + bb.w);
      }
      pk[2 * i] = pack_bf16(x.x, x.y);
      pk[2 * i + 1] = pack_bf16(x.z, x.w);
    }
    *reinterpret_cast<uint4*>(ta + t * 8192 + o0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    *reinterpret_cast<uint4*>(ta + t * 8192 + o1) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
  }
}

// 32-column half of the 64x64 tile product, W tile [64 k][64 n] (n contiguous)
__device__ __forceinline__ void mma_nn_half(float (&c)[4][4], const uint32_t (&a)[4][4], const bf16* tile, int nb0) {
  const int l = lane_id();
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    const int r = kk * 16 + (l & 7) + ((l >> 3) & 1) * 8;
#pragma unroll
    for (int nb = 0; nb < 4; nb += 2) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(tile_addr(tile, r, nb0 + nb + (l >> 4)), b0, b1, b2, b3);
      mma16816(c[nb], a[kk][0], a[kk][1], a[kk][2], a[kk][3], b0, b1);
      mma16816(c[nb + 1], a[kk][0], a[kk][1], a[kk][2], a[kk][3], b2, b3);
    }
  }
}

// out[0:B, n0:n0+64] += sA(64 x 64*KT) * slot(64*KT x 64) through red.global.add (split-K partial)
template <int KT>
__device__ __forceinline__ void mma_red(const bf16* sA, const bf16* slot, int B, int n0, float* __restrict__ out,
                                        long long ldo) {
  const int warp = threadIdx.x >> 5, l = lane_id();
  const int m = warp & 3, nh = warp >> 2;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll
  for (int t = 0; t < KT; ++t) {
    uint32_t a[4][4];
    load_a_frags(sA + t * 4096, m * 16, a);
    mma_nn_half(acc, a, slot + t * 4096, nh * 4);
  }
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    const int row = m * 16 + (l >> 2) + rr * 8;
    if (row >= B) continue;
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) {
      float* dst = out + (long long)row * ldo + n0 + (nh * 4 + nb) * 8 + (l & 3) * 2;
      asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(dst), "f"(acc[nb][2 * rr]), "f"(acc[nb][2 * rr + 1])
                   : "memory");
    }
  }
}

// One GEMM phase of a block: units (n chunk, k slice) round-robin over CTAs. The first unit's
// weights were prefetched into `slot` two barrier windows earlier; exactly one younger cp.async
// group (the prefetch for the next GEMM phase) is in flight when the phase starts.
template <int MODE, int KT>
__device__ __forceinline__ void gemm_phase(const void* A, long long lda, const float* cs, const float* bvec,
                                           const float* stats, const bf16* Wm, long long ldw, int B, int nunits,
                                           int ks, float* out, long long ldo, bf16* sA, bf16* slot,
                                           unsigned long long* sub = nullptr) {
  const int G = gridDim.x;
  for (int u = blockIdx.x; u < nunits; u += G) {
    const int n0 = (u / ks) * 64, k0 = (u % ks) * KT * 64;
    const bool own = u != (int)blockIdx.x;  // not the prefetched first unit
    if (own) issue_w_kn<KT>(Wm, ldw, n0, k0, slot);
    stage_a<MODE, KT>(A, lda, cs, bvec, stats, B, k0, sA);
    substamp(sub, 2);
    // the prefetched unit's group is the second-youngest: one newer prefetch may stay in flight
    if (own) cp_async_wait<0>();
    else cp_async_wait<1>();
    __syncthreads();
    substamp(sub, 3);
    mma_red<KT>(sA, slot, B, n0, out, ldo);
    __syncthreads();
    substamp(sub, 4);
  }
}
template <int KT>
__device__ __forceinline__ void prefetch_w(const bf16* Wm, long long ldw, int nunits, int ks, bf16* slot) {
  const int u = blockIdx.x;
  if (u < nunits) issue_w_kn<KT>(Wm, ldw, (u / ks) * 64, (u % ks) * KT * 64, slot);
}

// ---------------------------------------------------------------------------------------------
// Cached attention: one warp per (row, head). 64-key boxes of the K and V cache arrive through
// TMA (128B swizzle, two buffers per warp, the first two issued one phase ahead) so ~190 KB per
// SM is in flight: at late positions the KV read is the largest HBM stream of the step. The
// single query row is row 0 of an m16n8k16 A operand (lanes 0-3 hold it, other rows are zero):
// 32 + 32 tensor-core instructions per box instead of ~1000 scalar ones. Online softmax across
// boxes; keys past `pos` and padded keys are masked by selection, never by arithmetic.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void attn_issue(const MegaParams& p, int layer, int u, int c, uint8_t* wbuf, uint64_t* bars) {
  const int row = ((layer * p.B * p.NH + u) * p.Lmax) + c * 64;
  uint8_t* buf = wbuf + (c & 1) * 16384;
  mbar_arrive_expect_tx(&bars[c & 1], 16384u);
  tma_load_2d(buf, &p.tm_k, &bars[c & 1], 0, row);
  tma_load_2d(buf + 8192, &p.tm_v, &bars[c & 1], 0, row);
}

// lane 0: pull the K/V boxes 2.. of a unit (the ones that do not fit the two staged buffers)
// into L2 right behind the TMA loads of boxes 0 and 1
__device__ __forceinline__ void attn_prefetch_l2(const MegaParams& p, int layer, int u, int pos) {
  const int rows = min((pos / 64 + 1) * 64, p.Lmax) - 128;
  if (rows <= 0) return;
  const size_t base = ((size_t)layer * p.B * p.NH + u) * p.Lmax * 64 + 128 * 64;
  const uint32_t bytes = (uint32_t)rows * 128u;
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.w.kcache + base), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.w.vcache + base), "r"(bytes) : "memory");
}

__device__ __forceinline__ void attn_warp(const MegaParams& p, const float* __restrict__ fv, int layer, int u, int pos,
                                          uint8_t* wbuf, uint64_t* bars, uint32_t& par, bool prefetched) {
  const int l = lane_id(), g4 = l & 3;
  const bool act = l < 4;  // lanes holding row 0 of the MMA fragments
  const int E = p.E, NH = p.NH;
  const int b = u / NH, h = u - b * NH;
  const int nch = pos / 64 + 1;
  if (!prefetched && l == 0) {
    fence_proxy_async();
    attn_issue(p, layer, u, 0, wbuf, bars);
    if (nch > 1) attn_issue(p, layer, u, 1, wbuf, bars);
  }
  const float mean = __ldcg(p.w.row_stats + b * 2), rstd = __ldcg(p.w.row_stats + b * 2 + 1);
  const float* acc_row = p.w.qkv_acc + (long long)b * 3 * E + h * 64;
  const float* cs = fv + h * 64;          // cs_attn
  const float* bq = fv + 3 * E + h * 64;  // b'_attn
  // q as A-operand fragments: lane g4 holds columns kk*16 + 2 g4 + {0,1} (a0) and +8 (a2)
  uint32_t qa[4][2];
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int col = kk * 16 + hh * 8 + 2 * g4;
      const float2 a = __ldcg(reinterpret_cast<const float2*>(acc_row + col));
      const float2 c2 = __ldg(reinterpret_cast<const float2*>(cs + col));
      const float2 b2 = __ldg(reinterpret_cast<const float2*>(bq + col));
      const float q0 = (rstd * (a.x - mean * c2.x) + b2.x) * 0.125f;
      const float q1 = (rstd * (a.y - mean * c2.y) + b2.y) * 0.125f;
      qa[kk][hh] = act ? pack_bf16(q0, q1) : 0u;
    }
  }
  // this position's K (lanes 0-7) and V (lanes 8-15) rows, 8 dims per lane: to the cache and,
  // below, into the staged box
  const int ks = l >> 3, part = l & 7;
  uint4 newrow = make_uint4(0, 0, 0, 0);
  const size_t cbase = ((size_t)layer * p.B * NH + u) * p.Lmax * 64;
  if (ks < 2) {
    const int col = (ks + 1) * E + part * 8;
    float x[8];
    const float4 a0 = __ldcg(reinterpret_cast<const float4*>(acc_row + col)), a1 = __ldcg(reinterpret_cast<const float4*>(acc_row + col + 4));
    const float4 c0 = __ldg(reinterpret_cast<const float4*>(cs + col)), c1 = __ldg(reinterpret_cast<const float4*>(cs + col + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bq + col)), b1 = __ldg(reinterpret_cast<const float4*>(bq + col + 4));
    x[0] = rstd * (a0.x - mean * c0.x) + b0.x; x[1] = rstd * (a0.y - mean * c0.y) + b0.y;
    x[2] = rstd * (a0.z - mean * c0.z) + b0.z; x[3] = rstd * (a0.w - mean * c0.w) + b0.w;
    x[4] = rstd * (a1.x - mean * c1.x) + b1.x; x[5] = rstd * (a1.y - mean * c1.y) + b1.y;
    x[6] = rstd * (a1.z - mean * c1.z) + b1.z; x[7] = rstd * (a1.w - mean * c1.w) + b1.w;
    newrow = make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
    bf16* dst = (ks == 0 ? p.w.kcache : p.w.vcache) + cbase + (size_t)pos * 64 + part * 8;
    *reinterpret_cast<uint4*>(dst) = newrow;
  }
  float m = -INFINITY, lsum = 0.f;
  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  // key-padding mask bits of keys 0..pos, one load batch up front: lane i keeps word i
  const int* km = p.w.keymask + (long long)b * p.Lmax;
  uint32_t mword = 0;
  {
    const int nw = (pos + 32) / 32;  // <= 32 words (Lmax <= 1024)
    int kv[8];
#pragma unroll 1
    for (int w0 = 0; w0 < nw; w0 += 8) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int key = (w0 + i) * 32 + l;
        kv[i] = (w0 + i < nw && key <= pos) ? km[key] : 0;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t bits = __ballot_sync(0xffffffffu, kv[i] != 0);
        if (l == w0 + i) mword = bits;
      }
    }
  }
#pragma unroll 1
  for (int c = 0; c < nch; ++c) {
    uint8_t* kb = wbuf + (c & 1) * 16384;
    uint8_t* vb = kb + 8192;
    const uint32_t mlo = __shfl_sync(0xffffffffu, mword, (2 * c) & 31);
    const uint32_t mhi = __shfl_sync(0xffffffffu, mword, (2 * c + 1) & 31);
    mbar_wait<20>(&bars[c & 1], (par >> (c & 1)) & 1u);
    par ^= 1u << (c & 1);
    if (c == nch - 1) {
      const int rl = pos - c * 64;
      if (ks < 2) *reinterpret_cast<uint4*>((ks == 0 ? kb : vb) + rl * 128 + ((part ^ (rl & 7)) << 4)) = newrow;
      __syncwarp();
    }
    // S = q K^T (row 0 of the 16 x 64 product)
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
    {
      const bf16* tile = reinterpret_cast<const bf16*>(kb);
#pragma unroll
      for (int nb = 0; nb < 8; nb += 2) {
        const int r = nb * 8 + (l & 7) + ((l >> 4) & 1) * 8;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4(tile_addr(tile, r, kk * 2 + ((l >> 3) & 1)), b0, b1, b2, b3);
          mma16816(s[nb], qa[kk][0], 0u, qa[kk][1], 0u, b0, b1);
          mma16816(s[nb + 1], qa[kk][0], 0u, qa[kk][1], 0u, b2, b3);
        }
      }
    }
    // mask + online softmax on lanes 0-3 (keys nb*8 + 2 g4 + {0,1}); other lanes hold zero rows
    float cmax = -INFINITY;
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int kl = nb * 8 + 2 * g4 + e;
        const bool ok = kl < 32 ? ((mlo >> kl) & 1u) : ((mhi >> (kl - 32)) & 1u);
        s[nb][e] = ok ? s[nb][e] : -INFINITY;
        cmax = fmaxf(cmax, s[nb][e]);
      }
    }
    cmax = fmaxf(cmax, __shfl_xor_sync(0xffffffffu, cmax, 1));
    cmax = fmaxf(cmax, __shfl_xor_sync(0xffffffffu, cmax, 2));
    const float m_new = __shfl_sync(0xffffffffu, fmaxf(m, cmax), 0);
    if (m_new != -INFINITY) {  // warp-uniform
      const float scale = (m == -INFINITY) ? 0.f : __expf(m - m_new);
      lsum *= scale;
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        o[nb][0] *= scale;
        o[nb][1] *= scale;
      }
      uint32_t pa[4][2];
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        const float p0 = (s[nb][0] == -INFINITY) ? 0.f : __expf(s[nb][0] - m_new);
        const float p1 = (s[nb][1] == -INFINITY) ? 0.f : __expf(s[nb][1] - m_new);
        lsum += p0 + p1;
        pa[nb >> 1][nb & 1] = act ? pack_bf16(p0, p1) : 0u;
      }
      const bf16* tile = reinterpret_cast<const bf16*>(vb);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int r = kk * 16 + (l & 7) + ((l >> 3) & 1) * 8;
#pragma unroll
        for (int nb = 0; nb < 8; nb += 2) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4_t(tile_addr(tile, r, nb + (l >> 4)), b0, b1, b2, b3);
          mma16816(o[nb], pa[kk][0], 0u, pa[kk][1], 0u, b0, b1);
          mma16816(o[nb + 1], pa[kk][0], 0u, pa[kk][1], 0u, b2, b3);
        }
      }
      m = m_new;
    }
    __syncwarp();
    if (c + 2 < nch && l == 0) {
      fence_proxy_async();
      attn_issue(p, layer, u, c + 2, wbuf, bars);
    }
  }
  lsum += __shfl_xor_sync(0xffffffffu, lsum, 1);
  lsum += __shfl_xor_sync(0xffffffffu, lsum, 2);
  const float inv = lsum > 0.f ? 1.f / lsum : 0.f;
  if (act) {
    bf16* dst = p.w.att16 + (long long)b * E + h * 64 + 2 * g4;
#pragma unroll
    for (int nb = 0; nb < 8; ++nb)
      *reinterpret_cast<uint32_t*>(dst + nb * 8) = pack_bf16(o[nb][0] * inv, o[nb][1] * inv);
  }
}

// ---- lm_head: 32-vocab-row units over the full K = 768 (wte' is [V, K]), direct stores ----
__device__ __forceinline__ void issue_w_head(const bf16* __restrict__ wte, int V, int E, int n0, uint8_t* buf) {
  const int tid = threadIdx.x;
  const int kt = E / 64;
  for (int idx = tid; idx < kt * 256; idx += MG_THREADS) {
    const int t = idx >> 8, r = (idx >> 3) & 31, c = idx & 7;
    const bool ok = (n0 + r) < V;
    cp_async16(buf + t * 4096 + r * 128 + ((c ^ (r & 7)) << 4), wte + (long long)(ok ? n0 + r : 0) * E + t * 64 + c * 8,
               ok);
  }
  cp_async_commit();
}
// stage bf16(h) for all of K and compute the ln_f row statistics in the same pass: the 4 threads
// of a row read exactly that row (shifted sums, quad reduction)
__device__ __forceinline__ void stage_head(const float* __restrict__ h, int B, int E, bf16* sA, float* stats) {
  const int tid = threadIdx.x;
  const int r = tid >> 2, cq = tid & 3, c0 = cq * 2;
  uint8_t* ta = reinterpret_cast<uint8_t*>(sA) + r * 128;
  const uint32_t o0 = ((c0) ^ (r & 7)) << 4, o1 = ((c0 + 1) ^ (r & 7)) << 4;
  const bool valid = r < B;
  const float* hr = h + (long long)(valid ? r : 0) * E;
  const float shift = __ldcg(hr);
  float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
  for (int t0 = 0; t0 < E / 64; t0 += 6) {
    float4 v[6][4];
#pragma unroll
    for (int t = 0; t < 6; ++t) {
      const float4* s4 = reinterpret_cast<const float4*>(hr + (t0 + t) * 64 + cq * 16);
#pragma unroll
      for (int i = 0; i < 4; ++i) v[t][i] = __ldcg(s4 + i);
    }
#pragma unroll
    for (int t = 0; t < 6; ++t) {
      uint32_t pk[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float4 x = v[t][i];
        if (!valid) x = make_float4(0.f, 0.f, 0.f, 0.f);
        const float a = x.x - shift, b = x.y - shift, c = x.z - shift, d = x.w - shift;
        s1 += a + b + c + d;
        s2 += a * a + b * b + c * c + d * d;
        pk[2 * i] = pack_bf16(x.x, x.y);
        pk[2 * i + 1] = pack_bf16(x.z, x.w);
      }
      *reinterpret_cast<uint4*>(ta + (t0 + t) * 8192 + o0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      *reinterpret_cast<uint4*>(ta + (t0 + t) * 8192 + o1) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    }
  }
  s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
  s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
  s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
  s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
  if (cq == 0) {
    const float dm = s1 / (float)E;
    const float var = fmaxf(s2 / (float)E - dm * dm, 0.f);
    stats[r * 2] = valid ? shift + dm : 0.f;
    stats[r * 2 + 1] = rsqrtf(var + 1e-5f);
  }
}
__device__ __forceinline__ void head_unit(const bf16* sA, const uint8_t* buf, const float* stats,
                                          const float* __restrict__ cs, const float* __restrict__ bh, int B, int V,
                                          int E, int n0, float* __restrict__ logits) {
  const int warp = threadIdx.x >> 5, l = lane_id();
  const int m = warp & 3, nq = warp >> 2;
  float acc[2][4], csv[2][2], bhv[2][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll
  for (int nb = 0; nb < 2; ++nb)  // epilogue vectors: loads in flight during the product
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int col = n0 + nq * 16 + nb * 8 + (l & 3) * 2 + e;
      csv[nb][e] = col < V ? __ldg(cs + col) : 0.f;
      bhv[nb][e] = col < V ? __ldg(bh + col) : 0.f;
    }
  const int kt = E / 64;
#pragma unroll 4
  for (int t = 0; t < kt; ++t) {
    uint32_t a[4][4];
    load_a_frags(sA + t * 4096, m * 16, a);
    const bf16* tile = reinterpret_cast<const bf16*>(buf + t * 4096);
    const int r = nq * 16 + (l & 7) + ((l >> 4) & 1) * 8;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4(tile_addr(tile, r, kk * 2 + ((l >> 3) & 1)), b0, b1, b2, b3);
      mma16816(acc[0], a[kk][0], a[kk][1], a[kk][2], a[kk][3], b0, b1);
      mma16816(acc[1], a[kk][0], a[kk][1], a[kk][2], a[kk][3], b2, b3);
    }
  }
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    const int row = m * 16 + (l >> 2) + rr * 8;
    if (row >= B) continue;
    const float mean = stats[row * 2], rstd = stats[row * 2 + 1];
#pragma unroll
    for (int nb = 0; nb < 2; ++nb) {
      const int col = n0 + nq * 16 + nb * 8 + (l & 3) * 2;
      float* dst = logits + (long long)row * V + col;
      if (col < V) dst[0] = rstd * (acc[nb][2 * rr] - mean * csv[nb][0]) + bhv[nb][0];
      if (col + 1 < V) dst[1] = rstd * (acc[nb][2 * rr + 1] - mean * csv[nb][1]) + bhv[nb][1];
    }
  }
}

__global__ void __launch_bounds__(MG_THREADS, 1)
decode_mega_kernel(const __grid_constant__ MegaParams p) {
  extern __shared__ __align__(1024) uint8_t mg_smem[];
  bf16* sA = reinterpret_cast<bf16*>(mg_smem);
  bf16* sW = reinterpret_cast<bf16*>(mg_smem + OFF_SW);
  float* stats = reinterpret_cast<float*>(mg_smem + OFF_STATS);  // [64][2] (lm_head)
  uint64_t* bars = reinterpret_cast<uint64_t*>(mg_smem + OFF_BARS);
  // weight slots (64x64 tiles of sW): A 0-3, D 4-7, E 8-11, C 12; lm_head: two 48 KB buffers at 0
  bf16* slotA = sW;
  bf16* slotD = sW + 4 * 4096;
  bf16* slotE = sW + 8 * 4096;
  bf16* slotC = sW + 12 * 4096;
  const int G = gridDim.x, cta = blockIdx.x, tid = threadIdx.x, warp = tid >> 5;
  const int B = p.B, E = p.E, NH = p.NH;
  const int pos = p.Pl + *p.j_ptr;
  const long long gthreads = (long long)G * MG_THREADS, gtid = (long long)cta * MG_THREADS + tid;
  unsigned int phase = 0;
  int tr = 0;
  auto stamp = [&]() {
    if (p.trace && cta == 0 && tid == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      p.trace[tr++] = t;
    }
  };
  stamp();
  if (tid == 0) {
    for (int i = 0; i < 2 * ATT_WARPS; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
    tma_prefetch_desc(&p.tm_k);
    tma_prefetch_desc(&p.tm_v);
  }
  float* h = p.w.h;
  float* hn = p.w.h_alt;
  float* statsA = p.w.row_stats;        // LN1 stats of h
  float* statsD = p.w.row_stats + 128;  // LN2 stats of h2
  uint32_t att_par = 0;
  uint8_t* wbuf = mg_smem + warp * 32768;
  uint64_t* wbars = bars + 2 * warp;
  const int nA = (3 * E / 64) * 4, nC = (E / 64) * (E / 64), nD = (4 * E / 64) * 3, nE = (E / 64) * 12;
  const size_t wA = (size_t)E * 3 * E, wD = (size_t)E * 4 * E;
  prefetch_w<3>(p.w.f_attn, 3 * E, nA, 4, slotA);
  prefetch_w<1>(p.W + p.off.layer[0].proj_w, E, nC, E / 64, slotC);
  __syncthreads();

  for (int layer = 0; layer < p.NL; ++layer) {
    const mmtg_layer_offsets lo = p.off.layer[layer];
    const bool last = layer + 1 == p.NL;
    const float* fv = p.w.f_vec + (size_t)layer * 14 * E;  // cs_attn[3E] b_attn[3E] cs_fc[4E] b_fc[4E]
    unsigned long long* sub = (p.trace && last) ? p.trace + 64 : nullptr;
    // ---- A: qkv_acc += bf16(h) W'_attn ; side job: LN1 stats of h ----
    if (warp == 7 && cta < B) warp_row_stats(h, cta, E, statsA);
    substamp(sub, 1);
    gemm_phase<A_RAW, 3>(h, E, nullptr, nullptr, nullptr, p.w.f_attn + layer * wA, 3 * E, B, nA, 4, p.w.qkv_acc,
                         3 * E, sA, slotA, sub);
    if (sub && tid == 32) {  // per-CTA time of reaching the A -> B arrive
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      p.trace[80 + 320 + cta] = t;
    }
    barrier_arrive(p.w.barrier);
    // window: cached K/V boxes of this CTA's first attention units
    if (warp < ATT_WARPS && lane_id() == 0) {
      const int u = cta + G * warp;
      if (u < B * NH) {
        fence_proxy_async();
        attn_issue(p, layer, u, 0, wbuf, wbars);
        if (pos >= 64) attn_issue(p, layer, u, 1, wbuf, wbars);
      }
    }
    substamp(sub, 5);
    if (sub && tid == 0) {  // per-CTA arrival time at the A -> B barrier
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      p.trace[80 + cta] = t;
    }
    barrier_wait(p.w.barrier, ++phase * G);
    stamp();
    // ---- B: cached attention, one warp per (row, head) ----
    if (warp < ATT_WARPS) {
      bool first = true;
      for (int u = cta + G * warp; u < B * NH; u += G * ATT_WARPS) {
        attn_warp(p, fv, layer, u, pos, wbuf, wbars, att_par, first);
        first = false;
        __syncwarp();
      }
    } else {
      // the two warps without attention units: u_acc = 0 (for D) and h2 = h + b_proj (for C)
      const long long sth = (long long)G * 64, sid = (long long)cta * 64 + (tid - ATT_WARPS * 32);
      for (long long i = sid * 4; i < (long long)B * 4 * E; i += sth * 4)
        *reinterpret_cast<float4*>(p.w.u_acc + i) = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int i = (int)sid * 4; i < B * E; i += (int)sth * 4) {
        float4 v = __ldcg(reinterpret_cast<const float4*>(h + i));
        const float4 bb = __ldg(reinterpret_cast<const float4*>(p.P + lo.proj_b + i % E));
        v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
        *reinterpret_cast<float4*>(p.w.h2 + i) = v;
      }
    }
    if (sub && tid == 0) {  // per-CTA arrival time at the B -> C barrier
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      p.trace[80 + 160 + cta] = t;
    }
    barrier_arrive(p.w.barrier);
    prefetch_w<4>(p.w.f_fc + layer * wD, 4 * E, nD, 3, slotD);  // window: W'_fc for D
    barrier_wait(p.w.barrier, ++phase * G);
    stamp();
    // ---- C: h2 += att W_proj ----
    gemm_phase<A_BF16, 1>(p.w.att16, E, nullptr, nullptr, nullptr, p.W + lo.proj_w, E, B, nC, E / 64, p.w.h2, E, sA,
                          slotC);
    barrier_arrive(p.w.barrier);
    // window: W_proj2 for E; qkv_acc = 0 for the next block / next position; at the last block
    // pull this CTA's lm_head rows into L2
    prefetch_w<4>(p.W + lo.proj2_w, E, nE, 12, slotE);
    for (long long i = gtid * 4; i < (long long)B * 3 * E; i += gthreads * 4)
      *reinterpret_cast<float4*>(p.w.qkv_acc + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    if (last && tid == 0) {
      for (int u = cta; u * 32 < p.V; u += G) {
        const uint32_t bytes = (uint32_t)min(32, p.V - u * 32) * (uint32_t)E * 2u;
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.w.f_wte + (size_t)u * 32 * E), "r"(bytes)
                     : "memory");
      }
    }
    barrier_wait(p.w.barrier, ++phase * G);
    stamp();
    // ---- D: u_acc += bf16(h2) W'_fc ; side jobs: LN2 stats of h2, h_next = h2 + b_proj2 ----
    if (warp == 7 && cta < B) warp_row_stats(p.w.h2, cta, E, statsD);
    for (int i = (int)gtid * 4; i < B * E; i += (int)gthreads * 4) {
      float4 v = __ldcg(reinterpret_cast<const float4*>(p.w.h2 + i));
      const float4 bb = __ldg(reinterpret_cast<const float4*>(p.P + lo.proj2_b + i % E));
      v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
      *reinterpret_cast<float4*>(hn + i) = v;
    }
    gemm_phase<A_RAW, 4>(p.w.h2, E, nullptr, nullptr, nullptr, p.w.f_fc + layer * wD, 4 * E, B, nD, 3, p.w.u_acc,
                         4 * E, sA, slotD);
    barrier_arrive(p.w.barrier);
    // window: W'_attn of the next block, or the first lm_head unit
    if (!last) prefetch_w<3>(p.w.f_attn + (layer + 1) * wA, 3 * E, nA, 4, slotA);
    else issue_w_head(p.w.f_wte, p.V, E, cta * 32, reinterpret_cast<uint8_t*>(sW));
    barrier_wait(p.w.barrier, ++phase * G);
    stamp();
    // ---- E: h_next += gelu(rstd (u_acc - mean cs) + b') W_proj2 ----
    gemm_phase<A_GELU, 4>(p.w.u_acc, 4 * E, fv + 6 * E, fv + 10 * E, statsD, p.W + lo.proj2_w, E, B, nE, 12, hn, E, sA,
                          slotE);
    barrier_arrive(p.w.barrier);
    // window: W_proj of the next block, or the second lm_head unit
    if (!last) prefetch_w<1>(p.W + p.off.layer[layer + 1].proj_w, E, nC, E / 64, slotC);
    else issue_w_head(p.w.f_wte, p.V, E, (cta + G) * 32, reinterpret_cast<uint8_t*>(sW) + (E / 64) * 4096);
    barrier_wait(p.w.barrier, ++phase * G);
    stamp();
    float* t = h;
    h = hn;
    hn = t;
  }
  // ---- F: logits = rstd (bf16(h) wte'^T - mean cs) + b' (tied lm_head with ln_f folded in) ----
  {
    const int nun = cdiv(p.V, 32);
    uint8_t* hb = reinterpret_cast<uint8_t*>(sW);
    const int kt = E / 64, bufsz = kt * 4096;
    const float* hv = p.w.f_vec + (size_t)p.NL * 14 * E;  // cs_head[V] b_head[V]
    stage_head(h, B, E, sA, stats);
    unsigned long long* fsub = p.trace ? p.trace + 72 : nullptr;
    substamp(fsub, 0);
    int k = 0;
    for (int u = cta; u < nun; u += G, ++k) {
      if (u + G < nun) cp_async_wait<1>();
      else cp_async_wait<0>();
      __syncthreads();
      head_unit(sA, hb + (k & 1) * bufsz, stats, hv, hv + p.V, B, p.V, E, u * 32, p.logits);
      __syncthreads();
      substamp(fsub, 1 + k);
      if (u + 2 * G < nun) issue_w_head(p.w.f_wte, p.V, E, (u + 2 * G) * 32, hb + (k & 1) * bufsz);
    }
  }
  stamp();
}

unsigned long long* g_mega_trace = nullptr;

// ---- LayerNorm folding (once per set of weights) ----
// W'[k,n] = bf16(g[k] W[k,n]); cs[n] = sum_k W'[k,n]; bout[n] = bias[n] + sum_k beta[k] W[k,n]
__global__ void __launch_bounds__(256)
fold_kn_kernel(const float* __restrict__ Wsrc, const float* __restrict__ g, const float* __restrict__ beta,
               const float* __restrict__ bias, int K, int N, bf16* __restrict__ Wf, float* __restrict__ cs,
               float* __restrict__ bout) {
  // block = 32 columns x 8 row groups; row groups interleave over k, reduced through smem
  __shared__ float red[2][8][32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + tx;
  float s = 0.f, bb = 0.f;
  if (n < N) {
#pragma unroll 4
    for (int k = ty; k < K; k += 8) {
      const float w = Wsrc[(long long)k * N + n];
      const bf16 wf = __float2bfloat16(g[k] * w);
      Wf[(long long)k * N + n] = wf;
      s += __bfloat162float(wf);
      bb += beta[k] * w;
    }
  }
  red[0][ty][tx] = s;
  red[1][ty][tx] = bb;
  __syncthreads();
  if (ty == 0 && n < N) {
#pragma unroll
    for (int j = 1; j < 8; ++j) {
      s += red[0][j][tx];
      bb += red[1][j][tx];
    }
    cs[n] = s;
    bout[n] = bias[n] + bb;
  }
}
// tied lm_head: wte'[n,k] = bf16(g[k] wte[n,k]) (warp per vocabulary row)
__global__ void __launch_bounds__(256)
fold_head_kernel(const float* __restrict__ wte, const float* __restrict__ g, const float* __restrict__ beta, int V,
                 int E, bf16* __restrict__ Wf, float* __restrict__ cs, float* __restrict__ bout) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= V) return;
  const int l = lane_id();
  float s = 0.f, bb = 0.f;
  for (int k = l; k < E; k += 32) {
    const float w = wte[(long long)n * E + k];
    const bf16 wf = __float2bfloat16(g[k] * w);
    Wf[(long long)n * E + k] = wf;
    s += __bfloat162float(wf);
    bb += beta[k] * w;
  }
  s = warp_sum(s);
  bb = warp_sum(bb);
  if (l == 0) {
    cs[n] = s;
    bout[n] = bb;
  }
}

}  // namespace

int decode_fold_weights(const mmtg_model* m, const MegaBufs& w, cudaStream_t st) {
  const mmtg_dims& d = m->dims;
  const int E = d.E;
  const float* P = m->params;
  for (int l = 0; l < d.NL; ++l) {
    const mmtg_layer_offsets& lo = m->off.layer[l];
    float* fv = w.f_vec + (size_t)l * 14 * E;
    fold_kn_kernel<<<cdiv(3 * E, 32), 256, 0, st>>>(P + lo.attn_w, P + lo.ln1_w, P + lo.ln1_b, P + lo.attn_b, E, 3 * E,
                                                     w.f_attn + (size_t)l * E * 3 * E, fv, fv + 3 * E);
    MMTG_LAUNCH_OK();
    fold_kn_kernel<<<cdiv(4 * E, 32), 256, 0, st>>>(P + lo.fc_w, P + lo.ln2_w, P + lo.ln2_b, P + lo.fc_b, E, 4 * E,
                                                     w.f_fc + (size_t)l * E * 4 * E, fv + 6 * E, fv + 10 * E);
    MMTG_LAUNCH_OK();
  }
  float* hv = w.f_vec + (size_t)d.NL * 14 * E;
  fold_head_kernel<<<cdiv(d.V, 8), 256, 0, st>>>(P + m->off.wte, P + m->off.lnf_w, P + m->off.lnf_b, d.V, E, w.f_wte,
                                                 hv, hv + d.V);
  MMTG_LAUNCH_OK();
  count_launch(2 * d.NL + 1);
  return 0;
}

// Launch: all blocks + lm_head of one decode position. bufs.h holds the block-0 input (projector
// output + wpe + wte[type]); bufs.barrier must be zero at launch (decode_prep resets it).
int decode_mega_launch(const mmtg_model* m, int Lmax, const MegaBufs& bufs, const int* j_ptr, float* logits,
                       cudaStream_t st) {
  const mmtg_dims& d = m->dims;
  MMTG_CHECK_ARG(d.B <= 64 && d.E == 768 && d.NH * 64 == d.E && Lmax <= 1024,
                 "decode megakernel: unsupported shape (B=%d E=%d Lmax=%d)", d.B, d.E, Lmax);
  MMTG_PER_DEVICE_FLAG(attr_set);
  if (!attr_set) {
    MMTG_CUDA_OK(cudaFuncSetAttribute(decode_mega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MG_SMEM));
    attr_set = true;
  }
  MegaParams p;
  p.P = m->params;
  p.W = (const bf16*)m->params_bf16;
  p.off = m->off;
  const uint64_t rows = (uint64_t)d.NL * d.B * d.NH * Lmax;
  MMTG_TRY(make_tmap_bf16_2d(&p.tm_k, bufs.kcache, 64, rows, 64, 64, 64));
  MMTG_TRY(make_tmap_bf16_2d(&p.tm_v, bufs.vcache, 64, rows, 64, 64, 64));
  p.B = d.B; p.E = d.E; p.NH = d.NH; p.NL = d.NL; p.V = d.V; p.Pl = d.P; p.Lmax = Lmax;
  p.w = bufs;
  p.j_ptr = j_ptr;
  p.logits = logits;
  p.trace = g_mega_trace;
  // one CTA per SM, all co-resident (the grid barrier requires it): cooperative launch
  void* args[] = {(void*)&p};
  MMTG_CUDA_OK(cudaLaunchCooperativeKernel((const void*)decode_mega_kernel, dim3(num_sms()), dim3(MG_THREADS), args,
                                           MG_SMEM, st));
  count_launch();
  return 0;
}

}  // namespace mmtg

// Debug: CTA 0 writes a globaltimer stamp (ns) at kernel start, after every grid barrier and at
// its end into `dev_buf` (>= 512 entries; [64..69] = sub-stamps of the last block's phase A, [80..], [240..] = per-CTA arrival at its A->B and B->C barriers);
// pass NULL to disable.
extern "C" int mmtg_decode_set_trace(uint64_t* dev_buf) {
  mmtg::g_mega_trace = (unsigned long long*)dev_buf;
  return 0;
}
