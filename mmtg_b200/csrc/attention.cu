// Fused causal + key-padding softmax attention for the GPT-2 decoder (head_dim 64), forward and
// backward. Replaces HF GPT2Attention's SDPA call (transformers modeling_gpt2.py:54-72,144-226)
// and its autograd backward.
//
// Layout: qkv is the c_attn output [B*L, 3*E] bf16 (q | k | v, heads contiguous 64-wide slices);
// out is [B*L, E] bf16 with heads merged — exactly what c_proj consumes, so no permutes exist.
// Flash-style: 64-query x 64-key tiles, online softmax in the exp2 domain, K/V (fwd) or Q/dO
// (bwd) tiles double-buffered in shared memory with cp.async, XOR-swizzled 16-byte chunks so
// ldmatrix is conflict-free. Tensor-core math is mma.sync m16n8k16 bf16 (legacy HMMA path):
// attention is 2 % of the step's FLOPs; the tcgen05 version is tracked in DESIGN.md.
#include <stdlib.h>

#include "../../include/mmtg_b200.h"
#include "ops.h"
#include "mma_tiles.cuh"

namespace mmtg {

namespace {

constexpr int HD = 64;       // head dim
constexpr int BQ = 64;       // query rows per CTA
constexpr int BKV = 64;      // keys per tile

constexpr float LOG2E = 1.4426950408889634f;

struct AttnParams {
  const bf16* qkv;  // [B*L, 3E]
  const int* kmask;  // [B, L] 1 = attend, 0 = padding key
  bf16* out;        // [B*L, E]
  float* lse;       // [B, NH, L] natural-log LSE of the scaled scores
  const bf16* dout;  // [B*L, E]
  const float* delta;  // [B, NH, L] rowsum(dO * O)
  bf16* dqkv;       // [B*L, 3E]
  int B, L, NH, E;
  float scale;
  // dropout of the probabilities: element ((b*NH+h)*L + q) * Lp + key, Lp = L rounded up to even
  const unsigned long long* drop_seed;
  uint32_t drop_site;
  float drop_p;
};

// pair index (two neighbouring keys share one hash) of element (q, key) of head bh
__device__ __forceinline__ uint32_t attn_pair(int bh, int L, int q, int key) {
  return (uint32_t)(bh * L + q) * (uint32_t)((L + 1) >> 1) + (uint32_t)(key >> 1);
}

// --------------------------------------------------------------------------------------------
// forward
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ATT_THREADS)
attn_fwd_kernel(const AttnParams p) {
  __shared__ __align__(128) bf16 sQ[BQ * HD];
  __shared__ __align__(128) bf16 sK[2][BKV * HD];
  __shared__ __align__(128) bf16 sV[2][BKV * HD];
  __shared__ float sMask[2][BKV];

  const int qb = (gridDim.x - 1) - blockIdx.x;  // heavy (late) query blocks first
  const int bh = blockIdx.y;
  const bool dropping = p.drop_p > 0.f;
  DropKey dkey = {0u, 0u, 0u, 1.f};
  if (dropping) dkey = drop_key(p.drop_seed, p.drop_site, p.drop_p);
  const int b = bh / p.NH, h = bh - b * p.NH;
  const int warp = threadIdx.x >> 5, l = lane_id();
  const long long ld = 3LL * p.E;
  const bf16* base = p.qkv + (long long)b * p.L * ld;
  const int q0 = qb * BQ;

  load_tile_async(sQ, base, ld, q0, p.L, h * HD);
  const int nkv = qb + 1;  // causal: key tiles 0..qb
  auto issue_kv = [&](int t, int buf) {
    load_tile_async(sK[buf], base, ld, t * BKV, p.L, p.E + h * HD);
    load_tile_async(sV[buf], base, ld, t * BKV, p.L, 2 * p.E + h * HD);
    if (threadIdx.x < BKV) {
      const int key = t * BKV + threadIdx.x;
      const bool ok = key < p.L && (p.kmask == nullptr || p.kmask[b * p.L + key] != 0);
      sMask[buf][threadIdx.x] = ok ? 0.f : -INFINITY;
    }
  };
  issue_kv(0, 0);
  cp_async_commit();

  const float sl2 = p.scale * LOG2E;
  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  uint32_t qa[4][4];
  const int row_lo = q0 + warp * 16 + (l >> 2);  // this thread's rows: row_lo, row_lo + 8

  for (int t = 0; t < nkv; ++t) {
    const int buf = t & 1;
    if (t + 1 < nkv) issue_kv(t + 1, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (t == 0) load_a_frags(sQ, warp * 16, qa);

    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
    mma_nt(s, qa, sK[buf]);

    // mask + online softmax (exp2 domain)
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kc = nb * 8 + (l & 3) * 2 + (j & 1);
        const int key = t * BKV + kc;
        const int row = row_lo + (j >> 1) * 8;
        float v = s[nb][j] * sl2 + sMask[buf][kc];
        if (key > row) v = -INFINITY;
        s[nb][j] = v;
        mx[j >> 1] = fmaxf(mx[j >> 1], v);
      }
    }
    float corr[2], mnew[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      mnew[r] = fmaxf(m_run[r], mx[r]);
      const float msafe = (mnew[r] == -INFINITY) ? 0.f : mnew[r];
      corr[r] = exp2f(m_run[r] - msafe);  // m_run = -inf -> 0
      m_run[r] = mnew[r];
      mnew[r] = msafe;
    }
    float rs[2] = {0.f, 0.f};
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float e = exp2f(s[nb][j] - mnew[j >> 1]);
        s[nb][j] = e;
        rs[j >> 1] += e;
      }
      if (dropping) {  // O uses the dropped P; the 1/(1-p) factor is applied at the end
        const int key = t * BKV + nb * 8 + (l & 3) * 2;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const uint32_t bits = drop_bits(dkey, attn_pair(bh, p.L, row_lo + r * 8, key));
          if (!drop_keep_lo(dkey, bits)) s[nb][2 * r] = 0.f;
          if (!drop_keep_hi(dkey, bits)) s[nb][2 * r + 1] = 0.f;
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * corr[r] + rs[r];
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      o[nb][0] *= corr[0]; o[nb][1] *= corr[0];
      o[nb][2] *= corr[1]; o[nb][3] *= corr[1];
    }
    mma_nn(o, s, sV[buf]);
    __syncthreads();  // all warps done with buf before it is refilled next iteration
  }

  // finalize: O /= l, write bf16; LSE = (m + log2(l)) / log2(e)
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = row_lo + r * 8;
    if (row < p.L) {
      const float inv = (l_run[r] > 0.f ? 1.f / l_run[r] : 0.f) * dkey.inv_keep;
      bf16* dst = p.out + ((long long)b * p.L + row) * p.E + h * HD + (l & 3) * 2;
#pragma unroll
      for (int nb = 0; nb < 8; ++nb)
        *reinterpret_cast<uint32_t*>(dst + nb * 8) =
            pack_bf16(o[nb][2 * r] * inv, o[nb][2 * r + 1] * inv);
      if ((l & 3) == 0 && p.lse)
        p.lse[((long long)b * p.NH + h) * p.L + row] =
            l_run[r] > 0.f ? (m_run[r] + log2f(l_run[r])) / LOG2E : -INFINITY;
    }
  }
}

// --------------------------------------------------------------------------------------------
// backward preprocess: delta[b,h,q] = sum_d dO[q,d] * O[q,d]
// --------------------------------------------------------------------------------------------
// One warp per row: lane i-th 8-byte load covers elements [i*128 + lane*4, +4) -> head 2i + (lane>=16).
__global__ void __launch_bounds__(256)
attn_delta_kernel(const bf16* __restrict__ out, const bf16* __restrict__ dout,
                  float* __restrict__ delta, int B, int L, int NH, int E) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= (long long)B * L) return;
  const int l = lane_id();
  const int b = (int)(row / L), q = (int)(row - (long long)b * L);
  const uint2* o = reinterpret_cast<const uint2*>(out + row * E);
  const uint2* d = reinterpret_cast<const uint2*>(dout + row * E);
  for (int i = 0; i * 128 < E; ++i) {
    const uint2 a = __ldg(o + i * 32 + l), g = __ldg(d + i * 32 + l);
    const float2 a0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&a.x));
    const float2 a1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&a.y));
    const float2 g0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&g.x));
    const float2 g1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&g.y));
    float s = a0.x * g0.x + a0.y * g0.y + a1.x * g1.x + a1.y * g1.y;
#pragma unroll
    for (int off = 8; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if ((l & 15) == 0) {
      const int h = 2 * i + (l >> 4);
      delta[((long long)b * NH + h) * L + q] = s;
    }
  }
}

// --------------------------------------------------------------------------------------------
// backward, dK/dV: one CTA per (key tile, b, h); loops over query tiles >= key tile.
// Works on transposed tiles (rows = keys) so both outputs accumulate in registers.
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ATT_THREADS)
attn_bwd_dkdv_kernel(const AttnParams p) {
  extern __shared__ __align__(128) uint8_t att_smem[];
  bf16* sK = reinterpret_cast<bf16*>(att_smem);
  bf16* sV = sK + BKV * HD;
  bf16(*sQ)[BQ * HD] = reinterpret_cast<bf16(*)[BQ * HD]>(sV + BKV * HD);
  bf16(*sdO)[BQ * HD] = reinterpret_cast<bf16(*)[BQ * HD]>(sV + BKV * HD + 2 * BQ * HD);
  float(*sLse)[BQ] = reinterpret_cast<float(*)[BQ]>(sV + BKV * HD + 4 * BQ * HD);
  float(*sDelta)[BQ] = sLse + 2;

  const int kb = blockIdx.x;
  const int bh = blockIdx.y;
  const bool dropping = p.drop_p > 0.f;
  DropKey dkey = {0u, 0u, 0u, 1.f};
  if (dropping) dkey = drop_key(p.drop_seed, p.drop_site, p.drop_p);
  const int b = bh / p.NH, h = bh - b * p.NH;
  const int warp = threadIdx.x >> 5, l = lane_id();
  const long long ld = 3LL * p.E;
  const bf16* base = p.qkv + (long long)b * p.L * ld;
  const bf16* dobase = p.dout + (long long)b * p.L * p.E;
  const int k0 = kb * BKV;
  const int nq = (p.L + BQ - 1) / BQ;

  load_tile_async(sK, base, ld, k0, p.L, p.E + h * HD);
  load_tile_async(sV, base, ld, k0, p.L, 2 * p.E + h * HD);
  auto issue_q = [&](int t, int buf) {
    load_tile_async(sQ[buf], base, ld, t * BQ, p.L, h * HD);
    load_tile_async(sdO[buf], dobase, p.E, t * BQ, p.L, h * HD);
    if (threadIdx.x < BQ) {
      const int q = t * BQ + threadIdx.x;
      const long long o = ((long long)b * p.NH + h) * p.L + q;
      const float lv = q < p.L ? p.lse[o] : -INFINITY;
      sLse[buf][threadIdx.x] = lv == -INFINITY ? INFINITY : lv * LOG2E;  // +inf -> P = 0
      sDelta[buf][threadIdx.x] = q < p.L ? p.delta[o] : 0.f;
    }
  };
  issue_q(kb, 0);
  cp_async_commit();

  const float sl2 = p.scale * LOG2E;
  float dk[8][4], dv[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) dk[i][j] = dv[i][j] = 0.f;
  uint32_t ka[4][4], va[4][4];
  const int key_lo = k0 + warp * 16 + (l >> 2);
  bool keyok[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int key = key_lo + r * 8;
    keyok[r] = key < p.L && (p.kmask == nullptr || p.kmask[b * p.L + key] != 0);
  }

  for (int t = kb; t < nq; ++t) {
    const int buf = (t - kb) & 1;
    if (t + 1 < nq) issue_q(t + 1, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (t == kb) {
      load_a_frags(sK, warp * 16, ka);
      load_a_frags(sV, warp * 16, va);
    }
    // S^T = K Q^T  (rows = keys, cols = queries)
    float st[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) st[i][j] = 0.f;
    mma_nt(st, ka, sQ[buf]);
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int qc = nb * 8 + (l & 3) * 2 + (j & 1);
        const int q = t * BQ + qc;
        const int key = key_lo + (j >> 1) * 8;
        const bool ok = keyok[j >> 1] && key <= q;
        st[nb][j] = ok ? exp2f(st[nb][j] * sl2 - sLse[buf][qc]) : 0.f;  // P^T
      }
    }
    // with dropout D = keep/(1-p): dV += (P.D)^T dO and dS = P (D.dP - delta)
    float dm[8][4];
    if (dropping) {
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int q = t * BQ + nb * 8 + (l & 3) * 2 + (j & 1);
          const int key = key_lo + (j >> 1) * 8;
          const uint32_t bits = drop_bits(dkey, attn_pair(bh, p.L, q, key));
          const bool keep = (key & 1) ? drop_keep_hi(dkey, bits) : drop_keep_lo(dkey, bits);
          dm[nb][j] = keep ? dkey.inv_keep : 0.f;
        }
      }
      float pd[8][4];
#pragma unroll
      for (int nb = 0; nb < 8; ++nb)
#pragma unroll
        for (int j = 0; j < 4; ++j) pd[nb][j] = st[nb][j] * dm[nb][j];
      mma_nn(dv, pd, sdO[buf]);
    } else {
      mma_nn(dv, st, sdO[buf]);  // dV += P^T dO
    }
    // dP^T = V dO^T
    float dpt[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dpt[i][j] = 0.f;
    mma_nt(dpt, va, sdO[buf]);
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int qc = nb * 8 + (l & 3) * 2 + (j & 1);
        const float dpd = dropping ? dpt[nb][j] * dm[nb][j] : dpt[nb][j];
        dpt[nb][j] = st[nb][j] * (dpd - sDelta[buf][qc]);  // dS^T
      }
    }
    mma_nn(dk, dpt, sQ[buf]);  // dK += dS^T Q
    __syncthreads();
  }

#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int key = key_lo + r * 8;
    if (key < p.L) {
      bf16* dst = p.dqkv + ((long long)b * p.L + key) * ld + h * HD + (l & 3) * 2;
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        *reinterpret_cast<uint32_t*>(dst + p.E + nb * 8) =
            pack_bf16(dk[nb][2 * r] * p.scale, dk[nb][2 * r + 1] * p.scale);
        *reinterpret_cast<uint32_t*>(dst + 2 * p.E + nb * 8) =
            pack_bf16(dv[nb][2 * r], dv[nb][2 * r + 1]);
      }
    }
  }
}

// --------------------------------------------------------------------------------------------
// backward, dQ: one CTA per (query tile, b, h); loops over key tiles <= query tile.
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ATT_THREADS)
attn_bwd_dq_kernel(const AttnParams p) {
  extern __shared__ __align__(128) uint8_t att_smem[];
  bf16* sQ = reinterpret_cast<bf16*>(att_smem);
  bf16* sdO = sQ + BQ * HD;
  bf16(*sK)[BKV * HD] = reinterpret_cast<bf16(*)[BKV * HD]>(sdO + BQ * HD);
  bf16(*sV)[BKV * HD] = reinterpret_cast<bf16(*)[BKV * HD]>(sdO + BQ * HD + 2 * BKV * HD);
  float(*sMask)[BKV] = reinterpret_cast<float(*)[BKV]>(sdO + BQ * HD + 4 * BKV * HD);

  const int qb = (gridDim.x - 1) - blockIdx.x;
  const int bh = blockIdx.y;
  const bool dropping = p.drop_p > 0.f;
  DropKey dkey = {0u, 0u, 0u, 1.f};
  if (dropping) dkey = drop_key(p.drop_seed, p.drop_site, p.drop_p);
  const int b = bh / p.NH, h = bh - b * p.NH;
  const int warp = threadIdx.x >> 5, l = lane_id();
  const long long ld = 3LL * p.E;
  const bf16* base = p.qkv + (long long)b * p.L * ld;
  const bf16* dobase = p.dout + (long long)b * p.L * p.E;
  const int q0 = qb * BQ;

  load_tile_async(sQ, base, ld, q0, p.L, h * HD);
  load_tile_async(sdO, dobase, p.E, q0, p.L, h * HD);
  const int nkv = qb + 1;
  auto issue_kv = [&](int t, int buf) {
    load_tile_async(sK[buf], base, ld, t * BKV, p.L, p.E + h * HD);
    load_tile_async(sV[buf], base, ld, t * BKV, p.L, 2 * p.E + h * HD);
    if (threadIdx.x < BKV) {
      const int key = t * BKV + threadIdx.x;
      const bool ok = key < p.L && (p.kmask == nullptr || p.kmask[b * p.L + key] != 0);
      sMask[buf][threadIdx.x] = ok ? 0.f : -INFINITY;
    }
  };
  issue_kv(0, 0);
  cp_async_commit();

  const float sl2 = p.scale * LOG2E;
  float dq[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) dq[i][j] = 0.f;
  uint32_t qa[4][4], doa[4][4];
  const int row_lo = q0 + warp * 16 + (l >> 2);
  float lse2[2], dl[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = row_lo + r * 8;
    const long long o = ((long long)b * p.NH + h) * p.L + row;
    const float lv = row < p.L ? p.lse[o] : -INFINITY;
    lse2[r] = lv == -INFINITY ? INFINITY : lv * LOG2E;
    dl[r] = row < p.L ? p.delta[o] : 0.f;
  }

  for (int t = 0; t < nkv; ++t) {
    const int buf = t & 1;
    if (t + 1 < nkv) issue_kv(t + 1, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (t == 0) {
      load_a_frags(sQ, warp * 16, qa);
      load_a_frags(sdO, warp * 16, doa);
    }
    float s[8][4], dp[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = dp[i][j] = 0.f;
    mma_nt(s, qa, sK[buf]);
    mma_nt(dp, doa, sV[buf]);  // dP = dO V^T
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kc = nb * 8 + (l & 3) * 2 + (j & 1);
        const int key = t * BKV + kc;
        const int row = row_lo + (j >> 1) * 8;
        const float pv = (key <= row && sMask[buf][kc] == 0.f)
                             ? exp2f(s[nb][j] * sl2 - lse2[j >> 1]) : 0.f;
        float dpd = dp[nb][j];
        if (dropping) {
          const uint32_t bits = drop_bits(dkey, attn_pair(bh, p.L, row, key));
          dpd = ((key & 1) ? drop_keep_hi(dkey, bits) : drop_keep_lo(dkey, bits)) ? dpd * dkey.inv_keep : 0.f;
        }
        s[nb][j] = pv * (dpd - dl[j >> 1]);  // dS
      }
    }
    mma_nn(dq, s, sK[buf]);  // dQ += dS K
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = row_lo + r * 8;
    if (row < p.L) {
      bf16* dst = p.dqkv + ((long long)b * p.L + row) * ld + h * HD + (l & 3) * 2;
#pragma unroll
      for (int nb = 0; nb < 8; ++nb)
        *reinterpret_cast<uint32_t*>(dst + nb * 8) =
            pack_bf16(dq[nb][2 * r] * p.scale, dq[nb][2 * r + 1] * p.scale);
    }
  }
}

}  // namespace

void count_launch(int n = 1);

int attn_fwd_tc(const bf16* qkv, const int* kmask, bf16* out, float* lse, int B, int L, int NH,
                cudaStream_t st, const DropSpec* drop);
static inline bool drop_on(const DropSpec* d) { return d && d->seed && d->p > 0.f; }

// impl: 0 = default (tcgen05 unless MMTG_ATTN_TC=0), 1 = mma.sync kernels, 2 = tcgen05 kernels
int attn_fwd_impl(const bf16* qkv, const int* kmask, bf16* out, float* lse, int B, int L, int NH,
                  int impl, cudaStream_t st, const DropSpec* drop = nullptr) {
  static const bool env_tc = []() {
    const char* e = getenv("MMTG_ATTN_TC");
    return !(e && e[0] == '0');
  }();
  const bool use_tc = impl == 2 || (impl == 0 && env_tc);
  if (use_tc) return attn_fwd_tc(qkv, kmask, out, lse, B, L, NH, st, drop);
  AttnParams p{};
  p.qkv = qkv; p.kmask = kmask; p.out = out; p.lse = lse;
  p.B = B; p.L = L; p.NH = NH; p.E = NH * HD; p.scale = 0.125f;
  if (drop_on(drop)) { p.drop_seed = drop->seed; p.drop_site = drop->site; p.drop_p = drop->p; }
  dim3 grid(cdiv(L, BQ), B * NH);
  // causal FLOPs: 2 GEMMs x 2 x 64 x L(L+1)/2 per (b, h)
  ProfScope prof(1, 4.0 * 64 * 0.5 * L * (L + 1.0) * B * NH, 2.0 * 4 * B * L * NH * 64, st);
  attn_fwd_kernel<<<grid, ATT_THREADS, 0, st>>>(p);
  MMTG_LAUNCH_OK();
  count_launch();
  return 0;
}

int attn_bwd_tc(const bf16* qkv, const int* kmask, const bf16* out, const bf16* dout, const float* lse,
                bf16* dqkv, int B, int L, int NH, cudaStream_t st, const DropSpec* drop, float* dbias);

int attn_bwd_tiled_tc(const bf16* qkv, const int* kmask, const bf16* dout, const float* lse, const float* delta,
                      bf16* dqkv, int B, int L, int NH, cudaStream_t st, const DropSpec* drop, float* dbias);

int attn_fwd(const bf16* qkv, const int* kmask, bf16* out, float* lse, int B, int L, int NH,
             cudaStream_t st, const DropSpec* drop) {
  return attn_fwd_impl(qkv, kmask, out, lse, B, L, NH, 0, st, drop);
}

// dbias (optional, [3E]): += column sums of dqkv, the c_attn bias gradient - fused into the tcgen05
// kernel's drains, a separate pass behind the mma.sync kernels
int attn_bwd_impl(const bf16* qkv, const int* kmask, const bf16* out, const bf16* dout, const float* lse,
                  float* delta, bf16* dqkv, int B, int L, int NH, int impl, cudaStream_t st,
                  const DropSpec* drop = nullptr, float* dbias = nullptr) {
  // L <= 256: whole-head tcgen05 kernel (attention_tc.cu); longer sequences (config 5),
  // impl == 1 and MMTG_ATTN_BWD_TC=0 use the tiled mma.sync kernels below
  static const bool env_tc = []() {
    const char* e = getenv("MMTG_ATTN_BWD_TC");
    return !(e && e[0] == '0');
  }();
  const bool use_tc = impl == 2 || (impl == 0 && env_tc);
  AttnParams p{};
  p.qkv = qkv; p.kmask = kmask; p.lse = const_cast<float*>(lse); p.dout = dout; p.delta = delta;
  p.dqkv = dqkv; p.B = B; p.L = L; p.NH = NH; p.E = NH * HD; p.scale = 0.125f;
  if (drop_on(drop)) { p.drop_seed = drop->seed; p.drop_site = drop->site; p.drop_p = drop->p; }
  ProfScope prof(1, 2.5 * 4.0 * 64 * 0.5 * L * (L + 1.0) * B * NH, 2.0 * 8 * B * L * NH * 64, st);
  // (the whole-head tcgen05 kernel forms delta = rowsum(dO * O) in its own prologue)
  if (use_tc && L <= 256) return attn_bwd_tc(qkv, kmask, out, dout, lse, dqkv, B, L, NH, st, drop, dbias);
  const long long warps = (long long)B * L;
  attn_delta_kernel<<<(unsigned)cdivll(warps * 32, 256), 256, 0, st>>>(out, dout, delta, B, L, NH, p.E);
  MMTG_LAUNCH_OK();
  // 256 < L <= 1024: tiled tcgen05 kernels (dK/dV per key block, dQ per query block)
  if (use_tc && L <= 1024) {
    count_launch();
    return attn_bwd_tiled_tc(qkv, kmask, dout, lse, delta, dqkv, B, L, NH, st, drop, dbias);
  }
  dim3 grid(cdiv(L, BQ), B * NH);
  constexpr int BWD_SMEM = 6 * BQ * HD * 2 + 4 * BQ * 4;  // 6 bf16 tiles + 4 x 64 floats
  MMTG_PER_DEVICE_FLAG(attr_set);
  if (!attr_set) {
    MMTG_CUDA_OK(cudaFuncSetAttribute(attn_bwd_dkdv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
    MMTG_CUDA_OK(cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
    attr_set = true;
  }
  attn_bwd_dkdv_kernel<<<grid, ATT_THREADS, BWD_SMEM, st>>>(p);
  MMTG_LAUNCH_OK();
  attn_bwd_dq_kernel<<<grid, ATT_THREADS, BWD_SMEM, st>>>(p);
  MMTG_LAUNCH_OK();
  count_launch(3);
  if (dbias) MMTG_TRY(colsum(dqkv, 1, 3 * p.E, nullptr, 0, dbias, B * L, 3 * p.E, st));
  return 0;
}

int attn_bwd(const bf16* qkv, const int* kmask, const bf16* out, const bf16* dout, const float* lse,
             float* delta, bf16* dqkv, int B, int L, int NH, cudaStream_t st, const DropSpec* drop, float* dbias) {
  return attn_bwd_impl(qkv, kmask, out, dout, lse, delta, dqkv, B, L, NH, 0, st, drop, dbias);
}

}  // namespace mmtg

using namespace mmtg;

extern "C" int mmtg_attn_fwd_ex(const void* qkv, const int32_t* key_mask, void* out, float* lse,
                                int32_t B, int32_t L, int32_t n_head, int32_t impl, void* stream) {
  MMTG_CHECK_ARG(qkv && out && B > 0 && L > 0 && n_head > 0 && impl >= 0 && impl <= 2, "bad attention args");
  return attn_fwd_impl((const bf16*)qkv, key_mask, (bf16*)out, lse, B, L, n_head, impl, (cudaStream_t)stream);
}
extern "C" int mmtg_attn_bwd_ex(const void* qkv, const int32_t* key_mask, const void* out,
                                const void* dout, const float* lse, float* delta_ws, void* dqkv,
                                int32_t B, int32_t L, int32_t n_head, int32_t impl, void* stream) {
  MMTG_CHECK_ARG(qkv && out && dout && lse && delta_ws && dqkv && B > 0 && L > 0 && n_head > 0 && impl >= 0 && impl <= 2,
                 "bad attention bwd args");
  MMTG_CHECK_ARG(!(impl == 2 && L > 1024), "tcgen05 attention backward handles L <= 1024");
  return attn_bwd_impl((const bf16*)qkv, key_mask, (const bf16*)out, (const bf16*)dout, lse, delta_ws,
                       (bf16*)dqkv, B, L, n_head, impl, (cudaStream_t)stream);
}

// dropout variants (tcgen05 kernels): probabilities masked by (seed, site) with keep-rate 1 - p
extern "C" int mmtg_attn_fwd_drop(const void* qkv, const int32_t* key_mask, void* out, float* lse, int32_t B,
                                  int32_t L, int32_t n_head, const uint64_t* seed_dev, uint32_t site, float p,
                                  int32_t impl, void* stream) {
  MMTG_CHECK_ARG(qkv && out && B > 0 && L > 0 && n_head > 0 && p >= 0.f && p < 1.f && impl >= 0 && impl <= 2,
                 "bad attention args");
  DropSpec d{(const unsigned long long*)seed_dev, site, p, 0};
  return attn_fwd_impl((const bf16*)qkv, key_mask, (bf16*)out, lse, B, L, n_head, impl, (cudaStream_t)stream, &d);
}
extern "C" int mmtg_attn_bwd_drop(const void* qkv, const int32_t* key_mask, const void* out, const void* dout,
                                  const float* lse, float* delta_ws, void* dqkv, int32_t B, int32_t L,
                                  int32_t n_head, const uint64_t* seed_dev, uint32_t site, float p, int32_t impl,
                                  void* stream) {
  MMTG_CHECK_ARG(qkv && out && dout && lse && delta_ws && dqkv && B > 0 && L > 0 && n_head > 0 && p >= 0.f &&
                     p < 1.f && impl >= 0 && impl <= 2, "bad attention bwd args");
  MMTG_CHECK_ARG(!(impl == 2 && L > 1024), "tcgen05 attention backward handles L <= 1024");
  DropSpec d{(const unsigned long long*)seed_dev, site, p, 0};
  return attn_bwd_impl((const bf16*)qkv, key_mask, (const bf16*)out, (const bf16*)dout, lse, delta_ws,
                       (bf16*)dqkv, B, L, n_head, impl, (cudaStream_t)stream, &d);
}

extern "C" int mmtg_attn_fwd(const void* qkv, const int32_t* key_mask, void* out, float* lse,
                             int32_t B, int32_t L, int32_t n_head, void* stream) {
  MMTG_CHECK_ARG(qkv && out && B > 0 && L > 0 && n_head > 0, "bad attention args");
  return attn_fwd((const bf16*)qkv, key_mask, (bf16*)out, lse, B, L, n_head, (cudaStream_t)stream);
}

extern "C" int mmtg_attn_bwd(const void* qkv, const int32_t* key_mask, const void* out,
                             const void* dout, const float* lse, float* delta_ws, void* dqkv,
                             int32_t B, int32_t L, int32_t n_head, void* stream) {
  MMTG_CHECK_ARG(qkv && out && dout && lse && delta_ws && dqkv && B > 0 && L > 0 && n_head > 0,
                 "bad attention bwd args");
  return attn_bwd((const bf16*)qkv, key_mask, (const bf16*)out, (const bf16*)dout, lse, delta_ws,
                  (bf16*)dqkv, B, L, n_head, (cudaStream_t)stream);
}
