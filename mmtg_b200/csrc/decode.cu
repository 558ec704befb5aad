// KV-cached autoregressive decoding for MMTG generation (src/generate.py:97-145).
//
// The reference re-runs the whole model on the whole prefix for every token (no KV cache,
// batch 1). Every per-position input of its inference branch (token, fused context k = j/44,
// position, type id, key mask; src/model.py:291-326) depends only on tokens <= j, so caching
// K/V per layer is semantically exact. A batch of B rows here is B independent batch-1
// reference runs (the reference derives type ids / masks from row 0 only).
//
// Per step (one new token per row):  prep (embedding + type/mask rules) -> projector GEMMs ->
// 12 x [LN, c_attn GEMM, cached attention, c_proj GEMM(+res), LN, c_fc GEMM(gelu), c_proj GEMM
// (+res)] -> ln_f -> lm_head GEMM -> fused sampler (repetition penalty, temperature, bans,
// forced [#EOS#]/[#START#], PAD continuation, top-k, top-p, multinomial). The step index lives
// in device memory so the whole step is CUDA-graph capturable and replayable.
#include <string.h>

#include "../../include/mmtg_b200.h"
#include "ops.h"

namespace mmtg {

void count_launch(int n = 1);
int skinny_gemm(const mmtg_gemm_args* a, cudaStream_t st);

namespace {

// decode GEMMs: weight-streaming skinny kernel for <= 64 rows, tcgen05 kernel otherwise
inline int decode_gemm(const mmtg_gemm_args* a, void* stream) {
  if (a->M <= 64) return skinny_gemm(a, (cudaStream_t)stream);
  return mmtg_gemm_bf16(a, stream);
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct DWs {
  bf16 *kcache, *vcache;  // [NL][B][NH][Lmax][64]
  float* ctx;             // [S*B, Dw]
  int* keymask;           // [B, Lmax]
  int* types;             // [B]
  int* posidx;            // [B]
  bf16 *emb16, *p1, *x16, *qkv16, *att16, *a16;
  float *h, *h2;
  MegaBufs mega;  // fused step: accumulators, folded weights, grid barrier
  size_t bytes;
};

void carve_decode(const mmtg_dims& d, int Lmax, uint8_t* base, DWs* w) {
  size_t off = 0;
  auto take = [&](size_t bytes) -> uint8_t* {
    uint8_t* p = base ? base + off : nullptr;
    off += align_up(bytes, 256);
    return p;
  };
  const size_t B = d.B, E = d.E;
  const size_t cache = (size_t)d.NL * B * d.NH * Lmax * 64;
  w->kcache = (bf16*)take(cache * 2);
  w->vcache = (bf16*)take(cache * 2);
  w->ctx = (float*)take((size_t)d.S * B * d.Dw * 4);
  w->keymask = (int*)take(B * Lmax * 4);
  w->types = (int*)take(B * 4);
  w->posidx = (int*)take(B * 4);
  w->emb16 = (bf16*)take(B * d.Dw * 2);
  w->p1 = (bf16*)take(B * d.He * 2);
  w->x16 = (bf16*)take(B * E * 2);
  w->qkv16 = (bf16*)take(B * 3 * E * 2);
  w->att16 = (bf16*)take(B * E * 2);
  w->a16 = (bf16*)take(B * 4 * E * 2);
  w->h = (float*)take(B * E * 4);
  w->h2 = (float*)take(B * E * 4);
  MegaBufs& g = w->mega;
  g.h = w->h; g.h2 = w->h2; g.att16 = w->att16; g.kcache = w->kcache; g.vcache = w->vcache; g.keymask = w->keymask;
  g.h_alt = (float*)take(B * E * 4);
  g.qkv_acc = (float*)take(B * 3 * E * 4);
  g.u_acc = (float*)take(B * 4 * E * 4);
  g.row_stats = (float*)take(2 * 64 * 2 * 4);
  g.barrier = (unsigned int*)take(256);
  g.f_attn = (bf16*)take((size_t)d.NL * E * 3 * E * 2);
  g.f_fc = (bf16*)take((size_t)d.NL * E * 4 * E * 2);
  g.f_wte = (bf16*)take((size_t)d.V * E * 2);
  g.f_vec = (float*)take(((size_t)d.NL * 14 * E + 2 * (size_t)d.V) * 4);
  w->bytes = off;
}

// Per row: embedding of the new token (+ fused context of its sentence pair), inference-branch
// type id (src/model.py:300-306) and key mask (:309-312), absolute position.
__global__ void __launch_bounds__(256)
decode_prep_kernel(const int* __restrict__ gen, int gen_ld, const int* __restrict__ j_ptr,
                   const float* __restrict__ table, const float* __restrict__ ctx,
                   bf16* __restrict__ emb16, int* __restrict__ types, int* __restrict__ posidx,
                   int* __restrict__ keymask, int B, int P, int S, int sent_len, int n_sent, int D,
                   int Lmax, unsigned int* __restrict__ barrier, int table_rows) {
  const int b = blockIdx.x;
  if (b == 0 && threadIdx.x == 0) *barrier = 0u;  // the megakernel's grid barrier starts from zero
  const int j = *j_ptr;
  const int tok = gen[b * gen_ld + j];
  if ((unsigned)tok >= (unsigned)table_rows) {
    if (threadIdx.x == 0) printf("mmtg: token id %d (row %d, step %d) is outside the token table [0, %d)\n", tok, b, j, table_rows);
    __trap();
  }
  if (threadIdx.x == 0) {
    int ty;
    const int r = (j + 1) % sent_len;
    if (r == 0 || r == 1 || tok == 0) {
      ty = 0;
    } else {
      const int s = j / sent_len;  // type list [1..n_sent, 1]
      ty = s < n_sent ? s + 1 : 1;
    }
    types[b] = ty;
    posidx[b] = P + j;
    keymask[b * Lmax + P + j] = tok != 0 ? 1 : 0;
  }
  const int k = j / (2 * sent_len);
  const float* trow = table + (long long)tok * D;
  const float* crow = k < S ? ctx + ((long long)k * B + b) * D : nullptr;
  for (int c = threadIdx.x * 4; c < D; c += blockDim.x * 4) {
    float4 a = __ldg(reinterpret_cast<const float4*>(trow + c));
    if (crow) {
      const float4 e = __ldg(reinterpret_cast<const float4*>(crow + c));
      a.x += e.x; a.y += e.y; a.z += e.z; a.w += e.w;
    }
    __nv_bfloat162 lo = __floats2bfloat162_rn(a.x, a.y), hi = __floats2bfloat162_rn(a.z, a.w);
    *reinterpret_cast<uint2*>(emb16 + (long long)b * D + c) =
        make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
  }
}

// One 128-thread block per (row, head): append this step's K/V to the cache, then attend over
// keys 0..pos with the key-padding mask. The keys are split across the 4 warps (each lane scores
// whole keys: one 128-byte row per lane), partial (max, sum, output) merged in shared memory.
__global__ void __launch_bounds__(128)
decode_attn_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ kc, bf16* __restrict__ vc,
                   const int* __restrict__ keymask, const int* __restrict__ j_ptr,
                   bf16* __restrict__ out, int B, int NH, int P, int Lmax) {
  extern __shared__ float s_scores[];  // [Lmax] scores, then [4][66] partials
  float* s_part = s_scores + Lmax;
  const int w = blockIdx.x;
  const int b = w / NH, h = w - b * NH;
  const int warp = threadIdx.x >> 5, l = lane_id();
  const int pos = P + *j_ptr;
  const int E = NH * 64;
  const bf16* row = qkv + (long long)b * 3 * E + h * 64;
  bf16* kbase = kc + ((long long)b * NH + h) * Lmax * 64;
  bf16* vbase = vc + ((long long)b * NH + h) * Lmax * 64;
  if (warp == 0) {  // append (visible to the other warps after the barrier below)
    reinterpret_cast<uint32_t*>(kbase + (long long)pos * 64)[l] = reinterpret_cast<const uint32_t*>(row + E)[l];
    reinterpret_cast<uint32_t*>(vbase + (long long)pos * 64)[l] = reinterpret_cast<const uint32_t*>(row + 2 * E)[l];
  }
  float q[64];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint4 u = *reinterpret_cast<const uint4*>(row + i * 8);
    const uint32_t wds[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wds[e]));
      q[i * 8 + 2 * e] = f.x;
      q[i * 8 + 2 * e + 1] = f.y;
    }
  }
  __syncthreads();
  // scores: thread t handles keys t, t+128, ...
  float mx = -INFINITY;
  for (int key = threadIdx.x; key <= pos; key += 128) {
    float d = -INFINITY;
    if (keymask[b * Lmax + key] != 0) {
      const uint4* kr = reinterpret_cast<const uint4*>(kbase + (long long)key * 64);
      d = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint4 u = kr[i];
        const uint32_t wds[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wds[e]));
          d += q[i * 8 + 2 * e] * f.x + q[i * 8 + 2 * e + 1] * f.y;
        }
      }
      d *= 0.125f;
    }
    s_scores[key] = d;
    mx = fmaxf(mx, d);
  }
  mx = warp_max(mx);
  if (l == 0) s_part[warp] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(s_part[0], s_part[1]), fmaxf(s_part[2], s_part[3]));
  if (mx == -INFINITY) mx = 0.f;
  __syncthreads();
  float sum = 0.f;
  for (int key = threadIdx.x; key <= pos; key += 128) {
    const float e = __expf(s_scores[key] - mx);
    s_scores[key] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  if (l == 0) s_part[warp] = sum;
  __syncthreads();
  sum = s_part[0] + s_part[1] + s_part[2] + s_part[3];
  const float inv = sum > 0.f ? 1.f / sum : 0.f;
  __syncthreads();
  // output: warp w sums keys w, w+4, ...; lane l owns dims 2l, 2l+1
  float2 acc = make_float2(0.f, 0.f);
  for (int key = warp; key <= pos; key += 4) {
    const float pj = s_scores[key];
    if (pj != 0.f) {
      const float2 v = __bfloat1622float2(
          reinterpret_cast<const __nv_bfloat162*>(vbase + (long long)key * 64)[l]);
      acc.x += pj * v.x;
      acc.y += pj * v.y;
    }
  }
  s_part[4 + warp * 64 + 2 * l] = acc.x;
  s_part[4 + warp * 64 + 2 * l + 1] = acc.y;
  __syncthreads();
  if (warp == 0) {
    float ox = 0.f, oy = 0.f;
#pragma unroll
    for (int ww = 0; ww < 4; ++ww) {
      ox += s_part[4 + ww * 64 + 2 * l];
      oy += s_part[4 + ww * 64 + 2 * l + 1];
    }
    reinterpret_cast<__nv_bfloat162*>(out + (long long)b * E + h * 64)[l] = __floats2bfloat162_rn(ox * inv, oy * inv);
  }
}

// Copy the prefix K/V produced by the training-style prefill forward ([B*Lp, 3E] per layer) into
// the cache. grid = (B*Lp, NH), 32 threads.
__global__ void kv_scatter_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ kc,
                                  bf16* __restrict__ vc, int B, int Lp, int NH, int Lmax) {
  const int row = blockIdx.x, h = blockIdx.y, l = threadIdx.x;
  const int b = row / Lp, pos = row - b * Lp;
  const int E = NH * 64;
  const bf16* src = qkv + (long long)row * 3 * E + h * 64;
  const long long dst = (((long long)b * NH + h) * Lmax + pos) * 64;
  reinterpret_cast<uint32_t*>(kc + dst)[l] = reinterpret_cast<const uint32_t*>(src + E)[l];
  reinterpret_cast<uint32_t*>(vc + dst)[l] = reinterpret_cast<const uint32_t*>(src + 2 * E)[l];
}

// ------------------------------------------------------------------------------------------
// Fused sampler, one 1024-thread block per row (src/generate.py:118-142 + 64-94).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

struct ArgMax {
  float v;
  int i;
};
__device__ __forceinline__ ArgMax argmax_merge(ArgMax a, ArgMax b) {
  // larger value wins; ties -> smaller index (deterministic)
  if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
  return a;
}
__device__ ArgMax block_argmax(const float* s, int V, ArgMax* red) {
  ArgMax m{-INFINITY, 0x7fffffff};
  for (int c = threadIdx.x; c < V; c += blockDim.x) m = argmax_merge(m, ArgMax{s[c], c});
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ArgMax t{__shfl_xor_sync(0xffffffffu, m.v, o), __shfl_xor_sync(0xffffffffu, m.i, o)};
    m = argmax_merge(m, t);
  }
  if (lane_id() == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    ArgMax t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : ArgMax{-INFINITY, 0x7fffffff};
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ArgMax u{__shfl_xor_sync(0xffffffffu, t.v, o), __shfl_xor_sync(0xffffffffu, t.i, o)};
      t = argmax_merge(t, u);
    }
    if (threadIdx.x == 0) red[0] = t;
  }
  __syncthreads();
  const ArgMax r = red[0];
  __syncthreads();
  return r;
}

constexpr int MAX_SURV = 1024;  // top-k survivors kept in shared memory (pure top-p has no cap)

// order-preserving float -> uint32 key (larger float <=> larger key; -inf is the smallest real key)
__device__ __forceinline__ uint32_t float_key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// block-wide sum, result broadcast to every thread (`red`: >= 32 floats of shared memory)
__device__ float block_sum(float a, float* red) {
  a = warp_sum(a);
  if (lane_id() == 0) red[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) red[0] = t;
  }
  __syncthreads();
  const float r = red[0];
  __syncthreads();
  return r;
}

// Multinomial draw over {c : s[c] >= thr} with weights exp(s[c] - mx); u01 in [0, 1). Threads own
// contiguous chunks, so the prefix order is the vocabulary order (any fixed order is a valid
// inverse-CDF draw). Returns the picked id to every thread.
__device__ int block_draw(const float* s, int V, uint32_t thr_key, float mx, float u01, float* red, int* pick_slot) {
  const int per = (V + blockDim.x - 1) / blockDim.x;
  const int c0 = threadIdx.x * per, c1 = min(V, c0 + per);
  float local = 0.f;
  for (int c = c0; c < c1; ++c)
    if (float_key(s[c]) >= thr_key && s[c] != -INFINITY) local += __expf(s[c] - mx);
  // inclusive scan of the per-thread sums: warp scan + scan of the warp totals
  float incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_up_sync(0xffffffffu, incl, o);
    if ((int)lane_id() >= o) incl += t;
  }
  if (lane_id() == 31) red[threadIdx.x >> 5] = incl;
  if (threadIdx.x == 0) *pick_slot = -1;
  __syncthreads();
  float base = 0.f, total = 0.f;
  const int nw = blockDim.x >> 5;
  for (int w = 0; w < nw; ++w) {
    if (w < (int)(threadIdx.x >> 5)) base += red[w];
    total += red[w];
  }
  const float u = u01 * total;
  const float lo = base + incl - local, hi = base + incl;
  if (local > 0.f && u >= lo && u < hi) {
    float c2 = lo;
    int pick = -1;
    for (int c = c0; c < c1; ++c)
      if (float_key(s[c]) >= thr_key && s[c] != -INFINITY) {
        c2 += __expf(s[c] - mx);
        pick = c;
        if (u < c2) break;
      }
    *pick_slot = pick;  // intervals are disjoint: at most one writer
  }
  __syncthreads();
  int r = *pick_slot;
  if (r < 0) {  // u fell on a rounding gap at the very top: take the last kept id
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int c = V - 1; c >= 0; --c)
        if (float_key(s[c]) >= thr_key && s[c] != -INFINITY) {
          *pick_slot = c;
          break;
        }
    }
    __syncthreads();
    r = *pick_slot;
  }
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(1024)
sample_rows_kernel(const float* __restrict__ logits, long long ld, int* __restrict__ gen, int gen_ld,
                   int* __restrict__ j_ptr, int ban_specials, int V, int sent_len, float temperature,
                   int top_k, float top_p, float rep_penalty, unsigned long long seed,
                   const unsigned long long* __restrict__ seed_dev, float* __restrict__ dbg_probs) {
  if (seed_dev) seed = seed_dev[0];  // device-side seed: the launch stays CUDA-graph replayable
  extern __shared__ float s[];  // [V] working logits
  __shared__ ArgMax red[32];
  __shared__ float fred[32];
  __shared__ float sv[MAX_SURV];
  __shared__ int si[MAX_SURV];
  __shared__ int s_n;
  const int b = blockIdx.x;
  const int i = *j_ptr;  // reference loop index: decides token at position i + 1
  int* g = gen + (long long)b * gen_ld;
  int next = -1;
  if (i > 0 && (i + 2) % sent_len == 0) next = 2;        // forced [#EOS#]   (generate.py:118-120)
  else if (i > 0 && (i + 2) % sent_len == 1) next = 1;   // forced [#START#] (generate.py:121-123)
  else if (g[i] == 0 && !dbg_probs) next = 0;            // PAD continuation (generate.py:137-138)
  if (next < 0) {
    const float* z = logits + (long long)b * ld;
    for (int c = threadIdx.x; c < V; c += blockDim.x) s[c] = z[c];
    __syncthreads();
    if (threadIdx.x == 0) {
      // repetition penalty: plain division, once per OCCURRENCE, ids 0 and 102 exempt
      if (rep_penalty != 1.0f)
        for (int t = 0; t <= i; ++t) {
          const int id = g[t];
          if (id != 0 && id != 102 && id < V) s[id] = s[id] / rep_penalty;
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < V; c += blockDim.x) s[c] = s[c] / temperature;
    __syncthreads();
    if (threadIdx.x == 0 && ban_specials) {
      s[1] = -INFINITY; s[2] = -INFINITY; s[100] = -INFINITY; s[102] = -INFINITY;
    }
    __syncthreads();
    const uint64_t rbits = splitmix64(seed ^ splitmix64(((uint64_t)b << 32) | (uint32_t)i));
    const float u01 = (float)(rbits >> 40) * (1.0f / 16777216.0f);
    const int kk = top_k > 0 ? min(top_k, V) : 0;
    if (kk == 0) {
      // ---- no top-k: pure nucleus (top_p > 0) or plain softmax sampling (top_p == 0) ----
      // generate.py:81-92 keeps, in descending order, every token whose PRECEDING cumulative
      // probability is <= p (the first always). With F(v) = sum_{z_j > v} softmax(z)_j that set is
      // {c : F(z_c) <= p} = {c : z_c >= t*}, t* the smallest float with F(t*) <= p: found by
      // bisection over the order-preserving integer keys (32 block reductions), no sort and no
      // survivor cap. Exact ties at the threshold are kept or dropped together.
      const ArgMax m = block_argmax(s, V, red);
      uint32_t thr = 0u;  // key threshold: keep {c : key(s[c]) >= thr}, -inf excluded
      float zsum = 0.f;
      {
        float a = 0.f;
        for (int c = threadIdx.x; c < V; c += blockDim.x) a += __expf(s[c] - m.v);
        zsum = block_sum(a, fred);
      }
      if (top_p > 0.f) {
        uint32_t lo = 0u, hi = float_key(m.v);  // F(key(max)) = 0 <= p: hi always satisfies
        while (lo < hi) {
          const uint32_t mid = lo + ((hi - lo) >> 1);
          float a = 0.f;
          for (int c = threadIdx.x; c < V; c += blockDim.x)
            if (float_key(s[c]) > mid) a += __expf(s[c] - m.v);
          const float F = block_sum(a, fred) / zsum;
          if (F <= top_p) hi = mid;
          else lo = mid + 1;
        }
        thr = lo;  // smallest key with F <= p
      }
      const int pick = block_draw(s, V, thr, m.v, u01, fred, &s_n);
      if (dbg_probs) {  // test hook: dense probabilities of the filtered distribution
        float a = 0.f;
        for (int c = threadIdx.x; c < V; c += blockDim.x)
          if (float_key(s[c]) >= thr && s[c] != -INFINITY) a += __expf(s[c] - m.v);
        const float kept = block_sum(a, fred);
        float* d = dbg_probs + (long long)b * V;
        for (int c = threadIdx.x; c < V; c += blockDim.x)
          d[c] = (float_key(s[c]) >= thr && s[c] != -INFINITY) ? __expf(s[c] - m.v) / kept : 0.f;
      }
      next = pick;
    } else {
      // ---- top-k (<= 1024): descending selection of the survivors, then nucleus over them ----
      int n = 0;
      float kth = -INFINITY;
      while (n < MAX_SURV) {
        const ArgMax m = block_argmax(s, V, red);
        if (m.v == -INFINITY) break;
        if (n >= kk && m.v < kth) break;  // beyond the k-th value (ties at the k-th are kept)
        if (n == kk - 1) kth = m.v;
        if (threadIdx.x == 0) {
          sv[n] = m.v;
          si[n] = m.i;
          s[m.i] = -INFINITY;
        }
        __syncthreads();
        ++n;
      }
      if (dbg_probs) {
        float* d = dbg_probs + (long long)b * V;
        for (int c = threadIdx.x; c < V; c += blockDim.x) d[c] = 0.f;
        __syncthreads();
      }
      if (threadIdx.x == 0) {
        int keep = n;
        if (top_p > 0.f) {
          // nucleus over the top-k survivors (softmax over survivors only: the rest are -inf)
          float t = 0.f;
          for (int c = 0; c < n; ++c) t += __expf(sv[c] - sv[0]);
          float c2 = 0.f;
          keep = 0;
          for (int c = 0; c < n; ++c) {
            if (c > 0 && c2 > top_p) break;
            c2 += __expf(sv[c] - sv[0]) / t;
            ++keep;
          }
        }
        // multinomial over the kept survivors
        float t = 0.f;
        for (int c = 0; c < keep; ++c) t += __expf(sv[c] - sv[0]);
        const float u = u01 * t;
        float c2 = 0.f;
        int pick = keep > 0 ? si[keep - 1] : 0;
        for (int c = 0; c < keep; ++c) {
          c2 += __expf(sv[c] - sv[0]);
          if (u < c2) {
            pick = si[c];
            break;
          }
        }
        s_n = pick;
        if (dbg_probs) {
          float* d = dbg_probs + (long long)b * V;
          for (int c = 0; c < keep; ++c) d[si[c]] = __expf(sv[c] - sv[0]) / t;
        }
      }
      __syncthreads();
      next = s_n;
    }
  }
  if (threadIdx.x == 0) g[i + 1] = next;
  // the shared step index is advanced by a separate 1-thread kernel (advance_kernel): doing it
  // here would need a grid-wide barrier
}

__global__ void advance_kernel(int* j_ptr) { *j_ptr += 1; }

}  // namespace
}  // namespace mmtg

using namespace mmtg;

extern "C" int64_t mmtg_decode_workspace_bytes(const mmtg_dims* dims, int32_t Lmax) {
  if (!dims || Lmax <= 0) return -1;
  DWs w;
  carve_decode(*dims, Lmax, nullptr, &w);
  return (int64_t)w.bytes;
}

namespace mmtg {
namespace {
__global__ void keymask_init_kernel(const int* __restrict__ prefix_mask, int* __restrict__ keymask,
                                    int B, int Lp, int Lmax) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * Lmax) return;
  const int b = idx / Lmax, p = idx - b * Lmax;
  keymask[idx] = p < Lp ? prefix_mask[b * Lp + p] : 0;
}
}  // namespace

// After a training-style forward over the prefix (prompt + first token; Lp = d.L), move the
// prefix K/V of every layer, the fused context and the prefix key mask into the decode workspace.
int decode_load_prefix(const mmtg_dims& d, const bf16* const* qkv_layers, const float* ctx_out,
                       int Lmax, void* decode_ws, const int* prefix_mask, cudaStream_t st) {
  MMTG_CHECK_ARG(d.L <= Lmax, "prefix longer than the cache");
  DWs w;
  carve_decode(d, Lmax, (uint8_t*)decode_ws, &w);
  const size_t layer_cache = (size_t)d.B * d.NH * Lmax * 64;
  // fused step: the qkv accumulator starts from zero (each position leaves it zeroed again); cache
  // rows past the current position are read as whole 64-key boxes and must hold finite values
  MMTG_CUDA_OK(cudaMemsetAsync(w.mega.qkv_acc, 0, (size_t)d.B * 3 * d.E * 4, st));
  MMTG_CUDA_OK(cudaMemsetAsync(w.kcache, 0, (size_t)d.NL * layer_cache * 2, st));
  MMTG_CUDA_OK(cudaMemsetAsync(w.vcache, 0, (size_t)d.NL * layer_cache * 2, st));
  for (int l = 0; l < d.NL; ++l) {
    dim3 grid(d.B * d.L, d.NH);
    kv_scatter_kernel<<<grid, 32, 0, st>>>(qkv_layers[l], w.kcache + l * layer_cache,
                                           w.vcache + l * layer_cache, d.B, d.L, d.NH, Lmax);
    MMTG_LAUNCH_OK();
  }
  MMTG_CUDA_OK(cudaMemcpyAsync(w.ctx, ctx_out, (size_t)d.S * d.B * d.Dw * 4, cudaMemcpyDeviceToDevice, st));
  keymask_init_kernel<<<cdiv(d.B * Lmax, 256), 256, 0, st>>>(prefix_mask, w.keymask, d.B, d.L, Lmax);
  MMTG_LAUNCH_OK();
  count_launch(d.NL + 1);
  return 0;
}
}  // namespace mmtg

namespace mmtg {
namespace {
int decode_step_impl(const mmtg_model* m, int32_t Lmax, void* decode_ws, const int32_t* gen, int32_t gen_ld,
                     const int32_t* j_ptr, int32_t sent_len, int32_t n_sent, float* logits, void* stream,
                     bool fused) {
  MMTG_CHECK_ARG(m && decode_ws && gen && j_ptr && logits, "null argument");
  const mmtg_dims& d = m->dims;
  DWs w;
  carve_decode(d, Lmax, (uint8_t*)decode_ws, &w);
  cudaStream_t st = (cudaStream_t)stream;
  const float* P = m->params;
  const bf16* W = (const bf16*)m->params_bf16;
  const mmtg_param_offsets& o = m->off;
  const int B = d.B, E = d.E, He = d.He, Dw = d.Dw;
  const float eps = 1e-5f;
  decode_prep_kernel<<<B, 256, 0, st>>>(gen, gen_ld, j_ptr, m->token_table, w.ctx, w.emb16, w.types,
                                        w.posidx, w.keymask, B, d.P, d.S, sent_len, n_sent, Dw, Lmax, w.mega.barrier,
                                        m->table_rows);
  MMTG_LAUNCH_OK();
  count_launch();
  auto gemm = [&](const bf16* A, long long lda, const bf16* Bm, long long ldb, bool b_mn, int N, int K) {
    mmtg_gemm_args a;
    memset(&a, 0, sizeof(a));
    a.A = A; a.lda = lda; a.B = Bm; a.ldb = ldb; a.b_mn_major = b_mn;
    a.M = B; a.N = N; a.K = K; a.block_n = 128;
    return a;
  };
  {
    mmtg_gemm_args a = gemm(w.emb16, Dw, W + o.proj1_w, Dw, false, He, Dw);
    a.out = w.p1; a.ldo = He; a.out_dtype = MMTG_BF16; a.bias = P + o.proj1_b; a.act = MMTG_ACT_TANH;
    MMTG_TRY(decode_gemm(&a, stream));
    a = gemm(w.p1, He, W + o.proj2_w, He, false, E, He);
    a.out = w.h; a.ldo = E; a.out_dtype = MMTG_F32; a.bias = P + o.proj2_b;
    a.rowtab0 = P + o.wpe; a.ldt0 = E; a.rowidx0 = w.posidx;
    a.rowtab1 = P + o.wte; a.ldt1 = E; a.rowidx1 = w.types;
    MMTG_TRY(decode_gemm(&a, stream));
  }
  if (fused) return decode_mega_launch(m, Lmax, w.mega, j_ptr, logits, st);
  const size_t layer_cache = (size_t)B * d.NH * Lmax * 64;
  const int att_smem = (Lmax + 4 + 4 * 64) * 4;
  for (int l = 0; l < d.NL; ++l) {
    const mmtg_layer_offsets& lo = o.layer[l];
    MMTG_TRY(layernorm_fwd(w.h, P + lo.ln1_w, P + lo.ln1_b, w.x16, nullptr, nullptr, nullptr, B, E, eps, st));
    mmtg_gemm_args a = gemm(w.x16, E, W + lo.attn_w, 3 * E, true, 3 * E, E);
    a.out = w.qkv16; a.ldo = 3 * E; a.out_dtype = MMTG_BF16; a.bias = P + lo.attn_b;
    MMTG_TRY(decode_gemm(&a, stream));
    decode_attn_kernel<<<B * d.NH, 128, att_smem, st>>>(
        w.qkv16, w.kcache + l * layer_cache, w.vcache + l * layer_cache, w.keymask, j_ptr, w.att16, B, d.NH,
        d.P, Lmax);
    MMTG_LAUNCH_OK();
    count_launch();
    a = gemm(w.att16, E, W + lo.proj_w, E, true, E, E);
    a.out = w.h2; a.ldo = E; a.out_dtype = MMTG_F32; a.bias = P + lo.proj_b; a.residual = w.h; a.ldr = E;
    MMTG_TRY(decode_gemm(&a, stream));
    MMTG_TRY(layernorm_fwd(w.h2, P + lo.ln2_w, P + lo.ln2_b, w.x16, nullptr, nullptr, nullptr, B, E, eps, st));
    a = gemm(w.x16, E, W + lo.fc_w, 4 * E, true, 4 * E, E);
    a.out = w.a16; a.ldo = 4 * E; a.out_dtype = MMTG_BF16; a.bias = P + lo.fc_b; a.act = MMTG_ACT_GELU_NEW;
    MMTG_TRY(decode_gemm(&a, stream));
    a = gemm(w.a16, 4 * E, W + lo.proj2_w, E, true, E, 4 * E);
    a.out = w.h; a.ldo = E; a.out_dtype = MMTG_F32; a.bias = P + lo.proj2_b; a.residual = w.h2; a.ldr = E;
    MMTG_TRY(decode_gemm(&a, stream));
  }
  MMTG_TRY(layernorm_fwd(w.h, P + o.lnf_w, P + o.lnf_b, w.x16, nullptr, nullptr, nullptr, B, E, eps, st));
  mmtg_gemm_args a = gemm(w.x16, E, W + o.wte, E, false, d.V, E);
  a.out = logits; a.ldo = d.V; a.out_dtype = MMTG_F32;
  MMTG_TRY(decode_gemm(&a, stream));
  return 0;
}
}  // namespace
}  // namespace mmtg

extern "C" int mmtg_decode_step(const mmtg_model* m, int32_t Lmax, void* decode_ws, const int32_t* gen,
                                int32_t gen_ld, const int32_t* j_ptr, int32_t sent_len, int32_t n_sent,
                                float* logits, void* stream) {
  return decode_step_impl(m, Lmax, decode_ws, gen, gen_ld, j_ptr, sent_len, n_sent, logits, stream, false);
}

extern "C" int mmtg_decode_step_fused(const mmtg_model* m, int32_t Lmax, void* decode_ws, const int32_t* gen,
                                      int32_t gen_ld, const int32_t* j_ptr, int32_t sent_len, int32_t n_sent,
                                      float* logits, void* stream) {
  MMTG_CHECK_ARG(m && m->dims.B <= 64 && m->dims.E == 768 && m->dims.NH * 64 == m->dims.E && Lmax <= 1024,
                 "fused decode step supports B <= 64, E = 768, head dim 64, Lmax <= 1024");
  return decode_step_impl(m, Lmax, decode_ws, gen, gen_ld, j_ptr, sent_len, n_sent, logits, stream, true);
}

extern "C" int mmtg_decode_fold_weights(const mmtg_model* m, int32_t Lmax, void* decode_ws, void* stream) {
  MMTG_CHECK_ARG(m && decode_ws && m->params, "null argument");
  DWs w;
  carve_decode(m->dims, Lmax, (uint8_t*)decode_ws, &w);
  return decode_fold_weights(m, w.mega, (cudaStream_t)stream);
}

extern "C" int mmtg_sample_rows(const float* logits, int64_t ld, int32_t* gen, int32_t gen_ld,
                                int32_t* j_ptr, int32_t B, int32_t V, int32_t sent_len,
                                float temperature, int32_t top_k, float top_p, float rep_penalty,
                                uint64_t seed, const uint64_t* seed_dev, int32_t ban_specials,
                                float* dbg_probs, void* stream) {
  MMTG_CHECK_ARG(logits && gen && j_ptr && B > 0 && V > 102 && temperature > 0.f, "bad sampler args");
  MMTG_CHECK_ARG(top_k >= 0 && top_k <= MAX_SURV, "top_k must be in [0, %d] (0 = no top-k filter)", MAX_SURV);
  MMTG_CHECK_ARG(top_p >= 0.f, "top_p must be >= 0");
  MMTG_PER_DEVICE_FLAG(attr_set);
  const int smem = V * 4;
  if (!attr_set) {
    MMTG_CUDA_OK(cudaFuncSetAttribute(sample_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attr_set = true;
  }
  MMTG_CHECK_ARG(smem <= 96 * 1024, "vocabulary too large for the sampler's shared-memory copy");
  cudaStream_t st = (cudaStream_t)stream;
  sample_rows_kernel<<<B, 1024, smem, st>>>(logits, ld, gen, gen_ld, j_ptr, ban_specials, V, sent_len, temperature,
                                            top_k, top_p, rep_penalty, (unsigned long long)seed,
                                            (const unsigned long long*)seed_dev, dbg_probs);
  MMTG_LAUNCH_OK();
  advance_kernel<<<1, 1, 0, st>>>(j_ptr);
  MMTG_LAUNCH_OK();
  count_launch(2);
  return 0;
}
