// Fused sampler of MMTG generation (src/generate.py:64-94,118-142) as a block-level device
// function: shared by the standalone kernel (decode.cu, one 1024-thread block per row) and the
// persistent decode kernel (decode_mega.cu, one 256-thread CTA per row after the lm_head phase).
#pragma once
#include "common.cuh"

namespace mmtg {
namespace {

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

// Shared-memory scratch of one sampling block (besides the V working logits).
struct SamplerScratch {
  ArgMax red[32];
  float fred[32];
  float sv[MAX_SURV];
  int si[MAX_SURV];
  int s_n;
};

// One block samples row `b`: decides g[i + 1] from logits z[0..V) and the history g[0..i]
// (src/generate.py:118-142). Any block size that is a multiple of 32. `s`: V floats of shared memory.
__device__ void sample_row(const float* __restrict__ z, int* __restrict__ g, int b, int i, int ban_specials, int V,
                           int sent_len, float temperature, int top_k, float top_p, float rep_penalty,
                           unsigned long long seed, float* __restrict__ dbg_probs, float* s, SamplerScratch* sc) {
  ArgMax* red = sc->red;
  float* fred = sc->fred;
  float* sv = sc->sv;
  int* si = sc->si;
  int& s_n = sc->s_n;
  int next = -1;
  if (i > 0 && (i + 2) % sent_len == 0) next = 2;        // forced [#EOS#]   (generate.py:118-120)
  else if (i > 0 && (i + 2) % sent_len == 1) next = 1;   // forced [#START#] (generate.py:121-123)
  else if (g[i] == 0 && !dbg_probs) next = 0;            // PAD continuation (generate.py:137-138)
  if (next < 0) {
    for (int c = threadIdx.x; c < V; c += blockDim.x) s[c] = __ldcg(z + c);
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
        float* d = dbg_probs;
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
        float* d = dbg_probs;
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
          float* d = dbg_probs;
          for (int c = 0; c < keep; ++c) d[si[c]] = __expf(sv[c] - sv[0]) / t;
        }
      }
      __syncthreads();
      next = s_n;
    }
  }
  if (threadIdx.x == 0) g[i + 1] = next;
  __syncthreads();
}


}  // namespace
}  // namespace mmtg
