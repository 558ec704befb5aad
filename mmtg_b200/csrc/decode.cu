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
#include "sampler.cuh"

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
  g.h16 = (bf16*)take(64 * E * 2);
  g.h2_16 = (bf16*)take(64 * E * 2);
  g.qkv16 = (bf16*)take(64 * 3 * E * 2);
  g.u16 = (bf16*)take(64 * 4 * E * 2);
  g.stats1 = (float*)take(64 * 32 * 2 * 4);
  g.stats2 = (float*)take(64 * 32 * 2 * 4);
  g.barrier = (unsigned int*)take(256);
  g.f_attn = (bf16*)take((size_t)d.NL * E * 3 * E * 2);
  g.f_fc = (bf16*)take((size_t)d.NL * E * 4 * E * 2);
  g.f_wte = (bf16*)take((size_t)d.V * E * 2);
  g.f_vec = (float*)take(((size_t)d.NL * 14 * E + 2 * (size_t)d.V) * 4);
  g.T1 = (float*)take((size_t)d.V * d.He * 4);
  g.C1 = (float*)take((size_t)d.S * 64 * d.He * 4);
  g.w2t = (bf16*)take((size_t)d.He * E * 2);
  g.table16 = (bf16*)take((size_t)d.V * d.Dw * 2);
  g.ctx16 = (bf16*)take((size_t)d.S * 64 * d.Dw * 2);
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
  if (b == 0 && threadIdx.x == 0) {  // the megakernel's grid barrier starts from zero
    barrier[0] = 0u;
    barrier[1] = 0u;
  }
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

// Standalone sampler launch: one 1024-thread block per row (sampler.cuh holds the algorithm).
__global__ void __launch_bounds__(1024)
sample_rows_kernel(const float* __restrict__ logits, long long ld, int* __restrict__ gen, int gen_ld,
                   int* __restrict__ j_ptr, int ban_specials, int V, int sent_len, float temperature,
                   int top_k, float top_p, float rep_penalty, unsigned long long seed,
                   const unsigned long long* __restrict__ seed_dev, float* __restrict__ dbg_probs) {
  if (seed_dev) seed = seed_dev[0];  // device-side seed: the launch stays CUDA-graph replayable
  extern __shared__ float s[];  // [V] working logits
  __shared__ SamplerScratch sc;
  const int b = blockIdx.x;
  sample_row(logits + (long long)b * ld, gen + (long long)b * gen_ld, b, *j_ptr, ban_specials, V, sent_len, temperature,
             top_k, top_p, rep_penalty, seed, dbg_probs ? dbg_probs + (long long)b * V : nullptr, s, &sc);
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
  // fused step: cache rows past the current position are read as whole 64-key boxes and must hold
  // finite values; its grid-barrier counter starts from zero
  MMTG_CUDA_OK(cudaMemsetAsync(w.mega.barrier, 0, 256, st));
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
  if (fused) return decode_mega_launch(m, Lmax, w.mega, const_cast<int*>(j_ptr), logits, nullptr, st);
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

namespace mmtg {
namespace {
// out[k][n] = bf16(in[n][k])  (nn.Linear [N, K] fp32 -> [K, N] bf16), 32x32 tiles through smem
__global__ void transpose_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, int N, int K) {
  __shared__ float t[32][33];
  const int n0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int n = n0 + i, k = k0 + threadIdx.x;
    t[i][threadIdx.x] = (n < N && k < K) ? in[(long long)n * K + k] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int k = k0 + i, n = n0 + threadIdx.x;
    if (k < K && n < N) out[(long long)k * N + n] = __float2bfloat16(t[threadIdx.x][i]);
  }
}
// Full-step mode tables (once per generation call, after mmtg_decode_load_prefix): projector
// layer 1 (src/model.py:316) is linear in table[tok] + ctx, so table W1^T and ctx W1^T are
// computed once with the tcgen05 GEMM and each position gathers two rows instead of streaming W1.
int fold_projector(const mmtg_model* m, const DWs& w, void* stream) {
  const mmtg_dims& d = m->dims;
  cudaStream_t st = (cudaStream_t)stream;
  const float* P = m->params;
  const bf16* W = (const bf16*)m->params_bf16;
  const int rows = m->table_rows < d.V ? m->table_rows : d.V;
  transpose_bf16_kernel<<<dim3(cdiv(d.E, 32), cdiv(d.He, 32)), dim3(32, 8), 0, st>>>(P + m->off.proj2_w, w.mega.w2t, d.E, d.He);
  MMTG_LAUNCH_OK();
  count_launch();
  MMTG_TRY(cast_bf16(m->token_table, w.mega.table16, (long long)rows * d.Dw, st));
  MMTG_TRY(cast_bf16(w.ctx, w.mega.ctx16, (long long)d.S * d.B * d.Dw, st));
  mmtg_gemm_args a;
  memset(&a, 0, sizeof(a));
  a.A = w.mega.table16; a.lda = d.Dw; a.B = W + m->off.proj1_w; a.ldb = d.Dw;
  a.M = rows; a.N = d.He; a.K = d.Dw; a.block_n = 128;
  a.out = w.mega.T1; a.ldo = d.He; a.out_dtype = MMTG_F32;
  MMTG_TRY(mmtg_gemm_bf16(&a, stream));
  a.A = w.mega.ctx16; a.M = d.S * d.B; a.out = w.mega.C1;
  MMTG_TRY(decode_gemm(&a, stream));
  return 0;
}
}  // namespace
}  // namespace mmtg

extern "C" int mmtg_decode_fold_weights(const mmtg_model* m, int32_t Lmax, void* decode_ws, void* stream) {
  MMTG_CHECK_ARG(m && decode_ws && m->params, "null argument");
  DWs w;
  carve_decode(m->dims, Lmax, (uint8_t*)decode_ws, &w);
  MMTG_TRY(decode_fold_weights(m, w.mega, (cudaStream_t)stream));
  return fold_projector(m, w, stream);
}

extern "C" int mmtg_decode_steps_fused(const mmtg_model* m, int32_t Lmax, void* decode_ws, int32_t* gen, int32_t gen_ld,
                                       int32_t* j_ptr, int32_t sent_len, int32_t n_sent, int32_t n_steps,
                                       float temperature, int32_t top_k, float top_p, float rep_penalty,
                                       const uint64_t* seed_dev, float* logits, void* stream) {
  MMTG_CHECK_ARG(m && decode_ws && gen && j_ptr && logits && n_steps >= 1, "bad decode_steps_fused arguments");
  MMTG_CHECK_ARG(m->dims.B <= 64 && m->dims.E == 768 && m->dims.NH * 64 == m->dims.E && Lmax <= 1024,
                 "fused decode steps support B <= 64, E = 768, head dim 64, Lmax <= 1024");
  DWs w;
  carve_decode(m->dims, Lmax, (uint8_t*)decode_ws, &w);
  MegaStepArgs fa;
  fa.n_steps = n_steps; fa.gen = gen; fa.gen_ld = gen_ld; fa.sent_len = sent_len; fa.n_sent = n_sent;
  fa.temperature = temperature; fa.top_k = top_k; fa.top_p = top_p; fa.rep_penalty = rep_penalty;
  fa.seed_dev = (const unsigned long long*)seed_dev;
  return decode_mega_launch(m, Lmax, w.mega, j_ptr, logits, &fa, (cudaStream_t)stream);
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
