// Native orchestration of the MMTG training hot path on ONE stream:
//   forward  = MMTG.forward, training branch (src/model.py:356-400): encoder -> LayerNorms ->
//              alpha attention x2 (+KL) -> beta gate -> embedding build -> projector -> 12 GPT-2
//              blocks -> ln_f -> tied lm_head (+ fused row log-sum-exp) -> HF causal-LM loss;
//   backward = the hand-derived reverse of all of the above, split in stages so the host can
//              overlap a bucketed gradient all-reduce with the remaining backward work.
// No tensor library, no allocation: the caller provides one workspace, carved here.
//
// Operand layouts (why no transposes exist anywhere):
//   HF Conv1D weight W[in,out]:  forward  y = x W      -> B operand MN-major (native storage)
//                                dgrad   dx = dy W^T   -> B operand K-major  (native storage)
//                                wgrad   dW = x^T dy   -> A, B operands MN-major
//   nn.Linear weight W[out,in]:  forward K-major, dgrad MN-major, wgrad dW = dy^T x (MN, MN)
#include <string.h>

#include "../../include/mmtg_b200.h"
#include "ops.h"

namespace mmtg {

namespace {

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct LayerWs {
  float* h_in;  // alias of the previous layer's output (residual stream), fp32 [M,E]
  bf16* x1;
  float *mean1, *rstd1;
  bf16* qkv;
  bf16* att;
  float* lse;
  float* h_mid;
  bf16* x2;
  float *mean2, *rstd2;
  bf16 *u, *a;  // u = gelu_new'(pre-activation) (saved derivative), a = gelu_new(pre-activation)
  float* h_out;
};

struct Ws {
  // decoder forward
  bf16* emb16;
  bf16* p1;
  float* h0;
  LayerWs layer[MMTG_MAX_LAYERS];
  bf16* xf;
  float *meanf, *rstdf;
  float* lse_part;
  float* lse;
  float* hf_sum;
  // decoder backward
  bf16* dlogits16;
  float* dh;
  bf16 *g16x[2], *g16b[2], *du[2], *dx[2], *dx2[2], *datt[2], *dqkv[2], *dp1, *dE;  // [2]: ping-pong by backward stage parity
  float* delta;
  // encoder forward
  bf16 *x_topic16, *x_mod16[2];
  float* topic_pre;
  float* gi_all[2];
  float* gh;
  float* gru_save[2];
  float* hout[2];
  bf16* hout16[2];
  float *topic_ln, *topic_mean, *topic_rstd;
  bf16* ln16[2];
  float *ln_mean[2], *ln_rstd[2];
  float* aqkv[2];
  float* actx[2];
  float* aprobs[2];
  float* klpart;
  bf16* o16;
  float* att3;
  float* ctx_out;
  // encoder backward
  bf16 *dctx16, *do16;
  float *dtopic_ln, *dactx[2];
  bf16* daqkv16;
  float* dln;
  float* dhout;
  bf16 *dgi16, *dgh16;
  float *dhz, *dcarry;
  float* dtopic_pre;
  bf16* dtopic_pre16;
  size_t bytes;
};

// Carve the workspace. Called with base == nullptr to size it.
void carve(const mmtg_dims& d, uint8_t* base, Ws* w) {
  size_t off = 0;
  auto take = [&](size_t bytes) -> uint8_t* {
    uint8_t* p = base ? base + off : nullptr;
    off += align_up(bytes, 256);
    return p;
  };
  const size_t M = (size_t)d.B * d.L, E = d.E, SB = (size_t)d.S * d.B, He = d.He, Dw = d.Dw;
  auto f32 = [&](size_t n) { return (float*)take(n * 4); };
  auto b16 = [&](size_t n) { return (bf16*)take(n * 2); };
  w->emb16 = b16(M * Dw);
  w->p1 = b16(M * He);
  w->h0 = f32(M * E);
  float* prev = w->h0;
  for (int l = 0; l < d.NL; ++l) {
    LayerWs& L = w->layer[l];
    L.h_in = prev;
    L.x1 = b16(M * E);
    L.mean1 = f32(M);
    L.rstd1 = f32(M);
    L.qkv = b16(M * 3 * E);
    L.att = b16(M * E);
    L.lse = f32((size_t)d.B * d.NH * d.L);
    L.h_mid = f32(M * E);
    L.x2 = b16(M * E);
    L.mean2 = f32(M);
    L.rstd2 = f32(M);
    L.u = b16(M * 4 * E);
    L.a = b16(M * 4 * E);
    L.h_out = f32(M * E);
    prev = L.h_out;
  }
  w->xf = b16(M * E);
  w->meanf = f32(M);
  w->rstdf = f32(M);
  const size_t nt = (size_t)2 * cdiv(d.V, 128);  // half-tile partial slots, either tile width
  w->lse_part = f32(nt * M * 2);
  w->lse = f32(M);
  w->hf_sum = f32(d.B);
  w->dlogits16 = b16(M * d.Vp);
  w->dh = f32(M * E);
  w->g16x[0] = b16(M * E);
  w->g16x[1] = b16(M * E);
  // every per-stage temporary exists twice: the side-stream work of stage s (weight / parameter
  // gradients reading these buffers) may still run while the chain of stage s + 1 writes the others
  for (int k = 0; k < 2; ++k) {
    w->g16b[k] = b16(M * E);
    w->du[k] = b16(M * 4 * E);
    w->dx[k] = b16(M * E);
    w->dx2[k] = b16(M * E);
    w->datt[k] = b16(M * E);
    w->dqkv[k] = b16(M * 3 * E);
  }
  w->dp1 = b16(M * He);
  w->dE = b16(M * Dw);
  w->delta = f32((size_t)d.B * d.NH * d.L);
  // encoder
  w->x_topic16 = b16((size_t)d.B * Dw);
  for (int m = 0; m < 2; ++m) w->x_mod16[m] = b16(SB * Dw);
  w->topic_pre = f32((size_t)d.B * He);
  for (int m = 0; m < 2; ++m) w->gi_all[m] = f32(SB * 3 * He);
  w->gh = f32((size_t)d.B * 3 * He);
  for (int m = 0; m < 2; ++m) w->gru_save[m] = f32(SB * 4 * He);
  for (int m = 0; m < 2; ++m) w->hout[m] = f32(SB * He);
  for (int m = 0; m < 2; ++m) w->hout16[m] = b16(SB * He);
  w->topic_ln = f32((size_t)d.B * He);
  w->topic_mean = f32(d.B);
  w->topic_rstd = f32(d.B);
  for (int m = 0; m < 2; ++m) {
    w->ln16[m] = b16(SB * He);
    w->ln_mean[m] = f32(SB);
    w->ln_rstd[m] = f32(SB);
    w->aqkv[m] = f32(SB * 3 * He);
    w->actx[m] = f32(SB * He);
    w->aprobs[m] = f32((size_t)d.B * d.alpha_heads * d.S * d.S);
  }
  w->klpart = f32((size_t)2 * d.B * d.alpha_heads);
  w->o16 = b16(SB * He);
  w->att3 = f32(SB * 3);
  w->ctx_out = f32(SB * Dw);
  w->dctx16 = b16(SB * Dw);
  w->do16 = b16(SB * He);
  w->dtopic_ln = f32((size_t)d.B * He);
  for (int m = 0; m < 2; ++m) w->dactx[m] = f32(SB * He);
  w->daqkv16 = b16(SB * 3 * He);
  w->dln = f32(SB * He);
  w->dhout = f32(SB * He);
  w->dgi16 = b16(SB * 3 * He);
  w->dgh16 = b16(SB * 3 * He);
  w->dhz = f32((size_t)d.B * He);
  w->dcarry = f32((size_t)d.B * He);
  w->dtopic_pre = f32((size_t)d.B * He);
  w->dtopic_pre16 = b16((size_t)d.B * He);
  w->bytes = off;
}

int check_dims(const mmtg_dims& d) {
  MMTG_CHECK_ARG(d.B > 0 && d.L == d.P + d.T && d.T >= 1, "bad dims B=%d P=%d T=%d L=%d", d.B, d.P, d.T, d.L);
  MMTG_CHECK_ARG(d.NL > 0 && d.NL <= MMTG_MAX_LAYERS, "n_layer %d out of range", d.NL);
  MMTG_CHECK_ARG(d.E == d.NH * 64, "head_dim must be 64 (E=%d NH=%d)", d.E, d.NH);
  MMTG_CHECK_ARG(d.Vp >= d.V && d.Vp % 8 == 0, "Vp must be V rounded up to a multiple of 8");
  MMTG_CHECK_ARG(d.L <= d.n_pos, "sequence %d exceeds n_positions %d", d.L, d.n_pos);
  MMTG_CHECK_ARG(d.He % d.alpha_heads == 0, "bad alpha heads");
  return 0;
}

// Small builder around mmtg_gemm_args.
struct Gemm {
  mmtg_gemm_args a;
  Gemm(const void* A, long long lda, bool a_mn, const void* B, long long ldb, bool b_mn, int M, int N,
       int K) {
    memset(&a, 0, sizeof(a));
    a.A = A; a.lda = lda; a.a_mn_major = a_mn;
    a.B = B; a.ldb = ldb; a.b_mn_major = b_mn;
    a.M = M; a.N = N; a.K = K;
  }
  Gemm& out_f32(float* o, long long ld) { a.out = o; a.ldo = ld; a.out_dtype = MMTG_F32; return *this; }
  Gemm& out_bf16(bf16* o, long long ld) { a.out = o; a.ldo = ld; a.out_dtype = MMTG_BF16; return *this; }
  Gemm& bias(const float* b) { a.bias = b; return *this; }
  Gemm& act(int x) { a.act = x; return *this; }
  Gemm& out2(bf16* o, long long ld) { a.out2 = o; a.ldo2 = ld; return *this; }
  Gemm& residual(const float* r, long long ld) { a.residual = r; a.ldr = ld; return *this; }
  Gemm& dgelu(const bf16* u, long long ld) { a.dgelu_src = u; a.ldg = ld; return *this; }
  Gemm& dtanh(const bf16* y, long long ld) { a.dgelu_src = y; a.ldg = ld; a.dact_tanh_out = 1; return *this; }
  Gemm& dmul(const bf16* d, long long ld) { a.dgelu_src = d; a.ldg = ld; a.dact_tanh_out = 2; return *this; }
  Gemm& out2_deriv(bf16* o, long long ld) { a.out2 = o; a.ldo2 = ld; a.out2_mode = 1; return *this; }
  Gemm& colsum(float* c) { a.colsum = c; return *this; }
  Gemm& accumulate(int split = 1) { a.accumulate = 1; a.split_k = split; return *this; }
  Gemm& lse(float* p, int bn) { a.lse_partial = p; a.block_n = bn; return *this; }
  Gemm& dropout(const DropSpec& d) {
    if (d.seed && d.p > 0.f) { a.drop_seed = (const uint64_t*)d.seed; a.drop_site = d.site; a.drop_p = d.p; }
    return *this;
  }
  int run(cudaStream_t st) { return mmtg_gemm_bf16(&a, (void*)st); }
};

// dropout sites of the decoder (include/mmtg_b200.h): block l -> 4l + {0 attn probs, 1 attn
// c_proj, 2 mlp c_proj}; MMTG_DROP_SITE_EMBD for the embedding sum
inline DropSpec drop_spec(const mmtg_model* m, uint32_t site, float p, int mask_dx32 = 0) {
  DropSpec d;
  d.seed = (p > 0.f) ? (const unsigned long long*)m->drop_seed : nullptr;
  d.site = site;
  d.p = d.seed ? p : 0.f;
  d.mask_dx32 = mask_dx32;
  return d;
}

// wgrad: dW[rows, cols] += X^T Y with X [K, rows] and Y [K, cols] row-major activations.
// Output tiles are few (dW is at most 768 x 3072), so the reduction dimension (the tokens) is
// split: pick the tile width and split count whose (tile pair x split) work units fill whole
// waves of the 74 CTA-pair clusters, preferring fewer splits (fewer fp32 atomics) on ties.
int wgrad(const bf16* X, long long ldx, const bf16* Y, long long ldy, float* dW, long long ldw,
          int rows, int cols, int K, cudaStream_t st, int grid_mode = 0) {
  const int clusters = num_sms() / 2;
  const int pairs = (cdiv(rows, 128) + 1) / 2;
  const int nkb = cdiv(K, 64);
  int best_bn = 128, best_split = 1;
  double best_cost = 1e30;
  for (int bn = 128; bn <= 256; bn += 128) {
    if (bn == 256 && cols < 256) break;
    const int units = pairs * cdiv(cols, bn);
    const int max_split = nkb / 8 > 0 ? nkb / 8 : 1;  // keep >= 8 k-blocks per unit
    for (int sp = 1; sp <= max_split && sp <= 16; ++sp) {
      const int waves = cdiv(units * sp, clusters);
      // time ~ waves x (k-blocks per unit x per-k-block cost(bn) + fixed tile cost)
      const double kcost = bn == 256 ? 1.0 : 0.62;  // 128-wide tiles move more bytes per FLOP
      const double cost = waves * (cdiv(nkb, sp) * kcost + 6.0) * (1.0 + 0.01 * sp);
      if (cost < best_cost) {
        best_cost = cost;
        best_bn = bn;
        best_split = sp;
      }
    }
  }
  Gemm g(X, ldx, true, Y, ldy, true, rows, cols, K);
  g.out_f32(dW, ldw).accumulate(best_split);
  g.a.block_n = best_bn;
  g.a.grid_mode = grid_mode;
  return g.run(st);
}

}  // namespace

}  // namespace mmtg

using namespace mmtg;

extern "C" int64_t mmtg_train_workspace_bytes(const mmtg_dims* dims) {
  if (!dims || check_dims(*dims) != 0) return -1;
  Ws w;
  carve(*dims, nullptr, &w);
  return (int64_t)w.bytes;
}
extern "C" float* mmtg_ws_lse(const mmtg_dims* dims, void* workspace) {
  Ws w;
  carve(*dims, (uint8_t*)workspace, &w);
  return w.lse;
}
extern "C" void* mmtg_ws_dlogits_bf16(const mmtg_dims* dims, void* workspace) {
  Ws w;
  carve(*dims, (uint8_t*)workspace, &w);
  return w.dlogits16;
}
extern "C" int mmtg_dlogits_from_f32(const mmtg_dims* dims, void* workspace, const float* src, void* stream) {
  MMTG_CHECK_ARG(dims && workspace && src, "null argument");
  Ws w;
  carve(*dims, (uint8_t*)workspace, &w);
  return dlogits_f32_to_bf16(src, w.dlogits16, dims->B * dims->L, dims->V, dims->Vp, (cudaStream_t)stream);
}

namespace mmtg {
int decode_load_prefix(const mmtg_dims& d, const bf16* const* qkv_layers, const float* ctx_out,
                       int Lmax, void* decode_ws, const int* prefix_mask, cudaStream_t st);
}
extern "C" int mmtg_decode_load_prefix(const mmtg_dims* dp, void* train_ws, int32_t Lmax,
                                       void* decode_ws, const int32_t* prefix_mask, void* stream) {
  MMTG_CHECK_ARG(dp && train_ws && decode_ws && prefix_mask, "null argument");
  MMTG_TRY(check_dims(*dp));
  Ws w;
  carve(*dp, (uint8_t*)train_ws, &w);
  const bf16* qkv[MMTG_MAX_LAYERS];
  for (int l = 0; l < dp->NL; ++l) qkv[l] = w.layer[l].qkv;
  return decode_load_prefix(*dp, qkv, w.ctx_out, Lmax, decode_ws, prefix_mask, (cudaStream_t)stream);
}

extern "C" int mmtg_train_forward(const mmtg_model* m, const mmtg_batch* b, void* workspace,
                                  int64_t workspace_bytes, float* logits, float* scalars_out,
                                  int32_t save_for_backward, void* stream) {
  (void)save_for_backward;
  MMTG_CHECK_ARG(m && b && workspace && logits && scalars_out, "null argument");
  const mmtg_dims& d = m->dims;
  MMTG_TRY(check_dims(d));
  Ws w;
  carve(d, (uint8_t*)workspace, &w);
  MMTG_CHECK_ARG((int64_t)w.bytes <= workspace_bytes, "workspace too small: need %zu, have %lld", w.bytes,
                 (long long)workspace_bytes);
  MMTG_CHECK_ARG(((uintptr_t)workspace & 255) == 0, "workspace must be 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const float* P = m->params;
  const bf16* W = (const bf16*)m->params_bf16;
  const mmtg_param_offsets& o = m->off;
  const int B = d.B, S = d.S, He = d.He, Dw = d.Dw, E = d.E, SB = d.S * d.B, M = d.B * d.L;
  const float eps = 1e-5f;

  // ---------------- encoder (src/model.py:63-81) ----------------
  MMTG_TRY(pack_sb(b->topic_emb, w.x_topic16, B, 1, Dw, st));
  MMTG_TRY(pack_sb(b->img_embs, w.x_mod16[0], B, S, Dw, st));
  MMTG_TRY(pack_sb(b->txt_embs, w.x_mod16[1], B, S, Dw, st));
  MMTG_TRY(Gemm(w.x_topic16, Dw, false, W + o.topic_w, Dw, false, B, He, Dw)
               .out_f32(w.topic_pre, He).bias(P + o.topic_b).run(st));
  for (int md = 0; md < 2; ++md) {
    MMTG_TRY(Gemm(w.x_mod16[md], Dw, false, W + o.gru_w_ih[md], Dw, false, SB, 3 * He, Dw)
                 .out_f32(w.gi_all[md], 3 * He).bias(P + o.gru_b_ih[md]).run(st));
    for (int t = 0; t < S; ++t) {
      const float* hp = t > 0 ? w.hout[md] + (size_t)(t - 1) * B * He : nullptr;
      if (t > 0)
        MMTG_TRY(Gemm(w.hout16[md] + (size_t)(t - 1) * B * He, He, false, W + o.gru_w_hh[md], He, false, B,
                      3 * He, He)
                     .out_f32(w.gh, 3 * He).bias(P + o.gru_b_hh[md]).run(st));
      MMTG_TRY(gru_gate_fwd(w.gi_all[md] + (size_t)t * B * 3 * He, t > 0 ? w.gh : nullptr,
                            P + o.gru_b_hh[md], hp, w.hout[md] + (size_t)t * B * He,
                            w.hout16[md] + (size_t)t * B * He, w.gru_save[md] + (size_t)t * B * 4 * He, B,
                            He, st));
    }
  }
  // ---------------- LayerNorms (src/model.py:380-382) ----------------
  MMTG_TRY(layernorm_fwd(w.topic_pre, P + o.enc_ln_w[0], P + o.enc_ln_b[0], nullptr, w.topic_ln,
                         w.topic_mean, w.topic_rstd, B, He, eps, st));
  for (int md = 0; md < 2; ++md)
    MMTG_TRY(layernorm_fwd(w.hout[md], P + o.enc_ln_w[1 + md], P + o.enc_ln_b[1 + md], w.ln16[md],
                           nullptr, w.ln_mean[md], w.ln_rstd[md], SB, He, eps, st));
  // ---------------- alpha attention + KL (src/model.py:133-161) ----------------
  for (int md = 0; md < 2; ++md) {
    MMTG_TRY(Gemm(w.ln16[md], He, false, W + o.alpha_qkv_w[md], He, false, SB, 3 * He, He)
                 .out_f32(w.aqkv[md], 3 * He).bias(P + o.alpha_qkv_b[md]).run(st));
    MMTG_TRY(alpha_fwd(w.aqkv[md], w.actx[md], w.aprobs[md], w.klpart + (size_t)md * B * d.alpha_heads,
                       B, d.alpha_heads, S, He / d.alpha_heads, st));
  }
  MMTG_TRY(sum_scale(w.klpart, scalars_out + 1, 2 * B * d.alpha_heads, 1.f / (float)(S * B), st));
  // ---------------- beta gate + out_linear (src/model.py:181-202) ----------------
  MMTG_TRY(beta_fwd(w.topic_ln, w.actx[0], w.actx[1], P + o.beta_att_w, P + o.beta_att_b, w.o16,
                    w.att3, B, S, He, st));
  MMTG_TRY(Gemm(w.o16, He, false, W + o.beta_out_w, He, false, SB, Dw, He)
               .out_f32(w.ctx_out, Dw).bias(P + o.beta_out_b).run(st));
  // ---------------- decoder embedding + projector (src/model.py:253-281) ----------------
  MMTG_TRY(embed_fwd(m->token_table, b->topic_ids, b->targets, w.ctx_out, w.emb16, B, d.P, d.T, S,
                     d.two_sent, Dw, m->table_rows, st));
  MMTG_TRY(Gemm(w.emb16, Dw, false, W + o.proj1_w, Dw, false, M, He, Dw)
               .out_bf16(w.p1, He).bias(P + o.proj1_b).act(MMTG_ACT_TANH).run(st));
  {  // h0 = p1 W2^T + b2 + wpe[pos] + wte[type]   (HF modeling_gpt2.py:579-612)
    Gemm g(w.p1, He, false, W + o.proj2_w, He, false, M, E, He);
    g.out_f32(w.h0, E).bias(P + o.proj2_b);
    g.a.rowtab0 = P + o.wpe; g.a.ldt0 = E; g.a.rowmod0 = d.L;
    g.a.rowtab1 = P + o.wte; g.a.ldt1 = E; g.a.rowidx1 = b->type_ids;
    g.dropout(drop_spec(m, MMTG_DROP_SITE_EMBD, m->p_embd));  // embd_pdrop (HF modeling_gpt2.py:612)
    MMTG_TRY(g.run(st));
  }
  // ---------------- GPT-2 blocks (HF modeling_gpt2.py:262-309) ----------------
  for (int l = 0; l < d.NL; ++l) {
    const LayerWs& L = w.layer[l];
    const mmtg_layer_offsets& lo = o.layer[l];
    MMTG_TRY(layernorm_fwd(L.h_in, P + lo.ln1_w, P + lo.ln1_b, L.x1, nullptr, L.mean1, L.rstd1, M, E, eps, st));
    MMTG_TRY(Gemm(L.x1, E, false, W + lo.attn_w, 3 * E, true, M, 3 * E, E)
                 .out_bf16(L.qkv, 3 * E).bias(P + lo.attn_b).run(st));
    const DropSpec d_att = drop_spec(m, 4u * l, m->p_attn), d_r1 = drop_spec(m, 4u * l + 1, m->p_resid),
                   d_r2 = drop_spec(m, 4u * l + 2, m->p_resid);
    MMTG_TRY(attn_fwd(L.qkv, b->attn_mask, L.att, L.lse, B, d.L, d.NH, st, &d_att));
    MMTG_TRY(Gemm(L.att, E, false, W + lo.proj_w, E, true, M, E, E)
                 .out_f32(L.h_mid, E).bias(P + lo.proj_b).residual(L.h_in, E).dropout(d_r1).run(st));
    MMTG_TRY(layernorm_fwd(L.h_mid, P + lo.ln2_w, P + lo.ln2_b, L.x2, nullptr, L.mean2, L.rstd2, M, E, eps, st));
    MMTG_TRY(Gemm(L.x2, E, false, W + lo.fc_w, 4 * E, true, M, 4 * E, E)
                 .out_bf16(L.a, 4 * E).out2_deriv(L.u, 4 * E).bias(P + lo.fc_b).act(MMTG_ACT_GELU_NEW).run(st));
    MMTG_TRY(Gemm(L.a, 4 * E, false, W + lo.proj2_w, E, true, M, E, 4 * E)
                 .out_f32(L.h_out, E).bias(P + lo.proj2_b).residual(L.h_mid, E).dropout(d_r2).run(st));
  }
  // ---------------- ln_f + tied lm_head + HF loss ----------------
  const float* h_last = w.layer[d.NL - 1].h_out;
  MMTG_TRY(layernorm_fwd(h_last, P + o.lnf_w, P + o.lnf_b, w.xf, nullptr, w.meanf, w.rstdf, M, E, eps, st));
  const int bn = 256;
  MMTG_TRY(Gemm(w.xf, E, false, W + o.wte, E, false, M, d.V, E).out_f32(logits, d.V).lse(w.lse_part, bn).run(st));
  MMTG_TRY(lse_combine(w.lse_part, w.lse, M, 2 * cdiv(d.V, bn), st));
  MMTG_TRY(ce_reduce(logits, d.V, w.lse, b->topic_ids, b->targets, w.hf_sum, nullptr, scalars_out, B,
                     d.L, d.P, d.T, st));
  return 0;
}

namespace mmtg {
namespace {
// Weight-gradient GEMMs are off the backward's critical chain (only the optimizer consumes them):
// they run on a side stream, forked after their operands exist and joined at the end of the
// stage, so their CTAs fill the idle SMs of the chain's partial waves (the N = 768 dgrads run
// 90 tile pairs on 74 clusters; the whole-head attention backward 384 CTAs on 148 SMs) instead
// of queueing behind them. MMTG_WGRAD_STREAM=0 keeps everything on one stream.
struct SideStream {
  cudaStream_t side = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  cudaEvent_t done[2] = {nullptr, nullptr};  // side work of the backward stages of either parity
  bool used = false;
};
int g_side_enabled = -1;  // -1: from MMTG_WGRAD_STREAM (default on); 0 / 1: mmtg_set_wgrad_side_stream
SideStream* side_stream() {
  if (g_side_enabled < 0) {
    const char* e = getenv("MMTG_WGRAD_STREAM");
    g_side_enabled = (e && e[0] == '0') ? 0 : 1;
  }
  if (!g_side_enabled) return nullptr;
  static SideStream per_dev[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  SideStream& s = per_dev[dev];
  if (!s.side) {
    int lo = 0, hi = 0;  // lowest priority: the chain's kernels win SMs as they free up
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (cudaStreamCreateWithPriority(&s.side, cudaStreamNonBlocking, lo) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s.done[0], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s.done[1], cudaEventDisableTiming) != cudaSuccess)
      return nullptr;
  }
  return &s;
}
}  // namespace
}  // namespace mmtg

// Run the weight-gradient GEMMs of mmtg_train_backward on the low-priority side stream (1,
// default) or inline on the caller's stream (0: per-launch timings are then not perturbed by
// concurrent kernels — bench.py's profiled steps).
extern "C" int mmtg_set_wgrad_side_stream(int32_t enable) {
  mmtg::g_side_enabled = enable ? 1 : 0;
  return 0;
}

extern "C" int mmtg_train_backward(const mmtg_model* m, const mmtg_batch* b, void* workspace,
                                   int64_t workspace_bytes, const float* g_kl, int32_t stage_begin,
                                   int32_t stage_end, void* stream) {
  MMTG_CHECK_ARG(m && b && workspace && m->grads, "null argument");
  const mmtg_dims& d = m->dims;
  MMTG_TRY(check_dims(d));
  Ws w;
  carve(d, (uint8_t*)workspace, &w);
  MMTG_CHECK_ARG((int64_t)w.bytes <= workspace_bytes, "workspace too small");
  MMTG_CHECK_ARG(stage_begin >= 0 && stage_end <= d.NL + 3 && stage_begin <= stage_end, "bad stage range");
  cudaStream_t st = (cudaStream_t)stream;
  const float* P = m->params;
  const bf16* W = (const bf16*)m->params_bf16;
  float* G = m->grads;
  const mmtg_param_offsets& o = m->off;
  const int B = d.B, S = d.S, He = d.He, Dw = d.Dw, E = d.E, SB = d.S * d.B, M = d.B * d.L;
  SideStream* ss = side_stream();
  static const int side_grid_mode = [] {
    const char* e = getenv("MMTG_WGRAD_GRID");
    return e ? atoi(e) : 0;
  }();
  const int wg = ss ? side_grid_mode : 0;
  // run `fn(stream)` on the side stream, ordered after everything issued on `st` so far
  auto on_side = [&](auto&& fn) -> int {
    if (!ss) return fn(st);
    MMTG_CUDA_OK(cudaEventRecord(ss->fork, st));
    MMTG_CUDA_OK(cudaStreamWaitEvent(ss->side, ss->fork, 0));
    ss->used = true;
    return fn(ss->side);
  };
  // The side-stream work of stage s is joined ONE STAGE LATE: all per-stage temporaries are
  // double-buffered by stage parity, so the only buffer stage s + 1 writes that the side work of
  // stage s still reads is g16x[s & 1] (its g_in), rewritten by the LAST kernel of stage s + 1.
  // Joining at every stage end put the tail of the side work (the kernels forked last) on the
  // critical chain.
  bool pending[2] = {false, false};
  auto stage_done = [&](int stage) -> int {  // end of a stage: mark where its side work ends
    if (ss && ss->used) {
      MMTG_CUDA_OK(cudaEventRecord(ss->done[stage & 1], ss->side));
      pending[stage & 1] = true;
      ss->used = false;
    }
    return 0;
  };
  auto wait_side_of = [&](int stage) -> int {  // the chain waits for the side work of `stage`
    if (ss && stage >= 0 && pending[stage & 1]) {
      MMTG_CUDA_OK(cudaStreamWaitEvent(st, ss->done[stage & 1], 0));
      pending[stage & 1] = false;
    }
    return 0;
  };

  for (int stage = stage_begin; stage < stage_end; ++stage) {
    bf16* g_in = w.g16x[stage & 1];         // bf16 gradient of this stage's input (from the stage before)
    bf16* g_out = w.g16x[(stage + 1) & 1];  // ... emitted for the next stage
    const int sp = stage & 1;
    bf16 *t_du = w.du[sp], *t_dx = w.dx[sp], *t_dx2 = w.dx2[sp], *t_g16b = w.g16b[sp], *t_datt = w.datt[sp],
         *t_dqkv = w.dqkv[sp];
    if (stage > d.NL) MMTG_TRY(wait_side_of(stage - 1));  // (embedding / encoder stages: simply at their start)
    if (stage == 0) {
      // ---------------- lm_head (tied wte) + ln_f ----------------
      // dxf = dlogits · wte  (B operand: wte [V,E] as MN-major [N=E, K=V])
      MMTG_TRY(Gemm(w.dlogits16, d.Vp, false, W + o.wte, E, true, M, E, d.V).out_bf16(t_dx, E).run(st));
      // dwte += dlogits^T · xf
      MMTG_TRY(on_side([&](cudaStream_t s2) { return wgrad(w.dlogits16, d.Vp, w.xf, E, G + o.wte, E, d.V, E, M, s2, wg); }));
      // also emits g16 = bf16(dh) and the mlp c_proj bias gradient of the top block
      // (masked by the top block's mlp resid dropout: g16 is the gradient of the c_proj OUTPUT)
      const DropSpec dr = drop_spec(m, 4u * (d.NL - 1) + 2, m->p_resid);
      MMTG_TRY(layernorm_bwd(t_dx, 1, w.layer[d.NL - 1].h_out, w.meanf, w.rstdf, P + o.lnf_w, w.dh, 0,
                             nullptr, nullptr, g_out, nullptr, M, E, st, &dr));
      MMTG_TRY(on_side([&](cudaStream_t s2) {
        return ln_param_grads(t_dx, w.layer[d.NL - 1].h_out, w.meanf, w.rstdf, g_out, G + o.lnf_w, G + o.lnf_b,
                              G + o.layer[d.NL - 1].proj2_b, M, E, s2);
      }));
    } else if (stage <= d.NL) {
      const int l = d.NL - stage;
      const LayerWs& L = w.layer[l];
      const mmtg_layer_offsets& lo = o.layer[l];
      // ---- MLP ---- (g16 = bf16(dh) and d(proj2_b) were produced by the LayerNorm backward above)
      MMTG_TRY(on_side([&](cudaStream_t s2) { return wgrad(L.a, 4 * E, g_in, E, G + lo.proj2_w, E, 4 * E, E, M, s2, wg); }));
      // (the c_fc bias gradient = column sums of du is a separate HBM-bound pass on the side
      // stream: fused into this epilogue it cost 13 us of the critical chain, measured)
      MMTG_TRY(Gemm(g_in, E, false, W + lo.proj2_w, E, false, M, 4 * E, E)
                   .out_bf16(t_du, 4 * E).dmul(L.u, 4 * E).run(st));
      MMTG_TRY(on_side([&](cudaStream_t s2) {
        MMTG_TRY(colsum(t_du, 1, 4 * E, nullptr, 0, G + lo.fc_b, M, 4 * E, s2));
        return wgrad(L.x2, E, t_du, 4 * E, G + lo.fc_w, 4 * E, E, 4 * E, M, s2, wg);
      }));
      MMTG_TRY(Gemm(t_du, 4 * E, false, W + lo.fc_w, 4 * E, false, M, E, 4 * E).out_bf16(t_dx, E).run(st));
      const DropSpec d_att = drop_spec(m, 4u * l, m->p_attn), d_r1 = drop_spec(m, 4u * l + 1, m->p_resid);
      // (dgamma / dbeta / the c_proj bias gradient only feed the optimizer: side stream)
      MMTG_TRY(layernorm_bwd(t_dx, 1, L.h_mid, L.mean2, L.rstd2, P + lo.ln2_w, w.dh, 1, nullptr, nullptr, t_g16b,
                             nullptr, M, E, st, &d_r1));
      MMTG_TRY(on_side([&](cudaStream_t s2) {
        return ln_param_grads(t_dx, L.h_mid, L.mean2, L.rstd2, t_g16b, G + lo.ln2_w, G + lo.ln2_b, G + lo.proj_b, M, E, s2);
      }));
      // ---- attention ----
      MMTG_TRY(on_side([&](cudaStream_t s2) { return wgrad(L.att, E, t_g16b, E, G + lo.proj_w, E, E, E, M, s2, wg); }));
      MMTG_TRY(Gemm(t_g16b, E, false, W + lo.proj_w, E, false, M, E, E).out_bf16(t_datt, E).run(st));
      // (the c_attn bias gradient = column sums of dqkv comes out of the attention backward's drains)
      MMTG_TRY(attn_bwd(L.qkv, b->attn_mask, L.att, t_datt, L.lse, w.delta, t_dqkv, B, d.L, d.NH, st, &d_att,
                        G + lo.attn_b));
      MMTG_TRY(on_side([&](cudaStream_t s2) {
        return wgrad(L.x1, E, t_dqkv, 3 * E, G + lo.attn_w, 3 * E, E, 3 * E, M, s2, wg);
      }));
      MMTG_TRY(Gemm(t_dqkv, 3 * E, false, W + lo.attn_w, 3 * E, false, M, E, 3 * E).out_bf16(t_dx2, E).run(st));
      // dh is now the gradient of this block's input: its bf16 copy / column sums feed the block
      // below (mlp c_proj bias) or, for block 0, the projector (projector_layer2 bias)
      // (masked by the dropout that produced this block's input: the mlp resid dropout of the
      // block below, or the embedding dropout — there the fp32 dh is masked too)
      const DropSpec d_in = l > 0 ? drop_spec(m, 4u * (l - 1) + 2, m->p_resid)
                                  : drop_spec(m, MMTG_DROP_SITE_EMBD, m->p_embd, 1);
      MMTG_TRY(wait_side_of(stage - 1));  // g_out below is the g_in the side work of the previous stage read
      MMTG_TRY(layernorm_bwd(t_dx2, 1, L.h_in, L.mean1, L.rstd1, P + lo.ln1_w, w.dh, 1, nullptr, nullptr, g_out,
                             nullptr, M, E, st, &d_in));
      MMTG_TRY(on_side([&](cudaStream_t s2) {
        return ln_param_grads(t_dx2, L.h_in, L.mean1, L.rstd1, g_out, G + lo.ln1_w, G + lo.ln1_b,
                              l > 0 ? G + o.layer[l - 1].proj2_b : G + o.proj2_b, M, E, s2);
      }));
    } else if (stage == d.NL + 1) {
      // ---------------- embeddings + projector ----------------
      MMTG_TRY(posadd_bwd(w.dh, G + o.wpe, B, d.L, E, st));
      MMTG_TRY(typeadd_bwd(w.dh, b->type_ids, G + o.wte, M, E, st));
      // dp1 = (g · W2) ⊙ (1 - p1²); W2 [E,He] as MN-major [N=He, K=E]
      MMTG_TRY(on_side([&](cudaStream_t s2) { return wgrad(g_in, E, w.p1, He, G + o.proj2_w, He, E, He, M, s2, wg); }));
      MMTG_TRY(Gemm(g_in, E, false, W + o.proj2_w, He, true, M, He, E)
                   .out_bf16(w.dp1, He).dtanh(w.p1, He).colsum(G + o.proj1_b).run(st));
      MMTG_TRY(on_side([&](cudaStream_t s2) { return wgrad(w.dp1, He, w.emb16, Dw, G + o.proj1_w, Dw, He, Dw, M, s2, wg); }));
      MMTG_TRY(Gemm(w.dp1, He, false, W + o.proj1_w, Dw, true, M, Dw, He).out_bf16(w.dE, Dw).run(st));
      MMTG_TRY(embed_bwd(w.dE, w.dctx16, nullptr, B, d.P, d.T, S, d.two_sent, Dw, st));
    } else {
      // (a stage of its own: the projector / wpe / ln_f / tied-wte gradients above are final, so a
      // data-parallel caller can all-reduce that 50 MB bucket while the encoder side runs)
      // ---------------- beta gate ----------------
      MMTG_TRY(colsum(w.dctx16, 1, Dw, nullptr, 0, G + o.beta_out_b, SB, Dw, st));
      MMTG_TRY(Gemm(w.dctx16, Dw, false, W + o.beta_out_w, He, true, SB, He, Dw).out_bf16(w.do16, He).run(st));
      MMTG_TRY(wgrad(w.dctx16, Dw, w.o16, He, G + o.beta_out_w, He, Dw, He, SB, st));
      MMTG_TRY(beta_bwd(w.topic_ln, w.actx[0], w.actx[1], P + o.beta_att_w, w.att3, w.do16, w.dtopic_ln,
                        w.dactx[0], w.dactx[1], G + o.beta_att_w, G + o.beta_att_b, B, S, He, st));
      // ---------------- topic branch ----------------
      MMTG_TRY(layernorm_bwd(w.dtopic_ln, 0, w.topic_pre, w.topic_mean, w.topic_rstd, P + o.enc_ln_w[0],
                             w.dtopic_pre, 0, G + o.enc_ln_w[0], G + o.enc_ln_b[0], nullptr, nullptr, B, He, st));
      MMTG_TRY(colsum(w.dtopic_pre, 0, He, w.dtopic_pre16, He, G + o.topic_b, B, He, st));
      MMTG_TRY(wgrad(w.dtopic_pre16, He, w.x_topic16, Dw, G + o.topic_w, Dw, He, Dw, B, st));
      // ---------------- image / text branches: alpha attention, LayerNorm, GRU ----------------
      for (int md = 0; md < 2; ++md) {
        MMTG_TRY(alpha_bwd(w.aqkv[md], w.aprobs[md], w.dactx[md], g_kl, 1.f / (float)(S * B), w.daqkv16, B,
                           d.alpha_heads, S, He / d.alpha_heads, st));
        MMTG_TRY(colsum(w.daqkv16, 1, 3 * He, nullptr, 0, G + o.alpha_qkv_b[md], SB, 3 * He, st));
        // dln = dqkv · Wqkv  (Wqkv [3He, He] as MN-major [N=He, K=3He])
        MMTG_TRY(Gemm(w.daqkv16, 3 * He, false, W + o.alpha_qkv_w[md], He, true, SB, He, 3 * He)
                     .out_f32(w.dln, He).run(st));
        MMTG_TRY(wgrad(w.daqkv16, 3 * He, w.ln16[md], He, G + o.alpha_qkv_w[md], He, 3 * He, He, SB, st));
        MMTG_TRY(layernorm_bwd(w.dln, 0, w.hout[md], w.ln_mean[md], w.ln_rstd[md], P + o.enc_ln_w[1 + md],
                               w.dhout, 0, G + o.enc_ln_w[1 + md], G + o.enc_ln_b[1 + md], nullptr, nullptr, SB, He, st));
        // GRU backward through time
        for (int t = S - 1; t >= 0; --t) {
          const float* hp = t > 0 ? w.hout[md] + (size_t)(t - 1) * B * He : nullptr;
          MMTG_TRY(gru_gate_bwd(w.dhout + (size_t)t * B * He, t < S - 1 ? w.dcarry : nullptr,
                                w.gru_save[md] + (size_t)t * B * 4 * He, hp,
                                w.dgi16 + (size_t)t * B * 3 * He, w.dgh16 + (size_t)t * B * 3 * He, w.dhz, B,
                                He, st));
          if (t > 0)  // carry = dh*z + dgh · W_hh   (W_hh [3He, He] as MN-major [N=He, K=3He])
            MMTG_TRY(Gemm(w.dgh16 + (size_t)t * B * 3 * He, 3 * He, false, W + o.gru_w_hh[md], He, true, B, He,
                          3 * He)
                         .out_f32(w.dcarry, He).residual(w.dhz, He).run(st));
        }
        MMTG_TRY(colsum(w.dgi16, 1, 3 * He, nullptr, 0, G + o.gru_b_ih[md], SB, 3 * He, st));
        MMTG_TRY(colsum(w.dgh16, 1, 3 * He, nullptr, 0, G + o.gru_b_hh[md], SB, 3 * He, st));
        MMTG_TRY(wgrad(w.dgi16, 3 * He, w.x_mod16[md], Dw, G + o.gru_w_ih[md], Dw, 3 * He, Dw, SB, st));
        if (S > 1)
          MMTG_TRY(wgrad(w.dgh16 + (size_t)B * 3 * He, 3 * He, w.hout16[md], He, G + o.gru_w_hh[md], He,
                         3 * He, He, (S - 1) * B, st));
      }
    }
    MMTG_TRY(stage_done(stage));
  }
  // all side-stream work is complete when the call returns (the caller may all-reduce the
  // gradient buckets of these stages, and a stream capture must end fully joined)
  MMTG_TRY(wait_side_of(stage_end - 1));
  MMTG_TRY(wait_side_of(stage_end - 2));
  return 0;
}
