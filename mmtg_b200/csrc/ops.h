// Internal prototypes of the per-op launchers (defined across the .cu files of this library).
#pragma once
#include "common.cuh"

namespace mmtg {

// dropout of one site (common.cuh: drop_key / drop_bits); seed == nullptr or p == 0: off
struct DropSpec {
  const unsigned long long* seed;
  uint32_t site;
  float p;
  int mask_dx32;  // layernorm_bwd: also mask the fp32 dx (embedding dropout)
};

// elementwise.cu
int layernorm_fwd(const float* x, const float* gamma, const float* beta, bf16* y16, float* y32,
                  float* mean, float* rstd, int M, int E, float eps, cudaStream_t st);
int layernorm_bwd(const void* dy, int dy_bf16, const float* x, const float* mean, const float* rstd,
                  const float* gamma, float* dx, int accumulate_dx, float* dgamma, float* dbeta,
                  bf16* dx16, float* dx_colsum, int M, int E, cudaStream_t st, const DropSpec* drop = nullptr);
// dgamma / dbeta / column sums of dx16 of a LayerNorm backward whose chain part ran without them
int ln_param_grads(const bf16* dy, const float* x, const float* mean, const float* rstd, const bf16* dx16,
                   float* dgamma, float* dbeta, float* dx_colsum, int M, int E, cudaStream_t st);
int cast_bf16(const float* src, bf16* dst, long long n, cudaStream_t st);
int colsum(const void* x, int x_bf16, long long ld, bf16* copy16, long long ldc, float* out, int M,
           int N, cudaStream_t st);
int embed_fwd(const float* table, const int* topic_ids, const int* input_ids, const float* ctx,
              bf16* out, int B, int P, int T, int S, int two_sent, int D, int table_rows, cudaStream_t st);
int embed_bwd(const bf16* dE, bf16* dctx16, float* dctx32, int B, int P, int T, int S, int two_sent,
              int D, cudaStream_t st);
int posadd_bwd(const float* dh, float* dwpe, int B, int L, int E, cudaStream_t st);
int typeadd_bwd(const float* dh, const int* type_ids, float* dwte, int M, int E, cudaStream_t st);
int dlogits_f32_to_bf16(const float* src, bf16* dst, int M, int V, int Vp, cudaStream_t st);

// attention.cu
int attn_fwd(const bf16* qkv, const int* kmask, bf16* out, float* lse, int B, int L, int NH,
             cudaStream_t st, const DropSpec* drop = nullptr);
int attn_bwd(const bf16* qkv, const int* kmask, const bf16* out, const bf16* dout, const float* lse,
             float* delta, bf16* dqkv, int B, int L, int NH, cudaStream_t st, const DropSpec* drop = nullptr,
             float* dbias = nullptr);  // dbias [3E] += column sums of dqkv (c_attn bias gradient)

// loss.cu
int lse_rows(const float* logits, long long ld, float* lse, int M, int V, cudaStream_t st);
int lse_combine(const float* part, float* lse, int M, int ntiles, cudaStream_t st);
int ce_reduce(const float* logits, long long ld, const float* lse, const int* topic_ids,
              const int* targets, float* hf_sum, float* ce, float* hf_loss, int B, int L, int P, int T,
              cudaStream_t st);
int sum_scale(const float* x, float* out, int n, float scale, cudaStream_t st);

// encoder.cu
int pack_sb(const float* in, bf16* out, int B, int S, int D, cudaStream_t st);
int gru_gate_fwd(const float* gi, const float* gh, const float* b_hh, const float* h_prev,
                 float* h_out, bf16* h_out16, float* save, int B, int H, cudaStream_t st);
int gru_gate_bwd(const float* dh_out, const float* dh_carry, const float* save, const float* h_prev,
                 bf16* dgi, bf16* dgh, float* dhz, int B, int H, cudaStream_t st);
int alpha_fwd(const float* qkv, float* ctx, float* probs, float* klpart, int B, int heads, int S,
              int DH, cudaStream_t st);
int alpha_bwd(const float* qkv, const float* probs, const float* dctx, const float* g_kl,
              float kl_scale, bf16* dqkv, int B, int heads, int S, int DH, cudaStream_t st);
int beta_fwd(const float* topic, const float* img, const float* txt, const float* att_w,
             const float* att_b, bf16* o16, float* att, int B, int S, int H, cudaStream_t st);
int beta_bwd(const float* topic, const float* img, const float* txt, const float* att_w,
             const float* att, const bf16* do16, float* dtopic, float* dimg, float* dtxt,
             float* datt_w, float* datt_b, int B, int S, int H, cudaStream_t st);

// decode_mega.cu: buffers of the fused decode step (carved from the decode workspace)
struct MegaBufs {
  float *h, *h2;                            // fp32 residual stream [64][E]: block input / after attention
  bf16 *h16, *h2_16, *qkv16, *att16, *u16;  // bf16 GEMM operands: [64][E], [64][E], [64][3E], [64][E], [64][4E]
  float *stats1, *stats2;                   // [64][32][2] per-row (sum, sumsq) partials per column chunk (LN1/ln_f, LN2)
  bf16 *kcache, *vcache;
  int* keymask;
  unsigned int* barrier;  // [0] grid-barrier counter (runs on across launches), [1] its value at launch start
  // full-step mode: projector layer 1 folded into two tables (layer 1 is linear in table[tok] + ctx)
  float *T1, *C1;             // table W1^T [V][He], ctx W1^T [S*B][He]  (fp32, bias b1 NOT included)
  bf16 *w2t, *table16, *ctx16;  // projector layer 2 transposed [He][E]; bf16 staging of the GEMM operands
  // LayerNorm-folded weights (mmtg_decode_fold_weights)
  bf16 *f_attn, *f_fc, *f_wte;  // [NL][E][3E], [NL][E][4E], [V][E]
  float* f_vec;                 // per layer: cs_attn[3E] b_attn[3E] cs_fc[4E] b_fc[4E]; then cs_head[V] b_head[V]
};
struct MegaStepArgs {  // full-step mode (embedding + projector prologue and sampler inside the kernel)
  int n_steps;
  int* gen;
  int gen_ld, sent_len, n_sent;
  float temperature;
  int top_k;
  float top_p, rep_penalty;
  const unsigned long long* seed_dev;
};
int decode_mega_launch(const mmtg_model* m, int Lmax, const MegaBufs& bufs, int* j_ptr, float* logits,
                       const MegaStepArgs* full, cudaStream_t st);
int decode_fold_weights(const mmtg_model* m, const MegaBufs& bufs, cudaStream_t st);

}  // namespace mmtg
