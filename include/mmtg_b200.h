/* mmtg_b200 — C ABI of the B200-native MMTG hot path (libmmtg_b200.so).
 *
 * The reference (Aman-4-Real/MMTG) has no plugin/FFI seam: its hot path is the Python class
 * surface `MMTG.forward` (src/model.py:356-400), `MyLoss.forward` (src/loss.py:45-74) and
 * `sample_sequence` (src/generate.py:97-145), which bottom out in PyTorch/ATen library calls.
 * Every entry point below replaces the library call(s) named in its comment and is what the
 * reference-side binding (ctypes stub shown in INTEGRATION.md) would bind.
 *
 * Conventions
 *  - plain pointers and sizes only; all pointers are DEVICE pointers unless stated otherwise;
 *  - asynchronous on `stream` (a cudaStream_t passed as void*), never synchronises;
 *  - no allocation, no ownership transfer: callers (PyTorch host glue) own every buffer;
 *  - return 0 on success, <0 for invalid argument, >0 = cudaError_t; text via mmtg_last_error();
 *  - bf16 = raw uint16 storage of __nv_bfloat16; matrices are row-major with explicit pitches
 *    ("ld", in elements).
 */
#ifndef MMTG_B200_H_
#define MMTG_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMTG_ABI_VERSION 3

const char* mmtg_last_error(void);
int mmtg_abi_version(void);
/* number of kernel launches issued by this library since process start (bench.py gpu_launches) */
int64_t mmtg_launch_count(void);
/* optional per-launch CUDA-event timing used by bench.py's roofline block.
 * classes: 0 = tcgen05 GEMM, 1 = attention, 2 = row / reduction kernels */
void mmtg_prof_enable(int32_t on);
void mmtg_prof_reset(void);
int mmtg_prof_collect(int32_t cls, double* ms, double* flops, double* bytes, int64_t* count);
int mmtg_prof_dump(const char* csv_path);

/* ------------------------------------------------------------------------------------------
 * Dense contraction on tcgen05/TMEM fed by TMA:  out = epilogue(A · Bᵀ)
 * Replaces every cuBLAS call on the path: nn.Linear / HF Conv1D (`addmm`,
 * HF pytorch_utils.py:97-123) forward, their dgrad and wgrad in autograd backward, the
 * projector (src/model.py:279-281) and lm_head (HF modeling_gpt2.py:706).
 *
 *   A logical [M,K], B logical [N,K]; both bf16.
 *   a_mn_major = 0: A stored [M,K] (K contiguous, pitch lda); 1: stored [K,M] (M contiguous).
 *   b_mn_major likewise for B ([N,K] vs [K,N]).
 *   value = acc (+ bias[col]);  out2 (bf16, optional) receives this pre-activation value;
 *   value = act(value); value *= gelu_new'(dgelu_src[row,col]) (optional);
 *   value += residual[row,col] (optional); value += rowtab0[rowidx0[row] or row%rowmod0][col]
 *   (optional); value += rowtab1[rowidx1[row]][col] (optional); colsum[col] += Σ_row value
 *   (optional, atomics); out = value (or out += value with fp32 atomics when accumulate != 0 or
 *   split_k > 1).
 * ------------------------------------------------------------------------------------------ */
enum { MMTG_ACT_NONE = 0, MMTG_ACT_TANH = 1, MMTG_ACT_GELU_NEW = 2 };
enum { MMTG_F32 = 0, MMTG_BF16 = 1 };

typedef struct mmtg_gemm_args {
  const void* A;
  const void* B;
  int64_t lda, ldb;
  int32_t a_mn_major, b_mn_major;
  int32_t M, N, K;
  int32_t split_k;   /* <=1: none; >1 requires fp32 out, atomically accumulated */
  int32_t block_n;   /* 0 = auto, else 128 or 256 */
  void* out;
  int64_t ldo;
  int32_t out_dtype; /* MMTG_F32 / MMTG_BF16 */
  int32_t accumulate;
  void* out2;        /* optional bf16 pre-activation copy */
  int64_t ldo2;
  const float* bias;
  int32_t act;
  int32_t _pad0;
  const float* residual;
  int64_t ldr;
  const void* dgelu_src; /* bf16 */
  int64_t ldg;
  const float* rowtab0;
  const int32_t* rowidx0;
  int64_t ldt0;
  int32_t rowmod0;
  int32_t _pad1;
  const float* rowtab1;
  const int32_t* rowidx1;
  int64_t ldt1;
  float* colsum;
  /* optional per-row (max, sum-exp) partials of the stored value over each HALF tile's columns,
   * layout [2*ceil(N/block_n)][M][2] fp32 — lets the loss kernels skip a full re-read of logits */
  float* lse_partial;
  /* 0: dgelu_src holds the pre-activation u, value *= gelu_new'(u);
   * 1: dgelu_src holds a tanh OUTPUT y, value *= (1 - y*y)   (projector backward);
   * 2: dgelu_src holds the activation derivative itself, value *= src */
  int32_t dact_tanh_out;
  /* 0: out2 receives the pre-activation value; 1 (with MMTG_ACT_GELU_NEW): out2 receives
   * gelu_new'(pre-activation), so the backward epilogue is a single multiply */
  int32_t out2_mode;
  /* inverted dropout of the value (after bias / activation / row-table adds, BEFORE the residual
   * add): element (row, col) is kept iff the counter-based mask of (*drop_seed, drop_site,
   * row * N + col) says so (mmtg_dropout_mask gives the same mask). drop_p = 0 or NULL seed: off.
   * Needs N % 4 == 0 and 16-byte aligned operands. */
  const uint64_t* drop_seed;
  uint32_t drop_site;
  float drop_p;
  /* 0: persistent grid (one CTA per SM walks the tiles); 1: one work unit per CTA pair — short-lived
   * CTAs for background work on a low-priority stream (weight gradients next to the backward chain) */
  int32_t grid_mode;
  int32_t _pad2;
} mmtg_gemm_args;

int mmtg_gemm_bf16(const mmtg_gemm_args* args, void* stream);
/* Weight-streaming variant for generation (M <= 64 rows): same argument struct, subset of the
 * epilogue (bias, act, residual, rowtab0 via rowidx0, rowtab1, fp32/bf16 out); A must be K-major,
 * b_mn_major selects [K,N] (HF Conv1D) vs [N,K] (nn.Linear / tied wte) weight storage. */
int mmtg_skinny_gemm_bf16(const mmtg_gemm_args* args, void* stream);

/* ------------------------------------------------------------------------------------------
 * Row kernels (HBM-bound). Reference call sites in each comment.
 * ------------------------------------------------------------------------------------------ */
/* torch.nn.LayerNorm fwd (src/model.py:380-382; HF GPT2Block ln_1/ln_2, GPT2Model ln_f). E in {512,768} */
int mmtg_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y_bf16,
                       float* y_f32, float* mean, float* rstd, int32_t M, int32_t E, float eps,
                       void* stream);
/* its autograd backward; dx written or accumulated; dgamma/dbeta accumulated (atomics).
 * Optional fusions for the residual stream: dx_bf16 = bf16 copy of the final dx (next GEMM
 * operand), dx_colsum[E] += column sums of the final dx (bias gradient of the layer below). */
int mmtg_layernorm_bwd(const void* dy, int32_t dy_is_bf16, const float* x, const float* mean,
                       const float* rstd, const float* gamma, float* dx, int32_t accumulate_dx,
                       float* dgamma, float* dbeta, void* dx_bf16, float* dx_colsum, int32_t M,
                       int32_t E, void* stream);
/* The parameter-gradient half of the LayerNorm backward when mmtg_layernorm_bwd ran WITHOUT
 * dgamma / dbeta / dx_colsum (E = 768, bf16 dy: the engine's chain / side-stream split):
 * dgamma[E] += sum_rows dy * xhat, dbeta[E] += sum_rows dy, dx_colsum[E] += column sums of the bf16
 * dx copy (the bias gradient autograd computes for the Linear/Conv1D feeding this LayerNorm's
 * input; HF modeling_gpt2.py GPT2Block residual adds). E % 128 == 0. */
int mmtg_ln_param_grads(const void* dy_bf16, const float* x, const float* mean, const float* rstd,
                        const void* dx_bf16, float* dgamma, float* dbeta, float* dx_colsum, int32_t M,
                        int32_t E, void* stream);
int mmtg_cast_bf16(const float* src, void* dst, int64_t n, void* stream);
/* bias gradients: out[N] += column sums of x[M,N]; optional bf16 copy of an fp32 x */
int mmtg_colsum(const void* x, int32_t x_is_bf16, int64_t ld, void* copy_bf16, int64_t ldc,
                float* out, int32_t M, int32_t N, void* stream);
/* GPT2_Decoder embedding build, src/model.py:253-277 (token->WenLan gather + context add) */
int mmtg_embed_fwd(const float* table, const int32_t* topic_ids, const int32_t* input_ids,
                   const float* ctx, void* out_bf16, int32_t B, int32_t P, int32_t T, int32_t S,
                   int32_t two_sent, int32_t D, void* stream);
int mmtg_embed_bwd(const void* dE_bf16, void* dctx_bf16, float* dctx_f32, int32_t B, int32_t P,
                   int32_t T, int32_t S, int32_t two_sent, int32_t D, void* stream);

/* HF GPT2Attention (modeling_gpt2.py:54-72): causal + key-padding softmax(QK^T/8)V, head_dim 64.
 * qkv [B*L, 3E] bf16; out [B*L, E] bf16; lse [B, n_head, L] fp32. */
int mmtg_attn_fwd(const void* qkv, const int32_t* key_mask, void* out, float* lse, int32_t B,
                  int32_t L, int32_t n_head, void* stream);
int mmtg_attn_bwd(const void* qkv, const int32_t* key_mask, const void* out, const void* dout,
                  const float* lse, float* delta_ws, void* dqkv, int32_t B, int32_t L,
                  int32_t n_head, void* stream);
/* explicit kernel choice (tests / profiling): impl 0 = default, 1 = mma.sync tiles, 2 = tcgen05/TMEM
 * (forward: any L <= 1024; backward: whole head per CTA for L <= 256, tiled dK/dV + dQ kernels up to L = 1024) */
/* debugging aid: per-(block, warp) progress codes written to host-mapped memory (null = off) */
void mmtg_attn_set_trace(int32_t* host_mapped);
/* debugging aid: CTA 0 of the tcgen05 attention backward stamps %globaltimer (ns) at its stage
 * boundaries into dev_buf (>= 64 entries; scripts/attn_bwd_trace.py); NULL disables */
void mmtg_attn_set_clk(uint64_t* dev_buf);
int mmtg_attn_fwd_ex(const void* qkv, const int32_t* key_mask, void* out, float* lse, int32_t B,
                     int32_t L, int32_t n_head, int32_t impl, void* stream);
int mmtg_attn_bwd_ex(const void* qkv, const int32_t* key_mask, const void* out, const void* dout,
                     const float* lse, float* delta_ws, void* dqkv, int32_t B, int32_t L,
                     int32_t n_head, int32_t impl, void* stream);

/* Loss reductions: HF ForCausalLMLoss (loss/loss_utils.py:28-67) and MyLoss (src/loss.py:45-74) */
int mmtg_lse_rows(const float* logits, int64_t ld, float* lse, int32_t M, int32_t V, void* stream);
int mmtg_lse_combine(const float* partials, float* lse, int32_t M, int32_t ntiles, void* stream);
int mmtg_ce_reduce(const float* logits, int64_t ld, const float* lse, const int32_t* topic_ids,
                   const int32_t* targets, float* hf_sum_ws, float* ce, float* hf_loss, int32_t B,
                   int32_t L, int32_t P, int32_t T, void* stream);
int mmtg_negloss(const float* ce, const int32_t* ratings, int32_t stage, float* loss, float* coef,
                 int32_t B, void* stream);
int mmtg_ce_bwd(const float* logits, int64_t ld, const float* lse, const int32_t* topic_ids,
                const int32_t* targets, const float* coef, const float* g_my, const float* g_hf,
                void* out, int32_t out_is_bf16, int64_t ldo, int32_t B, int32_t L, int32_t P,
                int32_t T, int32_t V, void* stream);

/* ------------------------------------------------------------------------------------------
 * Whole-path engine: MMTG.forward (src/model.py:356-400, training branch) and its backward,
 * orchestrated natively on one stream. All parameters live in ONE flat fp32 buffer (plus a
 * bf16 shadow for GEMM operands and a flat fp32 gradient buffer) addressed by element offsets.
 * ------------------------------------------------------------------------------------------ */
#define MMTG_MAX_LAYERS 48

typedef struct mmtg_dims {
  int32_t B, P, T, L;          /* L = P + T */
  int32_t S, two_sent;         /* experience steps (5); tokens per sentence pair (44) */
  int32_t Dw, He, alpha_heads; /* 2048, 512, 4 */
  int32_t E, NH, NL, V, Vp;    /* 768, 12, 12, 13317, V rounded up to a multiple of 8 */
  int32_t n_pos, _pad;
} mmtg_dims;

typedef struct mmtg_layer_offsets {
  int64_t ln1_w, ln1_b, attn_w, attn_b, proj_w, proj_b, ln2_w, ln2_b, fc_w, fc_b, proj2_w, proj2_b;
} mmtg_layer_offsets;

typedef struct mmtg_param_offsets {
  int64_t topic_w, topic_b;
  int64_t gru_w_ih[2], gru_w_hh[2], gru_b_ih[2], gru_b_hh[2]; /* 0 = image, 1 = text */
  int64_t enc_ln_w[3], enc_ln_b[3];                           /* topic, image, text */
  int64_t alpha_qkv_w[2], alpha_qkv_b[2];                     /* q|k|v stacked: [3He,He], [3He] */
  int64_t beta_att_w, beta_att_b;                             /* [S,He], [S] */
  int64_t beta_out_w, beta_out_b;
  int64_t proj1_w, proj1_b, proj2_w, proj2_b;
  int64_t wte, wpe;
  int64_t lnf_w, lnf_b;
  mmtg_layer_offsets layer[MMTG_MAX_LAYERS];
} mmtg_param_offsets;

typedef struct mmtg_model {
  mmtg_dims dims;
  mmtg_param_offsets off;
  const float* params;      /* flat fp32 master weights */
  const void* params_bf16;  /* flat bf16 shadow (same offsets) */
  float* grads;             /* flat fp32 gradients (same offsets), accumulated into */
  const float* token_table; /* [V, Dw] fp32, frozen (vocab/token_id2emb_dict.pkl) */
  /* GPT-2 dropout of the training forward (HF configuration_gpt2.py embd_pdrop / resid_pdrop /
   * attn_pdrop, live in train mode: SURVEY §5). Masks are counter-based functions of
   * (*drop_seed, site, element index): backward regenerates them, nothing is stored. The caller
   * changes *drop_seed between steps (mmtg_dropout_next_seed). All p = 0 or NULL seed: off. */
  const uint64_t* drop_seed;
  float p_embd, p_resid, p_attn;
  /* rows of token_table: a token id outside [0, table_rows) makes the embedding kernels trap
   * (the reference raises KeyError on an id missing from its dict, src/model.py:256,263) */
  int32_t table_rows;
} mmtg_model;

typedef struct mmtg_batch {
  const int32_t* topic_ids; /* [B,P] */
  const int32_t* targets;   /* [B,T] */
  const int32_t* type_ids;  /* [B,L] = cat(tpw_type_ids, type_ids) */
  const int32_t* attn_mask; /* [B,L] = cat(tpw_attention_mask, attention_mask) */
  const float* topic_emb;   /* [B,Dw] */
  const float* img_embs;    /* [B,S,Dw] */
  const float* txt_embs;    /* [B,S,Dw] */
} mmtg_batch;

int64_t mmtg_train_workspace_bytes(const mmtg_dims* dims);
/* scalars_out: [0] = HF causal-LM loss, [1] = KL regulariser. logits: [B,L,V] fp32 contiguous. */
int mmtg_train_forward(const mmtg_model* m, const mmtg_batch* b, void* workspace,
                       int64_t workspace_bytes, float* logits, float* scalars_out, int32_t save_for_backward,
                       void* stream);
/* device pointers into the workspace that the loss path reads/writes */
float* mmtg_ws_lse(const mmtg_dims* dims, void* workspace);          /* [B*L] row log-sum-exp */
void* mmtg_ws_dlogits_bf16(const mmtg_dims* dims, void* workspace);  /* [B*L, Vp] bf16 */
/* Backward stages: 0 = lm_head + ln_f, 1..NL = blocks NL-1..0, NL+1 = embeddings + projector,
 * NL+2 = multi-modal attention + encoder (NL+3 stages in all).
 * Runs stages [stage_begin, stage_end). dlogits come from mmtg_ws_dlogits_bf16 (write them, or
 * convert an fp32 [B*L, V] gradient with mmtg_dlogits_from_f32, before stage 0).
 * g_kl: device scalar d(total)/d(kl) or null (= 0). */
int mmtg_train_backward(const mmtg_model* m, const mmtg_batch* b, void* workspace,
                        int64_t workspace_bytes, const float* g_kl, int32_t stage_begin,
                        int32_t stage_end, void* stream);
/* weight-gradient GEMMs on a low-priority side stream inside mmtg_train_backward (default 1) */
int mmtg_set_wgrad_side_stream(int32_t enable);
int mmtg_dlogits_from_f32(const mmtg_dims* dims, void* workspace, const float* dlogits_f32, void* stream);

/* Dropout sites: block l uses 4*l + {0: attention probabilities [B,NH,L,L'] with L' = L rounded
 * up to even, 1: attention c_proj output [B*L,E], 2: mlp c_proj output}; MMTG_DROP_SITE_EMBD is
 * the embedding sum [B*L,E]. mmtg_dropout_mask materialises keep flags (1 = kept) of elements
 * [0, n) of a site for tests; mmtg_dropout_next_seed advances the device-side seed (capturable). */
int mmtg_attn_fwd_drop(const void* qkv, const int32_t* key_mask, void* out, float* lse, int32_t B, int32_t L,
                       int32_t n_head, const uint64_t* seed_dev, uint32_t site, float p, int32_t impl,
                       void* stream);
int mmtg_attn_bwd_drop(const void* qkv, const int32_t* key_mask, const void* out, const void* dout,
                       const float* lse, float* delta_ws, void* dqkv, int32_t B, int32_t L, int32_t n_head,
                       const uint64_t* seed_dev, uint32_t site, float p, int32_t impl, void* stream);
#define MMTG_DROP_SITE_EMBD 0xFFFFu
int mmtg_dropout_mask(const uint64_t* seed_dev, uint32_t site, float p, int64_t n, uint8_t* keep_out, void* stream);
int mmtg_dropout_next_seed(uint64_t* seed_dev, void* stream);

/* ------------------------------------------------------------------------------------------
 * KV-cached generation: the per-token body of sample_sequence (src/generate.py:117-142) and the
 * decoder inference branch (src/model.py:291-326). A batch = independent batch-1 reference runs.
 *   1. run mmtg_train_forward over [prompt | first token] (inference-branch type ids / masks);
 *   2. mmtg_decode_load_prefix moves that prefix's K/V, fused context and key mask to the cache;
 *   3. per position: mmtg_sample_rows decides gen[:, j+1] from the last logits and bumps the
 *      device-side step index *j_ptr; mmtg_decode_step consumes gen[:, *j_ptr].
 * gen: [B, gen_ld] int32 token ids (device). Everything is CUDA-graph capturable.
 * ------------------------------------------------------------------------------------------ */
int64_t mmtg_decode_workspace_bytes(const mmtg_dims* dims, int32_t Lmax);
int mmtg_decode_load_prefix(const mmtg_dims* dims_prefix, void* train_workspace, int32_t Lmax,
                            void* decode_workspace, const int32_t* prefix_mask, void* stream);
int mmtg_decode_step(const mmtg_model* m, int32_t Lmax, void* decode_workspace, const int32_t* gen,
                     int32_t gen_ld, const int32_t* j_ptr, int32_t sent_len, int32_t n_sent,
                     float* logits, void* stream);
/* Fused form of mmtg_decode_step for B <= 64, E = 768: ONE persistent kernel runs every decoder
 * block + ln_f + lm_head of the position (grid-wide barriers between phases, weights and cached
 * K/V prefetched ahead, LayerNorm folded into the following linear layer). Same arguments and
 * results as mmtg_decode_step up to rounding (the folded LayerNorm changes the operation order).
 * Split-K partials are reduced in a FIXED order through distributed shared memory of 4-CTA
 * clusters: bit-reproducible run to run. Requires mmtg_decode_fold_weights on this workspace
 * after mmtg_decode_load_prefix and after every change of the weights. */
int mmtg_decode_fold_weights(const mmtg_model* m, int32_t Lmax, void* decode_workspace, void* stream);
int mmtg_decode_step_fused(const mmtg_model* m, int32_t Lmax, void* decode_workspace, const int32_t* gen,
                           int32_t gen_ld, const int32_t* j_ptr, int32_t sent_len, int32_t n_sent,
                           float* logits, void* stream);
/* n_steps consecutive positions in ONE launch of the persistent kernel, each = embedding build +
 * projector (src/model.py:296-318; layer 1 via the per-call tables table W1^T / ctx W1^T that
 * mmtg_decode_fold_weights prepares) + all blocks + lm_head + the sampler of mmtg_sample_rows
 * (same arguments, same RNG stream). On entry gen[:, *j_ptr] must hold the tokens to consume; on
 * return gen[:, *j_ptr + 1 .. *j_ptr + n_steps] are decided, *j_ptr is advanced by n_steps and
 * `logits` holds the last position's [B, V] logits. Deterministic: no atomics on the path. */
int mmtg_decode_steps_fused(const mmtg_model* m, int32_t Lmax, void* decode_workspace, int32_t* gen, int32_t gen_ld,
                            int32_t* j_ptr, int32_t sent_len, int32_t n_sent, int32_t n_steps, float temperature,
                            int32_t top_k, float top_p, float rep_penalty, const uint64_t* seed_dev, float* logits,
                            void* stream);
/* Debug: per-phase globaltimer stamps of the fused step (CTA 0) into dev_buf (>= 80 x u64). */
int mmtg_decode_set_trace(uint64_t* dev_buf);
/* ban_specials: set ids 1, 2, 100, 102 to -inf (src/generate.py:133-136).
 * seed_dev (optional, device): overrides `seed`, so a captured launch can be re-seeded.
 * top_k in [1, 1024]: descending selection; top_k == 0: pure nucleus (top_p > 0, no survivor cap)
 * or plain softmax sampling (top_p == 0).
 * dbg_probs (optional, test hook): [B, V] probabilities of the filtered distribution (0 = filtered
 * out); when given, the PAD-continuation shortcut is skipped so the distribution is always produced */
int mmtg_sample_rows(const float* logits, int64_t ld, int32_t* gen, int32_t gen_ld, int32_t* j_ptr,
                     int32_t B, int32_t V, int32_t sent_len, float temperature, int32_t top_k,
                     float top_p, float rep_penalty, uint64_t seed, const uint64_t* seed_dev,
                     int32_t ban_specials, float* dbg_probs, void* stream);

/* ------------------------------------------------------------------------------------------
 * Optimizer step over the flat buffers: clip_grad_norm_ (src/train.py:194) and
 * transformers.AdamW as the reference configures it (src/train.py:137: eps 1e-6 outside the
 * bias-corrected denominator, decoupled weight decay 0), refreshing the bf16 shadow in-pass.
 * ------------------------------------------------------------------------------------------ */
int mmtg_grad_norm_sq(const float* grads, int64_t n, float* partial_ws, int32_t partial_len,
                      float* out_normsq, void* stream);
/* lr_dev / step_dev (optional, device): learning rate and 0-based step counter kept on the device
 * (the counter is bumped after the update) so the call can be replayed from a CUDA graph. */
int mmtg_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                    void* params_bf16, int64_t n, float lr, double beta1, double beta2, float eps,
                    float weight_decay, int32_t step, int32_t correct_bias, const float* normsq,
                    float max_norm, const float* lr_dev, int32_t* step_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMTG_B200_H_ */
