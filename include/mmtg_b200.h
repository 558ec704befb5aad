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

#define MMTG_ABI_VERSION 1

const char* mmtg_last_error(void);
int mmtg_abi_version(void);
/* number of kernel launches issued by this library since process start (bench.py gpu_launches) */
int64_t mmtg_launch_count(void);

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
  /* optional per-row (max, sum-exp) partials of the stored value over this tile's columns,
   * layout [ceil(N/block_n)][M][2] fp32 — lets the loss kernels skip a full re-read of logits */
  float* lse_partial;
} mmtg_gemm_args;

int mmtg_gemm_bf16(const mmtg_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMTG_B200_H_ */
