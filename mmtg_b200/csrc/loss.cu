// Masked cross-entropy / curriculum negative-sampling loss reductions over the logits
// [B, L, V] fp32 (V = 13317: rows are NOT 16-byte aligned, so accesses are coalesced scalars).
//
// Replaces HF ForCausalLMLoss (transformers loss/loss_utils.py:28-67), MyLoss.forward's Python
// loop over the batch (src/loss.py:62-74) and their autograd backward (log_softmax + nll).
// HBM-bound: forward reads the logits once (or not at all when the lm_head GEMM epilogue
// already produced per-tile (max, sum-exp) partials); backward reads them once and writes
// dlogits once.
#include "../../include/mmtg_b200.h"
#include "common.cuh"

namespace mmtg {

void count_launch(int n = 1);

namespace {

constexpr float NEAR_0 = 1e-10f;

__device__ __forceinline__ void online_merge(float& m, float& s, float om, float os) {
  const float nm = fmaxf(m, om);
  if (nm == -INFINITY) return;
  s = s * __expf(m - nm) + os * __expf(om - nm);
  m = nm;
}

// lse[row] = logsumexp(logits[row, :V]); one 256-thread block per row.
__global__ void __launch_bounds__(256)
lse_rows_kernel(const float* __restrict__ logits, long long ld, float* __restrict__ lse, int V) {
  const long long row = blockIdx.x;
  const float* z = logits + row * ld;
  float m = -INFINITY, s = 0.f;
  for (int c = threadIdx.x; c < V; c += 256) {
    const float v = __ldg(z + c);
    if (v > m) {
      s = s * __expf(m - v) + 1.f;
      m = v;
    } else {
      s += __expf(v - m);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o), os = __shfl_xor_sync(0xffffffffu, s, o);
    online_merge(m, s, om, os);
  }
  __shared__ float sm[8], ss[8];
  if (lane_id() == 0) {
    sm[threadIdx.x >> 5] = m;
    ss[threadIdx.x >> 5] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) online_merge(m, s, sm[w], ss[w]);
    lse[row] = m + logf(s);
  }
}

// lse[row] from the GEMM epilogue's partials [ntiles][M][2] = (max, sum-exp).
__global__ void lse_combine_kernel(const float* __restrict__ part, float* __restrict__ lse, int M,
                                   int ntiles) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= M) return;
  float m = -INFINITY, s = 0.f;
  for (int t = 0; t < ntiles; ++t) {
    const float2 p = *reinterpret_cast<const float2*>(part + ((long long)t * M + row) * 2);
    online_merge(m, s, p.x, p.y);
  }
  lse[row] = m + logf(s);
}

// Labels are vocabulary ids: anything outside [0, bound) — including torch's ignore_index -100,
// which the reference path never produces (src/loss.py:62-71, no -100 anywhere) — traps instead
// of reading outside the logits row.
__device__ __forceinline__ int label_at(const int* topic_ids, const int* targets, int b, int pos,
                                        int P, int T, int bound) {
  const int lab = pos < P ? topic_ids[b * P + pos] : targets[b * T + (pos - P)];
  if ((unsigned)lab >= (unsigned)bound) {
    printf("mmtg: label %d at (row %d, position %d) is outside the vocabulary [0, %d)\n", lab, b, pos, bound);
    __trap();
  }
  return lab;
}

// Per sample b: nll(b,t) = lse[b,t] - z[b,t,label(b,t+1)] for t in [0, L-2].
//   hf_sum[b] = sum_t nll      (HF loss = sum_b hf_sum / (B*(L-1)), PAD labels included)
//   ce[b]     = mean over t in [P, L-2] of nll   (MyLoss per-sample CE, src/loss.py:62-71)
// topic_ids may be null when P == 0 rows are not needed (generic MyLoss path).
__global__ void __launch_bounds__(256)
ce_reduce_kernel(const float* __restrict__ logits, long long ld, const float* __restrict__ lse,
                 const int* __restrict__ topic_ids, const int* __restrict__ targets,
                 float* __restrict__ hf_sum, float* __restrict__ ce, int L, int P, int T) {
  const int b = blockIdx.x;
  float a_hf = 0.f, a_ce = 0.f;
  for (int t = threadIdx.x; t < L - 1; t += 256) {
    if (t + 1 < P && topic_ids == nullptr) continue;  // generic MyLoss path: prompt labels unknown
    const int lab = label_at(topic_ids, targets, b, t + 1, P, T, (int)ld);
    const long long row = (long long)b * L + t;
    const float nll = lse[row] - __ldg(logits + row * ld + lab);
    a_hf += nll;
    if (t >= P) a_ce += nll;
  }
  a_hf = warp_sum(a_hf);
  a_ce = warp_sum(a_ce);
  __shared__ float s1[8], s2[8];
  if (lane_id() == 0) {
    s1[threadIdx.x >> 5] = a_hf;
    s2[threadIdx.x >> 5] = a_ce;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) {
      a_hf += s1[w];
      a_ce += s2[w];
    }
    if (hf_sum) hf_sum[b] = a_hf;
    if (ce) ce[b] = (L - 1 - P) > 0 ? a_ce / (float)(L - 1 - P) : 0.f;
  }
}

// out[0] = sum(x[0..n)) * scale   (tiny; one block)
__global__ void sum_scale_kernel(const float* __restrict__ x, float* __restrict__ out, int n,
                                 float scale) {
  float a = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) a += x[i];
  a = warp_sum(a);
  __shared__ float s[32];
  if (lane_id() == 0) s[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) a += s[w];
    out[0] = a * scale;
  }
}

// MyLoss tail (src/loss.py:57-60,72-74): y = [rating > thr]; p = exp(-ce);
// l = -y log(p+e) - (1-y) log(1-p+e); loss = mean_b l; coef[b] = dl/dce.
__global__ void negloss_kernel(const float* __restrict__ ce, const int* __restrict__ ratings,
                               int thr, float* __restrict__ loss, float* __restrict__ coef, int B) {
  float a = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float y = ratings[b] > thr ? 1.f : 0.f;
    const float p = 1.f / expf(ce[b]);
    a += -y * logf(p + NEAR_0) - (1.f - y) * logf(1.f - p + NEAR_0);
    if (coef) coef[b] = y * p / (p + NEAR_0) - (1.f - y) * p / (1.f - p + NEAR_0);
  }
  a = warp_sum(a);
  __shared__ float s[32];
  if (lane_id() == 0) s[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) a += s[w];
    loss[0] = a / (float)B;
  }
}

// dlogits[b,t,c] = rowcoef(b,t) * (softmax(z[b,t])[c] - [c == label(b,t+1)])
//   rowcoef = g_my * coef[b] / (B*(L-1-P)) for t in [P, L-2]   (MyLoss)
//           + g_hf / (B*(L-1))             for t in [0, L-2]   (HF loss, normally unused)
// Rows with rowcoef == 0 are written as zeros. OUT_BF16 writes a padded [M, ldo] bf16 matrix
// that feeds the lm_head dgrad/wgrad GEMMs directly (columns V..ldo-1 zeroed).
template <bool OUT_BF16>
__global__ void __launch_bounds__(256)
ce_bwd_kernel(const float* __restrict__ logits, long long ld, const float* __restrict__ lse,
              const int* __restrict__ topic_ids, const int* __restrict__ targets,
              const float* __restrict__ coef, const float* __restrict__ g_my,
              const float* __restrict__ g_hf, void* __restrict__ out_, long long ldo, int B, int L,
              int P, int T, int V) {
  const long long row = blockIdx.x;
  const int b = (int)(row / L), t = (int)(row - (long long)b * L);
  float rc = 0.f;
  if (t <= L - 2) {
    if (g_hf) rc += g_hf[0] / ((float)B * (float)(L - 1));
    if (g_my && t >= P) rc += g_my[0] * coef[b] / ((float)B * (float)(L - 1 - P));
  }
  const int ncols = OUT_BF16 ? (int)ldo : V;
  if (rc == 0.f) {
    for (int c = threadIdx.x; c < ncols; c += 256) {
      if (OUT_BF16) ((bf16*)out_)[row * ldo + c] = __float2bfloat16(0.f);
      else ((float*)out_)[row * ldo + c] = 0.f;
    }
    return;
  }
  const int lab = label_at(topic_ids, targets, b, t + 1, P, T, V);
  const float* z = logits + row * ld;
  const float l = lse[row];
  for (int c = threadIdx.x; c < ncols; c += 256) {
    float v = 0.f;
    if (c < V) v = rc * (__expf(__ldg(z + c) - l) - (c == lab ? 1.f : 0.f));
    if (OUT_BF16) ((bf16*)out_)[row * ldo + c] = __float2bfloat16(v);
    else ((float*)out_)[row * ldo + c] = v;
  }
}

}  // namespace

int lse_rows(const float* logits, long long ld, float* lse, int M, int V, cudaStream_t st) {
  ProfScope prof(2, 0, (double)M * V * 4, st);
  lse_rows_kernel<<<M, 256, 0, st>>>(logits, ld, lse, V);
  MMTG_LAUNCH_OK();
  count_launch();
  return 0;
}
int lse_combine(const float* part, float* lse, int M, int ntiles, cudaStream_t st) {
  lse_combine_kernel<<<cdiv(M, 256), 256, 0, st>>>(part, lse, M, ntiles);
  MMTG_LAUNCH_OK();
  count_launch();
  return 0;
}
int ce_reduce(const float* logits, long long ld, const float* lse, const int* topic_ids,
              const int* targets, float* hf_sum, float* ce, float* hf_loss, int B, int L, int P, int T,
              cudaStream_t st) {
  ce_reduce_kernel<<<B, 256, 0, st>>>(logits, ld, lse, topic_ids, targets, hf_sum, ce, L, P, T);
  MMTG_LAUNCH_OK();
  count_launch();
  if (hf_loss) {
    sum_scale_kernel<<<1, 256, 0, st>>>(hf_sum, hf_loss, B, 1.f / ((float)B * (float)(L - 1)));
    MMTG_LAUNCH_OK();
    count_launch();
  }
  return 0;
}
int sum_scale(const float* x, float* out, int n, float scale, cudaStream_t st) {
  sum_scale_kernel<<<1, 256, 0, st>>>(x, out, n, scale);
  MMTG_LAUNCH_OK();
  count_launch();
  return 0;
}
int negloss(const float* ce, const int* ratings, int stage, float* loss, float* coef, int B,
            cudaStream_t st) {
  negloss_kernel<<<1, 256, 0, st>>>(ce, ratings, stage == 1 ? 4 : 3, loss, coef, B);
  MMTG_LAUNCH_OK();
  count_launch();
  return 0;
}
int ce_bwd(const float* logits, long long ld, const float* lse, const int* topic_ids,
           const int* targets, const float* coef, const float* g_my, const float* g_hf, void* out,
           int out_bf16, long long ldo, int B, int L, int P, int T, int V, cudaStream_t st) {
  ProfScope prof(2, 0, (double)B * L * ((double)V * 4 + (double)ldo * (out_bf16 ? 2 : 4)), st);
  if (out_bf16)
    ce_bwd_kernel<true><<<B * L, 256, 0, st>>>(logits, ld, lse, topic_ids, targets, coef, g_my, g_hf, out, ldo, B, L, P, T, V);
  else
    ce_bwd_kernel<false><<<B * L, 256, 0, st>>>(logits, ld, lse, topic_ids, targets, coef, g_my, g_hf, out, ldo, B, L, P, T, V);
  MMTG_LAUNCH_OK();
  count_launch();
  return 0;
}

}  // namespace mmtg

using namespace mmtg;

extern "C" int mmtg_lse_rows(const float* logits, int64_t ld, float* lse, int32_t M, int32_t V, void* stream) {
  MMTG_CHECK_ARG(logits && lse && M > 0 && V > 0, "bad lse args");
  return lse_rows(logits, ld, lse, M, V, (cudaStream_t)stream);
}
extern "C" int mmtg_lse_combine(const float* partials, float* lse, int32_t M, int32_t ntiles, void* stream) {
  MMTG_CHECK_ARG(partials && lse && M > 0 && ntiles > 0, "bad lse_combine args");
  return lse_combine(partials, lse, M, ntiles, (cudaStream_t)stream);
}
extern "C" int mmtg_ce_reduce(const float* logits, int64_t ld, const float* lse, const int32_t* topic_ids,
                              const int32_t* targets, float* hf_sum_ws, float* ce, float* hf_loss,
                              int32_t B, int32_t L, int32_t P, int32_t T, void* stream) {
  MMTG_CHECK_ARG(logits && lse && targets && B > 0 && L > P + 1 && L == P + T, "bad ce_reduce args");
  MMTG_CHECK_ARG(!(hf_loss && !hf_sum_ws), "hf_loss needs the hf_sum workspace");
  return ce_reduce(logits, ld, lse, topic_ids, targets, hf_sum_ws, ce, hf_loss, B, L, P, T, (cudaStream_t)stream);
}
extern "C" int mmtg_negloss(const float* ce, const int32_t* ratings, int32_t stage, float* loss,
                            float* coef, int32_t B, void* stream) {
  MMTG_CHECK_ARG(ce && ratings && loss && B > 0, "bad negloss args");
  return negloss(ce, ratings, stage, loss, coef, B, (cudaStream_t)stream);
}
extern "C" int mmtg_ce_bwd(const float* logits, int64_t ld, const float* lse, const int32_t* topic_ids,
                           const int32_t* targets, const float* coef, const float* g_my,
                           const float* g_hf, void* out, int32_t out_is_bf16, int64_t ldo, int32_t B,
                           int32_t L, int32_t P, int32_t T, int32_t V, void* stream) {
  MMTG_CHECK_ARG(logits && lse && targets && out && (g_my || g_hf), "bad ce_bwd args");
  MMTG_CHECK_ARG(!(g_my && !coef), "g_my needs coef");
  MMTG_CHECK_ARG(!(g_hf && !topic_ids && P > 0), "g_hf needs topic_ids");
  return ce_bwd(logits, ld, lse, topic_ids, targets, coef, g_my, g_hf, out, out_is_bf16, ldo, B, L, P, T, V, (cudaStream_t)stream);
}
