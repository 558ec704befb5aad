// Fused multi-tensor optimizer step over the FLAT parameter / gradient buffers:
// global-L2 gradient clipping (torch.nn.utils.clip_grad_norm_, src/train.py:194) + the HF AdamW
// variant the reference trains with (transformers.AdamW: eps added OUTSIDE the bias-corrected
// denominator, weight decay decoupled; src/train.py:137,195) + refresh of the bf16 weight
// shadow in the same pass. HBM-bound: 16 B read + 14 B written per parameter.
#include "../../include/mmtg_b200.h"
#include "common.cuh"

namespace mmtg {

void count_launch(int n = 1);

namespace {

__global__ void __launch_bounds__(256)
sumsq_partial_kernel(const float* __restrict__ g, long long n, float* __restrict__ partial) {
  float a = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x * 4;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 4 <= n) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(g + i));
      a += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    } else {
      for (long long j = i; j < n; ++j) a += g[j] * g[j];
    }
  }
  a = warp_sum(a);
  __shared__ float s[8];
  if (lane_id() == 0) s[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) a += s[w];
    partial[blockIdx.x] = a;
  }
}
__global__ void sumsq_final_kernel(const float* __restrict__ partial, int n, float* __restrict__ out) {
  float a = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) a += partial[i];
  a = warp_sum(a);
  __shared__ float s[32];
  if (lane_id() == 0) s[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) a += s[w];
    out[0] = a;
  }
}

__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
             float* __restrict__ v, bf16* __restrict__ p16, long long n, float lr, float b1, float b2,
             float omb1, float omb2, float eps, float wd, float step_size, const float* __restrict__ normsq,
             float max_norm, const float* __restrict__ lr_dev, const int* __restrict__ step_dev,
             int correct_bias) {
  // omb1 = 1 - beta1, omb2 = 1 - beta2 rounded ONCE from double on the host, as torch rounds the
  // python scalars of `exp_avg.mul_(beta1).add_(grad, alpha=1 - beta1)`: 1.f - 0.999f is 1.3e-5 off
  if (lr_dev) lr = lr_dev[0];
  if (step_dev) {
    // device-side schedule state: the launch is replayable from a CUDA graph
    step_size = lr;
    if (correct_bias) {
      const float t = (float)(step_dev[0] + 1);
      // 1 - beta^t = -expm1(t log1p(-(1 - beta))): no cancellation at small t
      const float bc1 = -expm1f(t * log1pf(-omb1)), bc2 = -expm1f(t * log1pf(-omb2));
      step_size = lr * sqrtf(bc2) / bc1;
    }
  }
  float clip = 1.f;
  if (normsq) clip = fminf(1.f, max_norm / (sqrtf(normsq[0]) + 1e-6f));
  const long long stride = (long long)gridDim.x * blockDim.x * 4;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    float4 pv = *reinterpret_cast<float4*>(p + i);
    const float4 gv = __ldg(reinterpret_cast<const float4*>(g + i));
    float4 mv = *reinterpret_cast<float4*>(m + i);
    float4 vv = *reinterpret_cast<float4*>(v + i);
    float* pp = reinterpret_cast<float*>(&pv);
    const float* gp = reinterpret_cast<const float*>(&gv);
    float* mp = reinterpret_cast<float*>(&mv);
    float* vp = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float gr = gp[e] * clip;
      mp[e] = b1 * mp[e] + omb1 * gr;
      vp[e] = b2 * vp[e] + omb2 * gr * gr;
      pp[e] -= step_size * mp[e] / (sqrtf(vp[e]) + eps);
      if (wd != 0.f) pp[e] -= lr * wd * pp[e];
    }
    *reinterpret_cast<float4*>(p + i) = pv;
    *reinterpret_cast<float4*>(m + i) = mv;
    *reinterpret_cast<float4*>(v + i) = vv;
    if (p16) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(pv.x, pv.y), hi = __floats2bfloat162_rn(pv.z, pv.w);
      *reinterpret_cast<uint2*>(p16 + i) =
          make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
    }
  }
}

__global__ void bump_kernel(int* c) { c[0] += 1; }

}  // namespace
}  // namespace mmtg

using namespace mmtg;

extern "C" int mmtg_grad_norm_sq(const float* grads, int64_t n, float* partial_ws, int32_t partial_len,
                                 float* out_normsq, void* stream) {
  MMTG_CHECK_ARG(grads && partial_ws && out_normsq && n > 0 && partial_len >= 64, "bad grad_norm args");
  MMTG_CHECK_ARG(((uintptr_t)grads & 15) == 0, "gradient buffer must be 16-byte aligned");
  int blocks = num_sms() * 4;
  if (blocks > partial_len) blocks = partial_len;
  sumsq_partial_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(grads, n, partial_ws);
  MMTG_LAUNCH_OK();
  sumsq_final_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(partial_ws, blocks, out_normsq);
  MMTG_LAUNCH_OK();
  count_launch(2);
  return 0;
}

extern "C" int mmtg_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                               void* params_bf16, int64_t n, float lr, double beta1, double beta2,
                               float eps, float weight_decay, int32_t step, int32_t correct_bias,
                               const float* normsq, float max_norm, const float* lr_dev,
                               int32_t* step_dev, void* stream) {
  MMTG_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && n > 0 && n % 4 == 0 && (step >= 1 || step_dev),
                 "bad adamw args (n must be a multiple of 4)");
  MMTG_CHECK_ARG(!(lr_dev && !step_dev), "lr_dev requires step_dev (device-side schedule state)");
  float step_size = lr;
  if (correct_bias && !step_dev) {
    const double bc1 = 1.0 - pow(beta1, (double)step);
    const double bc2 = 1.0 - pow(beta2, (double)step);
    step_size = (float)((double)lr * sqrt(bc2) / bc1);
  }
  const int blocks = num_sms() * 8;
  ProfScope prof(2, 0, (double)n * 30.0, (cudaStream_t)stream);
  adamw_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq,
                                                          (bf16*)params_bf16, n, lr, (float)beta1, (float)beta2,
                                                          (float)(1.0 - beta1), (float)(1.0 - beta2), eps,
                                                          weight_decay, step_size, normsq, max_norm,
                                                          lr_dev, step_dev, correct_bias);
  MMTG_LAUNCH_OK();
  count_launch();
  if (step_dev) {
    bump_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_dev);
    MMTG_LAUNCH_OK();
    count_launch();
  }
  return 0;
}
