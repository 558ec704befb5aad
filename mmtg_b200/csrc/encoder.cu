// Small fused kernels of the multi-modal encoder side (<0.3 % of the step's FLOPs, but on the
// critical path and needed for gradient parity):
//   * GRU cell gates forward / backward (torch.nn.GRU semantics, gate order r, z, n;
//     src/model.py:47-48,78-79) — the input and recurrent contractions run on the tcgen05 GEMM;
//   * inner-modal "alpha" attention over the 5 steps with its Gaussian-prior KL regulariser
//     (src/model.py:133-161), forward and backward, one warp per (sample, head);
//   * multi-modal "beta" gate (3-way softmax over topic / image_i / text_i, src/model.py:181-202),
//     forward and backward, one warp per sample.
// Activations on this side stay fp32 (the KL term needs log P); only GEMM operands are bf16.
// Row order of every [S*B, *] matrix here is (step s, sample b): row = s*B + b.
#include "../../include/mmtg_b200.h"
#include "common.cuh"

namespace mmtg {

void count_launch(int n = 1);

namespace {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// [B, S, D] fp32 -> [S, B, D] bf16 (encoder GEMM A operands); S == 1 is a plain cast.
__global__ void __launch_bounds__(256)
pack_sb_kernel(const float* __restrict__ in, bf16* __restrict__ out, int B, int S, int D) {
  const int row = blockIdx.x;  // output row s*B + b
  const int s = row / B, b = row - s * B;
  const float* src = in + ((long long)b * S + s) * D;
  for (int c = threadIdx.x * 4; c < D; c += blockDim.x * 4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(src + c));
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    *reinterpret_cast<uint2*>(out + (long long)row * D + c) =
        make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
  }
}

// ------------------------------------------------------------------------------------------
// GRU step. gi = x_t W_ih^T + b_ih (precomputed for all t), gh = h_{t-1} W_hh^T + b_hh
// (null at t = 0 where h = 0 -> gh = b_hh).
//   r = s(gi_r + gh_r), z = s(gi_z + gh_z), n = tanh(gi_n + r * gh_n), h = (1-z) n + z h_prev
// Saves r, z, n and gh_n for the backward pass.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gru_gate_fwd_kernel(const float* __restrict__ gi, const float* __restrict__ gh,
                    const float* __restrict__ b_hh, const float* __restrict__ h_prev,
                    float* __restrict__ h_out, bf16* __restrict__ h_out16, float* __restrict__ save,
                    int B, int H) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * H) return;
  const int b = i / H, c = i - b * H;
  const float* gir = gi + (long long)b * 3 * H;
  float ghr, ghz, ghn;
  if (gh) {
    const float* g = gh + (long long)b * 3 * H;
    ghr = g[c]; ghz = g[H + c]; ghn = g[2 * H + c];
  } else {
    ghr = b_hh[c]; ghz = b_hh[H + c]; ghn = b_hh[2 * H + c];
  }
  const float r = sigmoidf_(gir[c] + ghr);
  const float z = sigmoidf_(gir[H + c] + ghz);
  const float n = tanhf(gir[2 * H + c] + r * ghn);
  const float hp = h_prev ? h_prev[i] : 0.f;
  const float h = (1.f - z) * n + z * hp;
  h_out[i] = h;
  h_out16[i] = __float2bfloat16(h);
  float* sv = save + (long long)b * 4 * H;
  sv[c] = r; sv[H + c] = z; sv[2 * H + c] = n; sv[3 * H + c] = ghn;
}

// dh = dH_out[t] (+ carry); produces dgi, dgh (bf16, GEMM operands) and dh*z (fp32 residual of the
// carry GEMM dh_prev = dh*z + dgh W_hh).
__global__ void __launch_bounds__(256)
gru_gate_bwd_kernel(const float* __restrict__ dh_out, const float* __restrict__ dh_carry,
                    const float* __restrict__ save, const float* __restrict__ h_prev,
                    bf16* __restrict__ dgi, bf16* __restrict__ dgh, float* __restrict__ dhz, int B,
                    int H) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * H) return;
  const int b = i / H, c = i - b * H;
  const float* sv = save + (long long)b * 4 * H;
  const float r = sv[c], z = sv[H + c], n = sv[2 * H + c], ghn = sv[3 * H + c];
  const float dh = dh_out[i] + (dh_carry ? dh_carry[i] : 0.f);
  const float hp = h_prev ? h_prev[i] : 0.f;
  const float dn = dh * (1.f - z) * (1.f - n * n);
  const float dz = dh * (hp - n) * z * (1.f - z);
  const float dr = dn * ghn * r * (1.f - r);
  bf16* gi = dgi + (long long)b * 3 * H;
  bf16* gh = dgh + (long long)b * 3 * H;
  gi[c] = __float2bfloat16(dr); gi[H + c] = __float2bfloat16(dz); gi[2 * H + c] = __float2bfloat16(dn);
  gh[c] = __float2bfloat16(dr); gh[H + c] = __float2bfloat16(dz); gh[2 * H + c] = __float2bfloat16(dn * r);
  dhz[i] = dh * z;
}

// ------------------------------------------------------------------------------------------
// Alpha attention: one warp per (b, head); S = 5 steps, head_dim = 128 (4 dims per lane).
// qkv: [S*B, 3*Hd] fp32 (q | k | v). ctx out: [S*B, Hd] fp32. probs saved [B, heads, S, S].
// klpart[b*heads + h] = sum_{i,j} t_ij (log t_ij - log P_ij)   (caller scales by 1/(S*B)).
// ------------------------------------------------------------------------------------------
constexpr int AS = 5;
__constant__ float c_prior[AS * AS];  // t_ij, src/model.py:116-120

template <int DH>
__global__ void __launch_bounds__(128)
alpha_fwd_kernel(const float* __restrict__ qkv, float* __restrict__ ctx, float* __restrict__ probs,
                 float* __restrict__ klpart, int B, int heads) {
  constexpr int PER = DH / 32;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= B * heads) return;
  const int b = w / heads, h = w - b * heads, l = lane_id();
  const int Hd = heads * DH;
  float q[AS][PER], k[AS][PER], v[AS][PER];
#pragma unroll
  for (int s = 0; s < AS; ++s) {
    const float* row = qkv + ((long long)s * B + b) * 3 * Hd + h * DH + l * PER;
#pragma unroll
    for (int e = 0; e < PER; ++e) {
      q[s][e] = row[e];
      k[s][e] = row[Hd + e];
      v[s][e] = row[2 * Hd + e];
    }
  }
  const float scale = rsqrtf((float)DH);
  float kl = 0.f;
#pragma unroll
  for (int i = 0; i < AS; ++i) {
    float sc[AS];
#pragma unroll
    for (int j = 0; j < AS; ++j) {
      float d = 0.f;
#pragma unroll
      for (int e = 0; e < PER; ++e) d += q[i][e] * k[j][e];
      sc[j] = warp_sum(d) * scale;
    }
    float m = sc[0];
#pragma unroll
    for (int j = 1; j < AS; ++j) m = fmaxf(m, sc[j]);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < AS; ++j) {
      sc[j] = expf(sc[j] - m);
      sum += sc[j];
    }
    float o[PER];
#pragma unroll
    for (int e = 0; e < PER; ++e) o[e] = 0.f;
#pragma unroll
    for (int j = 0; j < AS; ++j) {
      const float p = sc[j] / sum;
      const float t = c_prior[i * AS + j];
      kl += t * (logf(t) - logf(p));
      if (l == 0) probs[((long long)w * AS + i) * AS + j] = p;
#pragma unroll
      for (int e = 0; e < PER; ++e) o[e] += p * v[j][e];
    }
    float* dst = ctx + ((long long)i * B + b) * Hd + h * DH + l * PER;
#pragma unroll
    for (int e = 0; e < PER; ++e) dst[e] = o[e];
  }
  if (l == 0) klpart[w] = kl;
}

// Backward of alpha attention + KL. dctx [S*B, Hd] fp32; g_kl: device scalar d(total)/d(kl).
// dqkv out: [S*B, 3*Hd] bf16 (operand of the QKV dgrad / wgrad GEMMs).
template <int DH>
__global__ void __launch_bounds__(128)
alpha_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ probs,
                 const float* __restrict__ dctx, const float* __restrict__ g_kl, float kl_scale,
                 bf16* __restrict__ dqkv, int B, int heads) {
  constexpr int PER = DH / 32;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= B * heads) return;
  const int b = w / heads, h = w - b * heads, l = lane_id();
  const int Hd = heads * DH;
  float q[AS][PER], k[AS][PER], v[AS][PER], dc[AS][PER];
#pragma unroll
  for (int s = 0; s < AS; ++s) {
    const float* row = qkv + ((long long)s * B + b) * 3 * Hd + h * DH + l * PER;
    const float* drow = dctx + ((long long)s * B + b) * Hd + h * DH + l * PER;
#pragma unroll
    for (int e = 0; e < PER; ++e) {
      q[s][e] = row[e];
      k[s][e] = row[Hd + e];
      v[s][e] = row[2 * Hd + e];
      dc[s][e] = drow[e];
    }
  }
  const float gk = (g_kl ? g_kl[0] : 0.f) * kl_scale;
  const float scale = rsqrtf((float)DH);
  float dq[AS][PER], dk[AS][PER], dv[AS][PER];
#pragma unroll
  for (int s = 0; s < AS; ++s)
#pragma unroll
    for (int e = 0; e < PER; ++e) dq[s][e] = dk[s][e] = dv[s][e] = 0.f;
#pragma unroll
  for (int i = 0; i < AS; ++i) {
    float p[AS], dp[AS];
    float dot = 0.f;
#pragma unroll
    for (int j = 0; j < AS; ++j) {
      p[j] = probs[((long long)w * AS + i) * AS + j];
      float d = 0.f;
#pragma unroll
      for (int e = 0; e < PER; ++e) d += dc[i][e] * v[j][e];
      dp[j] = warp_sum(d) - gk * c_prior[i * AS + j] / p[j];
      dot += p[j] * dp[j];
#pragma unroll
      for (int e = 0; e < PER; ++e) dv[j][e] += p[j] * dc[i][e];
    }
#pragma unroll
    for (int j = 0; j < AS; ++j) {
      const float ds = p[j] * (dp[j] - dot) * scale;
#pragma unroll
      for (int e = 0; e < PER; ++e) {
        dq[i][e] += ds * k[j][e];
        dk[j][e] += ds * q[i][e];
      }
    }
  }
#pragma unroll
  for (int s = 0; s < AS; ++s) {
    bf16* row = dqkv + ((long long)s * B + b) * 3 * Hd + h * DH + l * PER;
#pragma unroll
    for (int e = 0; e < PER; ++e) {
      row[e] = __float2bfloat16(dq[s][e]);
      row[Hd + e] = __float2bfloat16(dk[s][e]);
      row[2 * Hd + e] = __float2bfloat16(dv[s][e]);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Beta gate: one warp per sample b, loops over the S steps.
// ------------------------------------------------------------------------------------------
template <int H>
__global__ void __launch_bounds__(128)
beta_fwd_kernel(const float* __restrict__ topic, const float* __restrict__ img,
                const float* __restrict__ txt, const float* __restrict__ att_w,
                const float* __restrict__ att_b, bf16* __restrict__ o16, float* __restrict__ att,
                int B, int S) {
  constexpr int PER = H / 32;
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= B) return;
  const int l = lane_id();
  float x0[PER];
#pragma unroll
  for (int e = 0; e < PER; ++e) x0[e] = topic[(long long)b * H + l + 32 * e];
  for (int s = 0; s < S; ++s) {
    const long long row = (long long)s * B + b;
    float x1[PER], x2[PER], d0 = 0.f, d1 = 0.f, d2 = 0.f;
#pragma unroll
    for (int e = 0; e < PER; ++e) {
      const int c = l + 32 * e;
      const float wv = att_w[s * H + c];
      x1[e] = img[row * H + c];
      x2[e] = txt[row * H + c];
      d0 += wv * x0[e]; d1 += wv * x1[e]; d2 += wv * x2[e];
    }
    const float bb = att_b[s];
    d0 = warp_sum(d0) + bb; d1 = warp_sum(d1) + bb; d2 = warp_sum(d2) + bb;
    const float m = fmaxf(d0, fmaxf(d1, d2));
    const float e0 = expf(d0 - m), e1 = expf(d1 - m), e2 = expf(d2 - m);
    const float inv = 1.f / (e0 + e1 + e2);
    const float a0 = e0 * inv, a1 = e1 * inv, a2 = e2 * inv;
    if (l == 0) {
      att[row * 3] = a0; att[row * 3 + 1] = a1; att[row * 3 + 2] = a2;
    }
#pragma unroll
    for (int e = 0; e < PER; ++e)
      o16[row * H + l + 32 * e] = __float2bfloat16(a0 * x0[e] + a1 * x1[e] + a2 * x2[e]);
  }
}

// do16: [S*B, H] bf16 gradient of the gate output o (from the out_linear dgrad GEMM).
template <int H>
__global__ void __launch_bounds__(128)
beta_bwd_kernel(const float* __restrict__ topic, const float* __restrict__ img,
                const float* __restrict__ txt, const float* __restrict__ att_w,
                const float* __restrict__ att, const bf16* __restrict__ do16,
                float* __restrict__ dtopic, float* __restrict__ dimg, float* __restrict__ dtxt,
                float* __restrict__ datt_w, float* __restrict__ datt_b, int B, int S) {
  constexpr int PER = H / 32;
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= B) return;
  const int l = lane_id();
  float x0[PER], dx0[PER];
#pragma unroll
  for (int e = 0; e < PER; ++e) {
    x0[e] = topic[(long long)b * H + l + 32 * e];
    dx0[e] = 0.f;
  }
  for (int s = 0; s < S; ++s) {
    const long long row = (long long)s * B + b;
    float x1[PER], x2[PER], g[PER], wv[PER], da0 = 0.f, da1 = 0.f, da2 = 0.f;
#pragma unroll
    for (int e = 0; e < PER; ++e) {
      const int c = l + 32 * e;
      x1[e] = img[row * H + c];
      x2[e] = txt[row * H + c];
      g[e] = __bfloat162float(do16[row * H + c]);
      wv[e] = att_w[s * H + c];
      da0 += g[e] * x0[e]; da1 += g[e] * x1[e]; da2 += g[e] * x2[e];
    }
    da0 = warp_sum(da0); da1 = warp_sum(da1); da2 = warp_sum(da2);
    const float a0 = att[row * 3], a1 = att[row * 3 + 1], a2 = att[row * 3 + 2];
    const float dot = a0 * da0 + a1 * da1 + a2 * da2;
    const float de0 = a0 * (da0 - dot), de1 = a1 * (da1 - dot), de2 = a2 * (da2 - dot);
#pragma unroll
    for (int e = 0; e < PER; ++e) {
      const int c = l + 32 * e;
      dx0[e] += a0 * g[e] + de0 * wv[e];
      dimg[row * H + c] = a1 * g[e] + de1 * wv[e];
      dtxt[row * H + c] = a2 * g[e] + de2 * wv[e];
      atomicAdd(datt_w + s * H + c, de0 * x0[e] + de1 * x1[e] + de2 * x2[e]);
    }
    if (l == 0) atomicAdd(datt_b + s, de0 + de1 + de2);
  }
#pragma unroll
  for (int e = 0; e < PER; ++e) dtopic[(long long)b * H + l + 32 * e] = dx0[e];
}

}  // namespace

static int ensure_prior() {
  MMTG_PER_DEVICE_FLAG(g_prior_ready);  // c_prior is per-device __constant__ memory
  if (g_prior_ready) return 0;
  float t[AS * AS];
  for (int i = 0; i < AS; ++i) {
    double pdf[AS], sum = 0.0;
    for (int j = 0; j < AS; ++j) {
      pdf[j] = exp(-0.5 * (double)(j - i) * (double)(j - i)) / sqrt(2.0 * 3.14159265358979323846);
      sum += pdf[j];
    }
    for (int j = 0; j < AS; ++j) t[i * AS + j] = (float)(pdf[j] / sum);
  }
  MMTG_CUDA_OK(cudaMemcpyToSymbol(c_prior, t, sizeof(t)));
  g_prior_ready = true;
  return 0;
}

int pack_sb(const float* in, bf16* out, int B, int S, int D, cudaStream_t st) {
  MMTG_CHECK_ARG(D % 4 == 0, "pack width must be a multiple of 4");
  pack_sb_kernel<<<B * S, 256, 0, st>>>(in, out, B, S, D);
  MMTG_LAUNCH_OK();
  count_launch();
  return 0;
}
int gru_gate_fwd(const float* gi, const float* gh, const float* b_hh, const float* h_prev,
                 float* h_out, bf16* h_out16, float* save, int B, int H, cudaStream_t st) {
  gru_gate_fwd_kernel<<<cdiv(B * H, 256), 256, 0, st>>>(gi, gh, b_hh, h_prev, h_out, h_out16, save, B, H);
  MMTG_LAUNCH_OK();
  count_launch();
  return 0;
}
int gru_gate_bwd(const float* dh_out, const float* dh_carry, const float* save, const float* h_prev,
                 bf16* dgi, bf16* dgh, float* dhz, int B, int H, cudaStream_t st) {
  gru_gate_bwd_kernel<<<cdiv(B * H, 256), 256, 0, st>>>(dh_out, dh_carry, save, h_prev, dgi, dgh, dhz, B, H);
  MMTG_LAUNCH_OK();
  count_launch();
  return 0;
}
int alpha_fwd(const float* qkv, float* ctx, float* probs, float* klpart, int B, int heads, int S,
              int DH, cudaStream_t st) {
  MMTG_CHECK_ARG(S == AS && DH == 128, "alpha attention instantiated for 5 steps x head_dim 128");
  MMTG_TRY(ensure_prior());
  alpha_fwd_kernel<128><<<cdiv(B * heads, 4), 128, 0, st>>>(qkv, ctx, probs, klpart, B, heads);
  MMTG_LAUNCH_OK();
  count_launch();
  return 0;
}
int alpha_bwd(const float* qkv, const float* probs, const float* dctx, const float* g_kl,
              float kl_scale, bf16* dqkv, int B, int heads, int S, int DH, cudaStream_t st) {
  MMTG_CHECK_ARG(S == AS && DH == 128, "alpha attention instantiated for 5 steps x head_dim 128");
  MMTG_TRY(ensure_prior());
  alpha_bwd_kernel<128><<<cdiv(B * heads, 4), 128, 0, st>>>(qkv, probs, dctx, g_kl, kl_scale, dqkv, B, heads);
  MMTG_LAUNCH_OK();
  count_launch();
  return 0;
}
int beta_fwd(const float* topic, const float* img, const float* txt, const float* att_w,
             const float* att_b, bf16* o16, float* att, int B, int S, int H, cudaStream_t st) {
  MMTG_CHECK_ARG(H == 512, "beta gate instantiated for hidden 512");
  beta_fwd_kernel<512><<<cdiv(B, 4), 128, 0, st>>>(topic, img, txt, att_w, att_b, o16, att, B, S);
  MMTG_LAUNCH_OK();
  count_launch();
  return 0;
}
int beta_bwd(const float* topic, const float* img, const float* txt, const float* att_w,
             const float* att, const bf16* do16, float* dtopic, float* dimg, float* dtxt,
             float* datt_w, float* datt_b, int B, int S, int H, cudaStream_t st) {
  MMTG_CHECK_ARG(H == 512, "beta gate instantiated for hidden 512");
  beta_bwd_kernel<512><<<cdiv(B, 4), 128, 0, st>>>(topic, img, txt, att_w, att, do16, dtopic, dimg, dtxt, datt_w, datt_b, B, S);
  MMTG_LAUNCH_OK();
  count_launch();
  return 0;
}

}  // namespace mmtg
