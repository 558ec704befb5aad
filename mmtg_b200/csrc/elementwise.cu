// HBM-bound row kernels of the MMTG hot path: LayerNorm fwd/bwd, fp32->bf16 casts, column sums
// (bias gradients), and the decoder embedding build (token -> WenLan gather + fused-context add)
// with its backward. Every kernel is vectorised and coalesced; grids are sized in multiples of
// the SM count where the loop is grid-strided.
//
// Replaces: torch LayerNorm (src/model.py:380-382, HF GPT2Block ln_1/ln_2/ln_f), the Python
// per-token embedding loops of GPT2_Decoder.forward (src/model.py:253-268) and the ATen
// reductions autograd uses for bias gradients.
#include "../../include/mmtg_b200.h"
#include "ops.h"

namespace mmtg {

void count_launch(int n = 1);

namespace {

// ------------------------------------------------------------------------------------------
// LayerNorm forward: one warp per row, row kept in registers (two-pass mean / variance).
// ------------------------------------------------------------------------------------------
template <int E>
__global__ void __launch_bounds__(256)
ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
              const float* __restrict__ beta, bf16* __restrict__ y16, float* __restrict__ y32,
              float* __restrict__ mean_out, float* __restrict__ rstd_out, int M, float eps) {
  constexpr int V = E / 128;  // float4 per lane
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int l = lane_id();
  for (int row = warp; row < M; row += nwarps) {
    const float4* xr = reinterpret_cast<const float4*>(x + (long long)row * E);
    float4 v[V];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      v[i] = __ldg(xr + l + i * 32);
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
    const float mean = warp_sum(s) * (1.f / E);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += a * a + b * b + c * c + d * d;
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.f / E) + eps);
    if (l == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int c4 = l + i * 32;
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + c4);
      float4 o;
      o.x = (v[i].x - mean) * rstd * g.x + b.x;
      o.y = (v[i].y - mean) * rstd * g.y + b.y;
      o.z = (v[i].z - mean) * rstd * g.z + b.z;
      o.w = (v[i].w - mean) * rstd * g.w + b.w;
      if (y32) reinterpret_cast<float4*>(y32 + (long long)row * E)[c4] = o;
      if (y16) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
        reinterpret_cast<uint2*>(y16 + (long long)row * E)[c4] =
            make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// LayerNorm backward: dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma.
// dx is written or accumulated (residual-stream gradient); dgamma/dbeta are reduced per block
// in shared memory and added to global with one atomic per column per block.
// ------------------------------------------------------------------------------------------
template <int E, bool DY_BF16>
__global__ void __launch_bounds__(E / 2)
ln_bwd_kernel(const void* __restrict__ dy_, const float* __restrict__ x,
              const float* __restrict__ mean, const float* __restrict__ rstd,
              const float* __restrict__ gamma, float* __restrict__ dx, int accumulate_dx,
              float* __restrict__ dgamma, float* __restrict__ dbeta, bf16* __restrict__ dx16,
              float* __restrict__ dx_colsum, int M, const unsigned long long* __restrict__ drop_seed,
              uint32_t drop_site, float drop_p, int drop_dx32) {
  // Two passes over a tile of TR rows (second pass re-reads from L1/L2):
  //   1. warp per row: m1 = mean(g), m2 = mean(g * xhat), g = dy * gamma   -> shared memory
  //   2. thread per column pair: dx for every row of the tile, column sums in 6 registers
  // (row-wise accumulators cost 72 registers / 8 warps per SM; shared atomics were LSU-bound.)
  constexpr int NT = E / 2, NW = NT / 32, TR = 2 * NW, V = E / 128;
  __shared__ float s_m1[TR], s_m2[TR], s_mu[TR], s_rs[TR];
  const int warp = threadIdx.x >> 5, l = lane_id();
  const int c = threadIdx.x * 2;
  const float2 gam2 = *reinterpret_cast<const float2*>(gamma + c);
  float2 adg = make_float2(0.f, 0.f), adb = adg, acs = adg;
  // dx16 / dx_colsum feed the backward of a residual branch whose forward output went through
  // dropout: they carry dx * mask / (1 - p); the fp32 dx (residual stream) stays unmasked unless
  // drop_dx32 (embedding dropout: nothing below needs the unmasked gradient)
  const bool dropping = drop_p > 0.f;
  DropKey dk = {0u, 0u, 0u, 1.f};
  if (dropping) dk = drop_key(drop_seed, drop_site, drop_p);
  // each block owns ONE contiguous row range (balanced single wave), walked in tiles of TR rows
  const int rows_per_block = cdiv(M, (int)gridDim.x);
  const int r_begin = blockIdx.x * rows_per_block, r_end = min(M, r_begin + rows_per_block);
  for (int r0 = r_begin; r0 < r_end; r0 += TR) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int rl = warp * 2 + k, row = r0 + rl;
      if (row < r_end) {
        const float mu = mean[row], rs = rstd[row];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < V; ++i) {
          const int c4 = l + i * 32;
          const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (long long)row * E) + c4);
          float4 dyv;
          if (DY_BF16) {
            const uint2 u = __ldg(reinterpret_cast<const uint2*>((const bf16*)dy_ + (long long)row * E) + c4);
            const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
            const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
            dyv = make_float4(a.x, a.y, b.x, b.y);
          } else {
            dyv = __ldg(reinterpret_cast<const float4*>((const float*)dy_ + (long long)row * E) + c4);
          }
          const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
          const float g0 = dyv.x * gm.x, g1 = dyv.y * gm.y, g2 = dyv.z * gm.z, g3 = dyv.w * gm.w;
          s1 += g0 + g1 + g2 + g3;
          s2 += g0 * (xv.x - mu) + g1 * (xv.y - mu) + g2 * (xv.z - mu) + g3 * (xv.w - mu);
        }
        s1 = warp_sum(s1);
        s2 = warp_sum(s2) * rs;
        if (l == 0) {
          s_m1[rl] = s1 * (1.f / E);
          s_m2[rl] = s2 * (1.f / E);
          s_mu[rl] = mu;
          s_rs[rl] = rs;
        }
      }
    }
    __syncthreads();
    const int nr = min(TR, r_end - r0);
#pragma unroll 4
    for (int rl = 0; rl < nr; ++rl) {
      const long long off = (long long)(r0 + rl) * E + c;
      float2 dyv;
      if (DY_BF16) dyv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>((const bf16*)dy_ + off));
      else dyv = *reinterpret_cast<const float2*>((const float*)dy_ + off);
      const float2 xv = *reinterpret_cast<const float2*>(x + off);
      const float mu = s_mu[rl], rs = s_rs[rl], m1 = s_m1[rl], m2 = s_m2[rl];
      const float xh0 = (xv.x - mu) * rs, xh1 = (xv.y - mu) * rs;
      float2 o;
      o.x = rs * (dyv.x * gam2.x - m1 - xh0 * m2);
      o.y = rs * (dyv.y * gam2.y - m1 - xh1 * m2);
      if (accumulate_dx) {
        const float2 old = *reinterpret_cast<const float2*>(dx + off);
        o.x += old.x;
        o.y += old.y;
      }
      float2 om = o;
      if (dropping) {
        const uint32_t bits = drop_bits(dk, (uint32_t)(off >> 1));
        om.x = drop_keep_lo(dk, bits) ? o.x * dk.inv_keep : 0.f;
        om.y = drop_keep_hi(dk, bits) ? o.y * dk.inv_keep : 0.f;
      }
      *reinterpret_cast<float2*>(dx + off) = drop_dx32 ? om : o;
      if (dx16) *reinterpret_cast<__nv_bfloat162*>(dx16 + off) = __floats2bfloat162_rn(om.x, om.y);
      adg.x += dyv.x * xh0; adg.y += dyv.y * xh1;
      adb.x += dyv.x; adb.y += dyv.y;
      acs.x += om.x; acs.y += om.y;
    }
    __syncthreads();
  }
  if (dgamma) { atomicAdd(dgamma + c, adg.x); atomicAdd(dgamma + c + 1, adg.y); }
  if (dbeta) { atomicAdd(dbeta + c, adb.x); atomicAdd(dbeta + c + 1, adb.y); }
  if (dx_colsum) { atomicAdd(dx_colsum + c, acs.x); atomicAdd(dx_colsum + c + 1, acs.y); }
}

// ------------------------------------------------------------------------------------------
// LayerNorm backward of the decoder's residual stream, split in two (E = 768, bf16 dy):
//  * ln_bwd_rows_kernel - the part the backward CHAIN waits for: one warp per row, the row lives in
//    registers (x, dy, the running residual gradient: 18 independent 8/16-byte loads per lane are
//    in flight before the first use), dx += rstd (g - mean(g) - xhat mean(g xhat)) and its masked
//    bf16 copy are written once. No column accumulators, no atomics: 16 bytes per element.
//  * ln_param_grad_kernel - dgamma, dbeta and the column sums of the bf16 dx copy (the bias
//    gradient of the c_proj that feeds this LayerNorm's input): column-oriented partial sums
//    over a row range. Only the optimizer consumes them, so the engine runs this kernel on the
//    weight-gradient side stream.
// ------------------------------------------------------------------------------------------
template <int E>
__global__ void __launch_bounds__(128)
ln_bwd_rows_kernel(const bf16* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
                   const float* __restrict__ rstd, const float* __restrict__ gamma, float* __restrict__ dx,
                   int accumulate_dx, bf16* __restrict__ dx16, int M,
                   const unsigned long long* __restrict__ drop_seed, uint32_t drop_site, float drop_p, int drop_dx32) {
  constexpr int V = E / 128;  // float4 per lane
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= M) return;
  const int l = lane_id();
  const long long base = (long long)row * E;
  float4 xv[V], old[V];
  uint2 dyv[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int c4 = l + i * 32;
    xv[i] = __ldg(reinterpret_cast<const float4*>(x + base) + c4);
    dyv[i] = __ldg(reinterpret_cast<const uint2*>(dy + base) + c4);
    if (accumulate_dx) old[i] = *(reinterpret_cast<const float4*>(dx + base) + c4);
  }
  const float mu = mean[row], rs = rstd[row];
  const bool dropping = drop_p > 0.f;
  DropKey dk = {0u, 0u, 0u, 1.f};
  if (dropping) dk = drop_key(drop_seed, drop_site, drop_p);
  float4 g[V];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + l + i * 32);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&dyv[i].x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&dyv[i].y));
    g[i] = make_float4(a.x * gm.x, a.y * gm.y, b.x * gm.z, b.y * gm.w);
    xv[i] = make_float4((xv[i].x - mu) * rs, (xv[i].y - mu) * rs, (xv[i].z - mu) * rs, (xv[i].w - mu) * rs);  // xhat
    s1 += g[i].x + g[i].y + g[i].z + g[i].w;
    s2 += g[i].x * xv[i].x + g[i].y * xv[i].y + g[i].z * xv[i].z + g[i].w * xv[i].w;
  }
  const float m1 = warp_sum(s1) * (1.f / E), m2 = warp_sum(s2) * (1.f / E);
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int c4 = l + i * 32;
    float4 o;
    o.x = rs * (g[i].x - m1 - xv[i].x * m2);
    o.y = rs * (g[i].y - m1 - xv[i].y * m2);
    o.z = rs * (g[i].z - m1 - xv[i].z * m2);
    o.w = rs * (g[i].w - m1 - xv[i].w * m2);
    if (accumulate_dx) {
      o.x += old[i].x; o.y += old[i].y; o.z += old[i].z; o.w += old[i].w;
    }
    float4 om = o;
    if (dropping) {  // dx16 carries dx * mask / (1 - p) (see ln_bwd_kernel)
      const uint32_t pair = (uint32_t)((base + c4 * 4) >> 1);
      const uint32_t b0 = drop_bits(dk, pair), b1 = drop_bits(dk, pair + 1);
      om.x = drop_keep_lo(dk, b0) ? o.x * dk.inv_keep : 0.f;
      om.y = drop_keep_hi(dk, b0) ? o.y * dk.inv_keep : 0.f;
      om.z = drop_keep_lo(dk, b1) ? o.z * dk.inv_keep : 0.f;
      om.w = drop_keep_hi(dk, b1) ? o.w * dk.inv_keep : 0.f;
    }
    *(reinterpret_cast<float4*>(dx + base) + c4) = drop_dx32 ? om : o;
    if (dx16) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(om.x, om.y), hi = __floats2bfloat162_rn(om.z, om.w);
      *(reinterpret_cast<uint2*>(dx16 + base) + c4) =
          make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
    }
  }
}

// grid = (E / 128, row splits); each warp owns 128 columns (4 per lane) of a row subset
__global__ void __launch_bounds__(256)
ln_param_grad_kernel(const bf16* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
                     const float* __restrict__ rstd, const bf16* __restrict__ dx16, float* __restrict__ dgamma,
                     float* __restrict__ dbeta, float* __restrict__ dx_colsum, int M, int E) {
  __shared__ float4 part[3][8][32];
  const int warp = threadIdx.x >> 5, l = lane_id();
  const int col = blockIdx.x * 128 + l * 4;
  const int rows_per = cdiv(M, (int)gridDim.y);
  const int r0 = blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
  float4 adg = make_float4(0.f, 0.f, 0.f, 0.f), adb = adg, acs = adg;
#pragma unroll 4
  for (int r = r0 + warp; r < r1; r += 8) {
    const long long off = (long long)r * E + col;
    const uint2 du = __ldg(reinterpret_cast<const uint2*>(dy + off));
    const float4 xv = __ldg(reinterpret_cast<const float4*>(x + off));
    uint2 cu = make_uint2(0u, 0u);
    if (dx16) cu = __ldg(reinterpret_cast<const uint2*>(dx16 + off));
    const float mu = __ldg(mean + r), rs = __ldg(rstd + r);
    const float2 d0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&du.x));
    const float2 d1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&du.y));
    const float2 c0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&cu.x));
    const float2 c1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&cu.y));
    adg.x += d0.x * (xv.x - mu) * rs; adg.y += d0.y * (xv.y - mu) * rs;
    adg.z += d1.x * (xv.z - mu) * rs; adg.w += d1.y * (xv.w - mu) * rs;
    adb.x += d0.x; adb.y += d0.y; adb.z += d1.x; adb.w += d1.y;
    acs.x += c0.x; acs.y += c0.y; acs.z += c1.x; acs.w += c1.y;
  }
  part[0][warp][l] = adg;
  part[1][warp][l] = adb;
  part[2][warp][l] = acs;
  __syncthreads();
  if (warp < 3) {
    float4 t = part[warp][0][l];
#pragma unroll
    for (int w = 1; w < 8; ++w) {
      const float4 u = part[warp][w][l];
      t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
    }
    float* out = warp == 0 ? dgamma : (warp == 1 ? dbeta : dx_colsum);
    if (out) {
      atomicAdd(out + col, t.x);
      atomicAdd(out + col + 1, t.y);
      atomicAdd(out + col + 2, t.z);
      atomicAdd(out + col + 3, t.w);
    }
  }
}

// ------------------------------------------------------------------------------------------
// fp32 -> bf16 cast (weights, activations), 8 elements per thread.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
cast_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x * 8;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < n; i += stride) {
    if (i + 8 <= n) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(src + i));
      const float4 b = __ldg(reinterpret_cast<const float4*>(src + i + 4));
      __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
      __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
      *reinterpret_cast<uint4*>(dst + i) =
          make_uint4(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1),
                     *reinterpret_cast<uint32_t*>(&p2), *reinterpret_cast<uint32_t*>(&p3));
    } else {
      for (long long j = i; j < n; ++j) dst[j] = __float2bfloat16(src[j]);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Column sums of a [M, N] matrix (bf16 or fp32 in), optionally emitting a bf16 copy.
// grid = (ceil(N/64), row_splits); each warp owns 64 columns (2 per lane) of a row subset.
// ------------------------------------------------------------------------------------------
template <bool IN_BF16>
__global__ void __launch_bounds__(256)
colsum_kernel(const void* __restrict__ x_, long long ld, bf16* __restrict__ copy16, long long ldc,
              float* __restrict__ out, int M, int N) {
  __shared__ float2 part[8][32];
  const int warp = threadIdx.x >> 5, l = lane_id();
  const int col = blockIdx.x * 64 + l * 2;
  const int rows_per = cdiv(M, (int)gridDim.y);
  const int r0 = blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
  float2 acc = make_float2(0.f, 0.f);
  if (col < N) {
    for (int r = r0 + warp; r < r1; r += 8) {
      float2 v;
      if (IN_BF16) {
        v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>((const bf16*)x_ + (long long)r * ld + col));
      } else {
        v = *reinterpret_cast<const float2*>((const float*)x_ + (long long)r * ld + col);
        if (copy16)
          *reinterpret_cast<__nv_bfloat162*>(copy16 + (long long)r * ldc + col) = __floats2bfloat162_rn(v.x, v.y);
      }
      acc.x += v.x;
      acc.y += v.y;
    }
  }
  part[warp][l] = acc;
  __syncthreads();
  if (warp == 0 && col < N) {
    float2 s = part[0][l];
#pragma unroll
    for (int w = 1; w < 8; ++w) {
      s.x += part[w][l].x;
      s.y += part[w][l].y;
    }
    atomicAdd(out + col, s.x);
    atomicAdd(out + col + 1, s.y);
  }
}

// ------------------------------------------------------------------------------------------
// Decoder embedding build (src/model.py:253-277): E[b,p] = table[id(b,p)] (+ ctx[b, j/two_sent]
// for lyric position j = p - P < S*two_sent), cast to bf16 as the projector GEMM's A operand.
// ctx rows are ordered (s, b): row = s*B + b. One block per (b, p) row, 8 elements per thread.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
embed_fwd_kernel(const float* __restrict__ table, const int* __restrict__ topic_ids,
                 const int* __restrict__ input_ids, const float* __restrict__ ctx,
                 bf16* __restrict__ out, int B, int P, int T, int S, int two_sent, int D, int table_rows) {
  const int row = blockIdx.x;
  const int L = P + T;
  const int b = row / L, p = row - b * L;
  int id, k = -1;
  if (p < P) {
    id = topic_ids[b * P + p];
  } else {
    const int j = p - P;
    id = input_ids[b * T + j];
    if (j / two_sent < S) k = j / two_sent;
  }
  if ((unsigned)id >= (unsigned)table_rows) {  // the reference's dict lookup raises KeyError here
    if (threadIdx.x == 0)
      printf("mmtg: token id %d at (row %d, position %d) is outside the token table [0, %d)\n", id, b, p, table_rows);
    __trap();
  }
  const float* trow = table + (long long)id * D;
  const float* crow = k >= 0 ? ctx + ((long long)k * B + b) * D : nullptr;
  for (int c = threadIdx.x * 8; c < D; c += blockDim.x * 8) {
    float4 a = __ldg(reinterpret_cast<const float4*>(trow + c));
    float4 d = __ldg(reinterpret_cast<const float4*>(trow + c + 4));
    if (crow) {
      const float4 e = __ldg(reinterpret_cast<const float4*>(crow + c));
      const float4 f = __ldg(reinterpret_cast<const float4*>(crow + c + 4));
      a.x += e.x; a.y += e.y; a.z += e.z; a.w += e.w;
      d.x += f.x; d.y += f.y; d.z += f.z; d.w += f.w;
    }
    __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(d.x, d.y), p3 = __floats2bfloat162_rn(d.z, d.w);
    *reinterpret_cast<uint4*>(out + (long long)row * D + c) =
        make_uint4(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1),
                   *reinterpret_cast<uint32_t*>(&p2), *reinterpret_cast<uint32_t*>(&p3));
  }
}

// Backward of the context add: dctx[s*B+b, :] = sum over the two_sent lyric rows of pair s of
// dE[b, P + j, :]. The token table is frozen (not a parameter), so nothing else flows.
// grid = (B*S, D/256); one thread per column pair.
__global__ void __launch_bounds__(128)
embed_bwd_kernel(const bf16* __restrict__ dE, bf16* __restrict__ dctx16, float* __restrict__ dctx32,
                 int B, int P, int T, int S, int two_sent, int D) {
  const int bs = blockIdx.x;
  const int b = bs / S, s = bs - b * S;
  const int c = (blockIdx.y * blockDim.x + threadIdx.x) * 2;
  if (c >= D) return;
  const int L = P + T;
  const int j0 = s * two_sent, j1 = min(T, j0 + two_sent);
  float2 acc = make_float2(0.f, 0.f);
  for (int j = j0; j < j1; ++j) {
    const float2 v = __bfloat1622float2(
        *reinterpret_cast<const __nv_bfloat162*>(dE + ((long long)b * L + P + j) * D + c));
    acc.x += v.x;
    acc.y += v.y;
  }
  const long long o = ((long long)s * B + b) * D + c;
  if (dctx16) *reinterpret_cast<__nv_bfloat162*>(dctx16 + o) = __floats2bfloat162_rn(acc.x, acc.y);
  if (dctx32) *reinterpret_cast<float2*>(dctx32 + o) = acc;
}

// Backward of "h0 = proj + wpe[pos] + wte[type]" (HF modeling_gpt2.py:579-612):
// dwpe[p, :] += sum_b dh[b*L + p, :]  — one thread per (p, column), deterministic.
__global__ void __launch_bounds__(256)
posadd_bwd_kernel(const float* __restrict__ dh, float* __restrict__ dwpe, int B, int L, int E) {
  const int p = blockIdx.x;
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= E) return;
  float a = 0.f;
  for (int b = 0; b < B; ++b) a += dh[((long long)b * L + p) * E + c];
  dwpe[(long long)p * E + c] += a;
}
// dwte[type_ids[row], :] += dh[row, :]. Only a handful of distinct type ids exist, so each block
// first reduces its 32 rows per type in shared memory, then issues one atomic per (type, col).
constexpr int TYPE_SLOTS = 16;
__global__ void __launch_bounds__(128)
typeadd_bwd_kernel(const float* __restrict__ dh, const int* __restrict__ type_ids,
                   float* __restrict__ dwte, int M, int E) {
  __shared__ float acc[TYPE_SLOTS][128];
  __shared__ int touched[TYPE_SLOTS];
  const int c = blockIdx.y * 128 + threadIdx.x;
  for (int t = 0; t < TYPE_SLOTS; ++t) acc[t][threadIdx.x] = 0.f;
  if (threadIdx.x < TYPE_SLOTS) touched[threadIdx.x] = 0;
  __syncthreads();
  // 32 rows per block, fetched 8 at a time (independent loads in flight) before the shared-memory
  // accumulation: the 128-row serial walk of round 1 was latency-bound (67 us for 23 MB)
  const int r0 = blockIdx.x * 32, r1 = min(M, r0 + 32);
  for (int rb = r0; rb < r1; rb += 8) {
    float v[8];
    int ty[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int r = rb + k;
      ty[k] = r < r1 ? type_ids[r] : -1;
      v[k] = (r < r1 && c < E) ? dh[(long long)r * E + c] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int t = ty[k];
      if (t < 0) continue;
      if (t < TYPE_SLOTS) {
        acc[t][threadIdx.x] += v[k];
        if (threadIdx.x == 0) touched[t] = 1;
      } else if (c < E) {
        atomicAdd(dwte + (long long)t * E + c, v[k]);
      }
    }
  }
  __syncthreads();
  if (c < E)
    for (int t = 0; t < TYPE_SLOTS; ++t)
      if (touched[t]) atomicAdd(dwte + (long long)t * E + c, acc[t][threadIdx.x]);
}
// fp32 [M, V] contiguous gradient -> bf16 [M, Vp] padded GEMM operand (generic MyLoss path)
__global__ void __launch_bounds__(256)
dlogits_cvt_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int V, int Vp) {
  const long long row = blockIdx.x;
  for (int c = threadIdx.x; c < Vp; c += 256)
    dst[row * Vp + c] = __float2bfloat16(c < V ? __ldg(src + row * V + c) : 0.f);
}

}  // namespace

int layernorm_fwd(const float* x, const float* gamma, const float* beta, bf16* y16, float* y32,
                  float* mean, float* rstd, int M, int E, float eps, cudaStream_t st) {
  MMTG_CHECK_ARG(E == 768 || E == 512, "LayerNorm width %d not instantiated (512, 768)", E);
  const int blocks = min(cdiv(M, 8), num_sms() * 8);
  ProfScope prof(2, 0, (double)M * E * (4 + (y16 ? 2 : 0) + (y32 ? 4 : 0)), st);
  if (E == 768) ln_fwd_kernel<768><<<blocks, 256, 0, st>>>(x, gamma, beta, y16, y32, mean, rstd, M, eps);
  else ln_fwd_kernel<512><<<blocks, 256, 0, st>>>(x, gamma, beta, y16, y32, mean, rstd, M, eps);
  MMTG_LAUNCH_OK();
  count_launch();
  return 0;
}

int layernorm_bwd(const void* dy, int dy_bf16, const float* x, const float* mean, const float* rstd,
                  const float* gamma, float* dx, int accumulate_dx, float* dgamma, float* dbeta,
                  bf16* dx16, float* dx_colsum, int M, int E, cudaStream_t st, const DropSpec* drop) {
  const unsigned long long* dseed = drop ? drop->seed : nullptr;
  const uint32_t dsite = drop ? drop->site : 0u;
  const float dp = (drop && drop->seed) ? drop->p : 0.f;
  const int ddx32 = drop ? drop->mask_dx32 : 0;
  MMTG_CHECK_ARG(E == 768 || E == 512, "LayerNorm width %d not instantiated (512, 768)", E);
  if (E == 768 && dy_bf16 && !dgamma && !dbeta && !dx_colsum) {
    // chain part only (the parameter gradients come from ln_param_grads on the side stream)
    ProfScope prof(2, 0, (double)M * E * (4 + 2 + 4 + (accumulate_dx ? 4 : 0) + (dx16 ? 2 : 0)), st);
    ln_bwd_rows_kernel<768><<<cdiv(M, 4), 128, 0, st>>>((const bf16*)dy, x, mean, rstd, gamma, dx, accumulate_dx, dx16, M,
                                                        dseed, dsite, dp, ddx32);
    MMTG_LAUNCH_OK();
    count_launch();
    return 0;
  }
  // 3 blocks of E/2 threads are resident per SM (56 registers): one balanced wave
  const int blocks = max(1, min(M, num_sms() * 3));
  ProfScope prof(2, 0, (double)M * E * (4 + (dy_bf16 ? 2 : 4) + 4 + (accumulate_dx ? 4 : 0) + (dx16 ? 2 : 0)), st);
#define LNB(EE, BF) ln_bwd_kernel<EE, BF><<<blocks, EE / 2, 0, st>>>(dy, x, mean, rstd, gamma, dx, accumulate_dx, dgamma, dbeta, dx16, dx_colsum, M, dseed, dsite, dp, ddx32)
  if (E == 768) { if (dy_bf16) LNB(768, true); else LNB(768, false); }
  else { if (dy_bf16) LNB(512, true); else LNB(512, false); }
#undef LNB
  MMTG_LAUNCH_OK();
  count_launch();
  return 0;
}

int ln_param_grads(const bf16* dy, const float* x, const float* mean, const float* rstd, const bf16* dx16,
                   float* dgamma, float* dbeta, float* dx_colsum, int M, int E, cudaStream_t st) {
  MMTG_CHECK_ARG(E % 128 == 0, "ln_param_grads needs E %% 128 == 0");
  int splits = cdiv(num_sms() * 4, E / 128);
  splits = max(1, min(splits, cdiv(M, 32)));
  dim3 grid(E / 128, splits);
  ProfScope prof(2, 0, (double)M * E * (2 + 4 + (dx16 ? 2 : 0)), st);
  ln_param_grad_kernel<<<grid, 256, 0, st>>>(dy, x, mean, rstd, dx16, dgamma, dbeta, dx_colsum, M, E);
  MMTG_LAUNCH_OK();
  count_launch();
  return 0;
}

int cast_bf16(const float* src, bf16* dst, long long n, cudaStream_t st) {
  if (n <= 0) return 0;
  const long long want = cdivll(n, 8 * 256);
  const int blocks = (int)(want < (long long)num_sms() * 16 ? want : (long long)num_sms() * 16);
  cast_bf16_kernel<<<blocks, 256, 0, st>>>(src, dst, n);
  MMTG_LAUNCH_OK();
  count_launch();
  return 0;
}

int colsum(const void* x, int x_bf16, long long ld, bf16* copy16, long long ldc, float* out, int M,
           int N, cudaStream_t st) {
  MMTG_CHECK_ARG(N % 2 == 0 && ld % 2 == 0, "colsum needs even N and pitch");
  int splits = cdiv(num_sms() * 4, cdiv(N, 64));
  splits = max(1, min(splits, cdiv(M, 32)));
  dim3 grid(cdiv(N, 64), splits);
  ProfScope prof(2, 0, (double)M * N * ((x_bf16 ? 2 : 4) + (copy16 ? 2 : 0)), st);
  if (x_bf16) colsum_kernel<true><<<grid, 256, 0, st>>>(x, ld, nullptr, 0, out, M, N);
  else colsum_kernel<false><<<grid, 256, 0, st>>>(x, ld, copy16, ldc, out, M, N);
  MMTG_LAUNCH_OK();
  count_launch();
  return 0;
}

int embed_fwd(const float* table, const int* topic_ids, const int* input_ids, const float* ctx,
              bf16* out, int B, int P, int T, int S, int two_sent, int D, int table_rows, cudaStream_t st) {
  MMTG_CHECK_ARG(D % 8 == 0, "embedding width must be a multiple of 8");
  MMTG_CHECK_ARG(table_rows > 0, "token table row count missing (mmtg_model.table_rows)");
  ProfScope prof(2, 0, (double)B * (P + T) * D * 6, st);
  embed_fwd_kernel<<<B * (P + T), 256, 0, st>>>(table, topic_ids, input_ids, ctx, out, B, P, T, S, two_sent, D,
                                                table_rows);
  MMTG_LAUNCH_OK();
  count_launch();
  return 0;
}

int embed_bwd(const bf16* dE, bf16* dctx16, float* dctx32, int B, int P, int T, int S, int two_sent,
              int D, cudaStream_t st) {
  dim3 grid(B * S, cdiv(D, 256));
  embed_bwd_kernel<<<grid, 128, 0, st>>>(dE, dctx16, dctx32, B, P, T, S, two_sent, D);
  MMTG_LAUNCH_OK();
  count_launch();
  return 0;
}

int posadd_bwd(const float* dh, float* dwpe, int B, int L, int E, cudaStream_t st) {
  dim3 grid(L, cdiv(E, 256));
  posadd_bwd_kernel<<<grid, 256, 0, st>>>(dh, dwpe, B, L, E);
  MMTG_LAUNCH_OK();
  count_launch();
  return 0;
}
int typeadd_bwd(const float* dh, const int* type_ids, float* dwte, int M, int E, cudaStream_t st) {
  dim3 grid(cdiv(M, 32), cdiv(E, 128));
  typeadd_bwd_kernel<<<grid, 128, 0, st>>>(dh, type_ids, dwte, M, E);
  MMTG_LAUNCH_OK();
  count_launch();
  return 0;
}
int dlogits_f32_to_bf16(const float* src, bf16* dst, int M, int V, int Vp, cudaStream_t st) {
  dlogits_cvt_kernel<<<M, 256, 0, st>>>(src, dst, V, Vp);
  MMTG_LAUNCH_OK();
  count_launch();
  return 0;
}

}  // namespace mmtg

using namespace mmtg;

extern "C" int mmtg_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y_bf16,
                                  float* y_f32, float* mean, float* rstd, int32_t M, int32_t E,
                                  float eps, void* stream) {
  MMTG_CHECK_ARG(x && gamma && beta && (y_bf16 || y_f32) && M > 0, "bad layernorm args");
  return layernorm_fwd(x, gamma, beta, (bf16*)y_bf16, y_f32, mean, rstd, M, E, eps, (cudaStream_t)stream);
}
extern "C" int mmtg_ln_param_grads(const void* dy_bf16, const float* x, const float* mean, const float* rstd,
                                   const void* dx_bf16, float* dgamma, float* dbeta, float* dx_colsum, int32_t M,
                                   int32_t E, void* stream) {
  MMTG_CHECK_ARG(dy_bf16 && x && mean && rstd && M > 0 && (dgamma || dbeta || dx_colsum), "bad ln_param_grads args");
  MMTG_CHECK_ARG(!dx_colsum || dx_bf16, "dx_colsum needs the bf16 dx copy");
  return ln_param_grads((const bf16*)dy_bf16, x, mean, rstd, (const bf16*)dx_bf16, dgamma, dbeta, dx_colsum, M, E,
                        (cudaStream_t)stream);
}
extern "C" int mmtg_layernorm_bwd(const void* dy, int32_t dy_is_bf16, const float* x, const float* mean,
                                  const float* rstd, const float* gamma, float* dx,
                                  int32_t accumulate_dx, float* dgamma, float* dbeta, void* dx_bf16,
                                  float* dx_colsum, int32_t M, int32_t E, void* stream) {
  MMTG_CHECK_ARG(dy && x && mean && rstd && gamma && dx && M > 0, "bad layernorm bwd args");
  return layernorm_bwd(dy, dy_is_bf16, x, mean, rstd, gamma, dx, accumulate_dx, dgamma, dbeta,
                       (bf16*)dx_bf16, dx_colsum, M, E, (cudaStream_t)stream);
}
extern "C" int mmtg_cast_bf16(const float* src, void* dst, int64_t n, void* stream) {
  MMTG_CHECK_ARG(src && dst && ((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0), "cast needs 16-byte aligned buffers");
  return cast_bf16(src, (bf16*)dst, n, (cudaStream_t)stream);
}
extern "C" int mmtg_colsum(const void* x, int32_t x_is_bf16, int64_t ld, void* copy_bf16, int64_t ldc,
                           float* out, int32_t M, int32_t N, void* stream) {
  MMTG_CHECK_ARG(x && out && M > 0 && N > 0, "bad colsum args");
  return colsum(x, x_is_bf16, ld, (bf16*)copy_bf16, ldc, out, M, N, (cudaStream_t)stream);
}
extern "C" int mmtg_embed_fwd(const float* table, const int32_t* topic_ids, const int32_t* input_ids,
                              const float* ctx, void* out_bf16, int32_t B, int32_t P, int32_t T,
                              int32_t S, int32_t two_sent, int32_t D, void* stream) {
  MMTG_CHECK_ARG(table && topic_ids && input_ids && out_bf16, "bad embed args");
  return embed_fwd(table, topic_ids, input_ids, ctx, (bf16*)out_bf16, B, P, T, S, two_sent, D, 0x7fffffff,
                   (cudaStream_t)stream);  // standalone op (tests): the caller vouches for the ids
}
extern "C" int mmtg_embed_bwd(const void* dE_bf16, void* dctx_bf16, float* dctx_f32, int32_t B, int32_t P,
                              int32_t T, int32_t S, int32_t two_sent, int32_t D, void* stream) {
  MMTG_CHECK_ARG(dE_bf16 && (dctx_bf16 || dctx_f32), "bad embed bwd args");
  return embed_bwd((const bf16*)dE_bf16, (bf16*)dctx_bf16, dctx_f32, B, P, T, S, two_sent, D, (cudaStream_t)stream);
}

// ---- dropout helpers (tests / step seed) ----
namespace mmtg {
namespace {
__global__ void __launch_bounds__(256)
dropout_mask_kernel(const unsigned long long* __restrict__ seed, uint32_t site, float p, long long n,
                    uint8_t* __restrict__ keep) {
  const DropKey dk = drop_key(seed, site, p);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long pr = (long long)blockIdx.x * blockDim.x + threadIdx.x; pr * 2 < n; pr += stride) {
    const uint32_t bits = drop_bits(dk, (uint32_t)pr);
    keep[pr * 2] = drop_keep_lo(dk, bits) ? 1 : 0;
    if (pr * 2 + 1 < n) keep[pr * 2 + 1] = drop_keep_hi(dk, bits) ? 1 : 0;
  }
}
__global__ void seed_next_kernel(unsigned long long* seed) {
  // splitmix64 step: consecutive steps get unrelated keys
  unsigned long long z = (*seed += 0x9E3779B97F4A7C15ull);
  (void)z;
}
}  // namespace
}  // namespace mmtg

extern "C" int mmtg_dropout_mask(const uint64_t* seed_dev, uint32_t site, float p, int64_t n, uint8_t* keep_out,
                                 void* stream) {
  using namespace mmtg;
  MMTG_CHECK_ARG(seed_dev && keep_out && n > 0 && p >= 0.f && p < 1.f, "bad dropout_mask args");
  const int blocks = (int)((n / 2 + 255) / 256 < 4096 ? (n / 2 + 255) / 256 + 1 : 4096);
  dropout_mask_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const unsigned long long*)seed_dev, site, p,
                                                                (long long)n, keep_out);
  MMTG_LAUNCH_OK();
  count_launch();
  return 0;
}

extern "C" int mmtg_dropout_next_seed(uint64_t* seed_dev, void* stream) {
  using namespace mmtg;
  MMTG_CHECK_ARG(seed_dev, "null seed");
  seed_next_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((unsigned long long*)seed_dev);
  MMTG_LAUNCH_OK();
  count_launch();
  return 0;
}
