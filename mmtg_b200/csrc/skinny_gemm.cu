// Skinny GEMM for generation: y[M <= 64, N] = epilogue(x[M, K] · W) with bf16 operands.
// At decode time every dense layer is a weight-streaming (HBM-bound) problem with at most 64
// activation rows; the persistent tcgen05 kernel's fixed cost (cluster launch, TMEM allocation,
// 128-row tiles) dominated there (~10 us per launch). This kernel: one CTA per 64 output
// columns, 4 warps x 16 rows, K streamed through a double-buffered cp.async pipeline of
// 64x64 swizzled tiles, mma.sync m16n8k16, fused bias / tanh / gelu_new / residual /
// row-gather adds (wpe[pos] + wte[type]) epilogue. Both weight layouts are consumed in place:
// HF Conv1D [K, N] (ldmatrix.trans) and nn.Linear / tied wte [N, K].
#include "../../include/mmtg_b200.h"
#include "common.cuh"
#include "mma_tiles.cuh"

namespace mmtg {

void count_launch(int n = 1);

namespace {

struct SkinnyParams {
  const bf16* x;   // [M, K]
  const bf16* w;   // [K, N] (w_kn = 1) or [N, K]
  long long ldx, ldw;
  int M, N, K, w_kn;
  void* out;
  long long ldo;
  int out_bf16, act;
  const float* bias;
  const float* residual;
  long long ldr;
  const float* rowtab0;
  const int* rowidx0;
  long long ldt0;
  const float* rowtab1;
  const int* rowidx1;
  long long ldt1;
};

__global__ void __launch_bounds__(ATT_THREADS)
skinny_gemm_kernel(const SkinnyParams p) {
  __shared__ __align__(128) bf16 sX[2][64 * 64];
  __shared__ __align__(128) bf16 sW[2][64 * 64];
  const int n0 = blockIdx.x * 64;
  const int warp = threadIdx.x >> 5, l = lane_id();
  const int nk = cdiv(p.K, 64);
  auto issue = [&](int kb, int buf) {
    // x tile: rows = tokens, cols = k (zero-fill beyond M; K is a multiple of 64 on this path)
    load_tile_async(sX[buf], p.x, p.ldx, 0, p.M, kb * 64);
    if (p.w_kn) {
      // W rows = k, cols = n: tile [64 k][64 n]; columns beyond N are handled by clamping the row
      // count to K and masking the store (N is a multiple of 8 for every [K, N] weight here)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = threadIdx.x + i * ATT_THREADS;
        const int r = idx >> 3, c = idx & 7;
        const bool ok = (kb * 64 + r) < p.K && (n0 + c * 8) < p.N;
        const bf16* src = p.w + (long long)(ok ? kb * 64 + r : 0) * p.ldw + (ok ? n0 + c * 8 : 0);
        cp_async16(reinterpret_cast<uint8_t*>(sW[buf]) + r * 128 + ((c ^ (r & 7)) << 4), src, ok);
      }
    } else {
      // W rows = n, cols = k: tile [64 n][64 k]; rows beyond N zero-filled
      load_tile_async(sW[buf], p.w + (long long)n0 * p.ldw, p.ldw, 0, p.N - n0, kb * 64);
    }
  };
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  issue(0, 0);
  cp_async_commit();
  for (int kb = 0; kb < nk; ++kb) {
    const int buf = kb & 1;
    if (kb + 1 < nk) issue(kb + 1, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    uint32_t a[4][4];
    load_a_frags(sX[buf], warp * 16, a);
    if (p.w_kn) mma_nn_a(acc, a, sW[buf]);
    else mma_nt(acc, a, sW[buf]);
    __syncthreads();
  }
  // epilogue: thread holds rows (warp*16 + l/4, +8), columns nb*8 + 2*(l%4) + {0,1}
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = warp * 16 + (l >> 2) + r * 8;
    if (row >= p.M) continue;
    long long t0 = 0, t1 = 0;
    if (p.rowtab0) t0 = (long long)p.rowidx0[row] * p.ldt0;
    if (p.rowtab1) t1 = (long long)p.rowidx1[row] * p.ldt1;
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      const int col = n0 + nb * 8 + (l & 3) * 2;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int cc = col + e;
        if (cc >= p.N) continue;
        float v = acc[nb][2 * r + e];
        if (p.bias) v += __ldg(p.bias + cc);
        if (p.act == MMTG_ACT_TANH) v = tanhf(v);
        else if (p.act == MMTG_ACT_GELU_NEW) v = gelu_new_f(v);
        if (p.residual) v += __ldg(p.residual + (long long)row * p.ldr + cc);
        if (p.rowtab0) v += __ldg(p.rowtab0 + t0 + cc);
        if (p.rowtab1) v += __ldg(p.rowtab1 + t1 + cc);
        if (p.out_bf16) ((bf16*)p.out)[(long long)row * p.ldo + cc] = __float2bfloat16(v);
        else ((float*)p.out)[(long long)row * p.ldo + cc] = v;
      }
    }
  }
}

}  // namespace

// Uses the mmtg_gemm_args struct (subset): A = x [M,K] K-major, B = W; b_mn_major = 1 means W is
// stored [K, N]. Supported epilogue: bias, act, residual, rowtab0(rowidx0)/rowtab1, fp32/bf16 out.
int skinny_gemm(const mmtg_gemm_args* a, cudaStream_t st) {
  MMTG_CHECK_ARG(a->M > 0 && a->M <= 64, "skinny GEMM handles at most 64 rows (got %d)", a->M);
  MMTG_CHECK_ARG(!a->a_mn_major && !a->out2 && !a->dgelu_src && !a->colsum && !a->lse_partial &&
                     !a->accumulate && a->split_k <= 1,
                 "unsupported epilogue for the skinny GEMM");
  MMTG_CHECK_ARG(a->K % 8 == 0 && a->lda % 8 == 0 && a->ldb % 8 == 0, "skinny GEMM needs 16-byte aligned rows");
  MMTG_CHECK_ARG(!(a->b_mn_major && a->N % 8 != 0), "[K,N] weights need N %% 8 == 0");
  MMTG_CHECK_ARG(!(a->rowtab0 && !a->rowidx0), "skinny GEMM rowtab0 needs rowidx0");
  SkinnyParams p;
  p.x = (const bf16*)a->A; p.w = (const bf16*)a->B; p.ldx = a->lda; p.ldw = a->ldb;
  p.M = a->M; p.N = a->N; p.K = a->K; p.w_kn = a->b_mn_major ? 1 : 0;
  p.out = a->out; p.ldo = a->ldo; p.out_bf16 = a->out_dtype == MMTG_BF16; p.act = a->act;
  p.bias = a->bias; p.residual = a->residual; p.ldr = a->ldr;
  p.rowtab0 = a->rowtab0; p.rowidx0 = a->rowidx0; p.ldt0 = a->ldt0;
  p.rowtab1 = a->rowtab1; p.rowidx1 = a->rowidx1; p.ldt1 = a->ldt1;
  ProfScope prof(0, 2.0 * a->M * a->N * a->K, 2.0 * a->N * a->K, st);
  skinny_gemm_kernel<<<cdiv(a->N, 64), ATT_THREADS, 0, st>>>(p);
  MMTG_LAUNCH_OK();
  count_launch();
  return 0;
}

}  // namespace mmtg

extern "C" int mmtg_skinny_gemm_bf16(const mmtg_gemm_args* args, void* stream) {
  MMTG_CHECK_ARG(args && args->A && args->B && args->out, "null args");
  return mmtg::skinny_gemm(args, (cudaStream_t)stream);
}
