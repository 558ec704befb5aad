// Skinny GEMM for generation: y[M <= 64, N] = epilogue(x[M, K] · W) with bf16 operands.
// At decode time every dense layer is a weight-streaming (HBM-bound) problem with at most 64
// activation rows; the persistent tcgen05 kernel's fixed cost (cluster launch, TMEM allocation,
// 128-row tiles) dominated there (~10 us per launch). This kernel: split-K over a thread-block
// cluster (see below), 4 warps x 16 rows per CTA, 64x64 swizzled tiles, mma.sync m16n8k16,
// DSMEM reduction, fused bias / tanh / gelu_new / residual / row-gather (wpe[pos] + wte[type])
// epilogue. Both weight layouts are consumed in place:
// HF Conv1D [K, N] (ldmatrix.trans) and nn.Linear / tied wte [N, K].
#include <string.h>

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

// Split-K across a thread-block cluster: cluster dim (1, KS, 1); CTA (x = 64-column block,
// y = k-slice). Every CTA issues ALL cp.async loads of its k-slice at once (<= MAX_KB tiles, one
// memory round trip, no per-tile latency), multiplies, parks its 64x64 fp32 partial in its own
// shared memory, and after a cluster barrier reduces 64/KS rows of the tile over all peers
// through distributed shared memory, applying the epilogue to those rows.
constexpr int MAX_KB = 6;   // 64-wide k tiles per CTA
constexpr int PART_PITCH = 68;  // floats

__global__ void __launch_bounds__(ATT_THREADS)
skinny_gemm_kernel(const SkinnyParams p, int ks, int kb_per) {
  extern __shared__ __align__(128) uint8_t sk_smem[];
  bf16* sX = reinterpret_cast<bf16*>(sk_smem);                    // [kb_per][64*64]
  bf16* sW = sX + kb_per * 64 * 64;                               // [kb_per][64*64]
  float* part = reinterpret_cast<float*>(sW + kb_per * 64 * 64);  // [64][PART_PITCH]
  const int n0 = blockIdx.x * 64;
  const int krank = blockIdx.y;  // == rank in cluster (cluster spans the y dimension only)
  const int warp = threadIdx.x >> 5, l = lane_id();
  const int nk = cdiv(p.K, 64);
  const int kb0 = krank * kb_per, kb1 = min(nk, kb0 + kb_per);
  for (int kb = kb0; kb < kb1; ++kb) {
    const int st = kb - kb0;
    bf16* tx = sX + st * 4096;
    bf16* tw = sW + st * 4096;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = threadIdx.x + i * ATT_THREADS;
      const int r = idx >> 3, c = idx & 7;
      // x tile: rows = tokens, cols = k
      const bool okx = r < p.M && (kb * 64 + c * 8) < p.K;
      cp_async16(reinterpret_cast<uint8_t*>(tx) + r * 128 + ((c ^ (r & 7)) << 4),
                 p.x + (long long)(okx ? r : 0) * p.ldx + (okx ? kb * 64 + c * 8 : 0), okx);
      if (p.w_kn) {  // W rows = k, cols = n
        const bool ok = (kb * 64 + r) < p.K && (n0 + c * 8) < p.N;
        cp_async16(reinterpret_cast<uint8_t*>(tw) + r * 128 + ((c ^ (r & 7)) << 4),
                   p.w + (long long)(ok ? kb * 64 + r : 0) * p.ldw + (ok ? n0 + c * 8 : 0), ok);
      } else {       // W rows = n, cols = k
        const bool ok = (n0 + r) < p.N && (kb * 64 + c * 8) < p.K;
        cp_async16(reinterpret_cast<uint8_t*>(tw) + r * 128 + ((c ^ (r & 7)) << 4),
                   p.w + (long long)(ok ? n0 + r : 0) * p.ldw + (ok ? kb * 64 + c * 8 : 0), ok);
      }
    }
  }
  cp_async_commit();
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  cp_async_wait<0>();
  __syncthreads();
  for (int kb = kb0; kb < kb1; ++kb) {
    const int st = kb - kb0;
    uint32_t a[4][4];
    load_a_frags(sX + st * 4096, warp * 16, a);
    if (p.w_kn) mma_nn_a(acc, a, sW + st * 4096);
    else mma_nt(acc, a, sW + st * 4096);
  }
  // park the partial tile: thread holds rows (warp*16 + l/4, +8), cols nb*8 + 2*(l%4) + {0,1}
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = warp * 16 + (l >> 2) + r * 8;
#pragma unroll
    for (int nb = 0; nb < 8; ++nb)
      *reinterpret_cast<float2*>(part + row * PART_PITCH + nb * 8 + (l & 3) * 2) =
          make_float2(acc[nb][2 * r], acc[nb][2 * r + 1]);
  }
  cluster_sync_all();
  // reduce rows [krank*rows_per, +rows_per) over the ks partials (DSMEM) and finish them
  const int rows_per = 64 / ks;
  const uint32_t my_part = smem_u32(part);
  for (int e = threadIdx.x; e < rows_per * 32; e += ATT_THREADS) {  // 32 column pairs per row
    const int row = krank * rows_per + (e >> 5);
    const int cp = (e & 31) * 2;
    float2 v = make_float2(0.f, 0.f);
    for (int rk = 0; rk < ks; ++rk) {
      uint32_t ra;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(my_part + (uint32_t)((row * PART_PITCH + cp) * 4)), "r"(rk));
      float2 t;
      asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(t.x), "=f"(t.y) : "r"(ra));
      v.x += t.x;
      v.y += t.y;
    }
    if (row < p.M) {
      long long t0 = 0, t1 = 0;
      if (p.rowtab0) t0 = (long long)p.rowidx0[row] * p.ldt0;
      if (p.rowtab1) t1 = (long long)p.rowidx1[row] * p.ldt1;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int cc = n0 + cp + q;
        if (cc >= p.N) continue;
        float o = q == 0 ? v.x : v.y;
        if (p.bias) o += __ldg(p.bias + cc);
        if (p.act == MMTG_ACT_TANH) o = tanhf(o);
        else if (p.act == MMTG_ACT_GELU_NEW) o = gelu_new_f(o);
        if (p.residual) o += __ldg(p.residual + (long long)row * p.ldr + cc);
        if (p.rowtab0) o += __ldg(p.rowtab0 + t0 + cc);
        if (p.rowtab1) o += __ldg(p.rowtab1 + t1 + cc);
        if (p.out_bf16) ((bf16*)p.out)[(long long)row * p.ldo + cc] = __float2bfloat16(o);
        else ((float*)p.out)[(long long)row * p.ldo + cc] = o;
      }
    }
  }
  cluster_sync_all();  // peers may still be reading this CTA's partial
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
  // k-split: enough CTAs to pull the weights at full bandwidth, at most MAX_KB tiles per CTA
  // k-split: aim at <= 3 k-tiles per CTA (66 KB of smem -> 3 CTAs per SM, the whole grid is
  // resident in one wave and pulls the weights with one memory round trip), at most 8-way
  const int nk = cdiv(a->K, 64);
  int ks = 1;
  while (ks < 8 && cdiv(nk, ks) > 3) ks *= 2;
  while (ks < 8 && cdiv(a->N, 64) * ks < 64 && ks * 2 <= nk) ks *= 2;
  MMTG_CHECK_ARG(cdiv(nk, ks) <= MAX_KB, "skinny GEMM: K=%d too large (max %d)", a->K, 8 * MAX_KB * 64);
  const int kb_per = cdiv(nk, ks);
  const int SMEM = 2 * kb_per * 64 * 64 * 2 + 64 * PART_PITCH * 4;
  MMTG_PER_DEVICE_FLAG(attr_set);
  if (!attr_set) {
    MMTG_CUDA_OK(cudaFuncSetAttribute(skinny_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      2 * MAX_KB * 64 * 64 * 2 + 64 * PART_PITCH * 4));
    attr_set = true;
  }
  ProfScope prof(0, 2.0 * a->M * a->N * a->K, 2.0 * a->N * a->K, st);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(cdiv(a->N, 64), ks);
  cfg.blockDim = dim3(ATT_THREADS);
  cfg.dynamicSmemBytes = SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = ks;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  MMTG_CUDA_OK(cudaLaunchKernelEx(&cfg, skinny_gemm_kernel, p, ks, kb_per));
  MMTG_LAUNCH_OK();
  count_launch();
  return 0;
}

}  // namespace mmtg

extern "C" int mmtg_skinny_gemm_bf16(const mmtg_gemm_args* args, void* stream) {
  MMTG_CHECK_ARG(args && args->A && args->B && args->out, "null args");
  return mmtg::skinny_gemm(args, (cudaStream_t)stream);
}
