// tcgen05 / TMEM / TMA GEMM for sm_100a:  out = epilogue(A · Bᵀ), bf16 operands, fp32 accumulate.
//
// Replaces the cuBLAS calls under nn.Linear / HF Conv1D on the MMTG hot path (forward, dgrad
// and wgrad) — see include/mmtg_b200.h for the reference call sites.
//
// Design (B200-first, not a translation of anything in the reference, which has no kernels):
//  * persistent CTAs (one per SM), static round-robin over (tile, k-split) work units;
//  * warp-specialised: warp 0 = TMA producer, warp 1 = TMEM owner + single-thread tcgen05.mma
//    issuer, warps 2..5 = epilogue (one TMEM lane quadrant each);
//  * operands staged by TMA into a multi-stage SWIZZLE_128B shared-memory ring; both K-major
//    and MN-major operand layouts are consumed straight from HBM (no transposes for wgrad);
//  * fp32 accumulators double-buffered in TMEM (2 x BN columns) so the epilogue of tile i
//    overlaps the MMAs of tile i+1;
//  * epilogue: tcgen05.ld -> per-warp padded smem transpose -> coalesced global stores with
//    fused bias / tanh / gelu_new / dgelu / residual / row-gather adds / column sums / split-K
//    atomics / per-row log-sum-exp partials.
#include <mutex>
#include <unordered_map>

#include "../../include/mmtg_b200.h"
#include "common.cuh"

namespace mmtg {

struct GemmParams {
  int M, N, K;
  int num_m, num_n, num_kb, kb_per_split, splits;
  int a_mn, b_mn;
  // epilogue
  void* out;
  long long ldo;
  int out_bf16;
  int atomic;
  bf16* out2;
  long long ldo2;
  const float* bias;
  int act;
  const float* residual;
  long long ldr;
  const bf16* dgelu_src;
  long long ldg;
  const float* rowtab0;
  const int* rowidx0;
  long long ldt0;
  int rowmod0;
  const float* rowtab1;
  const int* rowidx1;
  long long ldt1;
  float* colsum;
  float* lse_partial;
  int vec4;
  int dact_tanh_out;
};

template <int BN>
struct GemmCfg {
  static constexpr int BM = 128;
  static constexpr int BK = 64;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int EPI_PITCH = 36;  // floats; 16-B aligned rows, conflict-free
  static constexpr int EPI_WARP_FLOATS = 32 * EPI_PITCH;
  static constexpr int EPI_BYTES = 4 * EPI_WARP_FLOATS * 4;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES + 1024;
  static constexpr int TMEM_COLS = 2 * BN;  // 256 or 512: power of two
  static constexpr int THREADS = 192;
};

template <int BN>
__global__ void __launch_bounds__(192, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA,
                         const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using C = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B atoms need 1024-B alignment; offset (not cast) keeps the shared address space
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* epi_all = (float*)(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* full = (uint64_t*)(smem + C::STAGES * C::STAGE_BYTES + C::EPI_BYTES);
  uint64_t* empty = full + C::STAGES;
  uint64_t* tfull = empty + C::STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = (uint32_t*)(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 4);
    }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total = p.num_m * p.num_n * p.splits;

  if (warp == 0) {
    // ===================== TMA producer =====================
    int s = 0;
    uint32_t ph = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
      const int tile = w / p.splits, ks = w - tile * p.splits;
      const int n_t = tile / p.num_m, m_t = tile - n_t * p.num_m;
      const int m0 = m_t * C::BM, n0 = n_t * BN;
      const int kb0 = ks * p.kb_per_split;
      const int kb1 = min(p.num_kb, kb0 + p.kb_per_split);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty[s], ph ^ 1);
        if (lane == 0) {
          uint8_t* sA = smem + s * C::STAGE_BYTES;
          uint8_t* sB = sA + C::A_BYTES;
          mbar_arrive_expect_tx(&full[s], C::STAGE_BYTES);
          if (!p.a_mn) {
            tma_load_2d(sA, &tmA, &full[s], kb * C::BK, m0);
          } else {
#pragma unroll
            for (int j = 0; j < C::BM / 64; ++j)
              tma_load_2d(sA + j * 8192, &tmA, &full[s], m0 + 64 * j, kb * C::BK);
          }
          if (!p.b_mn) {
            tma_load_2d(sB, &tmB, &full[s], kb * C::BK, n0);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(sB + j * 8192, &tmB, &full[s], n0 + 64 * j, kb * C::BK);
          }
        }
        __syncwarp();
        if (++s == C::STAGES) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = umma_idesc_bf16(C::BM, BN, p.a_mn, p.b_mn);
    // K-major SW128: 8-row atoms 1024 B apart (SBO), UMMA_K=16 advances 32 B inside the row.
    // MN-major SW128: 64-element MN chunks 8192 B apart (LBO), 8-row K groups 1024 B apart
    // (SBO), UMMA_K=16 advances 16 rows = 2048 B.
    const uint32_t a_lbo = p.a_mn ? 8192u : 16u, b_lbo = p.b_mn ? 8192u : 16u;
    const uint32_t a_kstep = p.a_mn ? 2048u : 32u, b_kstep = p.b_mn ? 2048u : 32u;
    int s = 0;
    uint32_t ph = 0;
    int acc = 0;
    uint32_t acc_ph = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
      const int tile = w / p.splits, ks = w - tile * p.splits;
      const int kb0 = ks * p.kb_per_split;
      const int kb1 = min(p.num_kb, kb0 + p.kb_per_split);
      mbar_wait(&tempty[acc], acc_ph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_addr = smem_u32(smem + s * C::STAGE_BYTES);
          const uint32_t b_addr = a_addr + C::A_BYTES;
#pragma unroll
          for (int k = 0; k < C::BK / 16; ++k) {
            const uint64_t ad = umma_desc_sw128(a_addr + k * a_kstep, a_lbo, 1024u);
            const uint64_t bd = umma_desc_sw128(b_addr + k * b_kstep, b_lbo, 1024u);
            umma_bf16(d_tmem, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty[s]);  // frees the smem slot when these MMAs retire
        }
        __syncwarp();
        if (++s == C::STAGES) {
          s = 0;
          ph ^= 1;
        }
      }
      if (lane == 0) umma_commit(&tfull[acc]);  // accumulator complete -> epilogue
      __syncwarp();
      if (++acc == 2) {
        acc = 0;
        acc_ph ^= 1;
      }
    }
  } else {
    // ===================== epilogue warps (2..5) =====================
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    float* epi = epi_all + (warp - 2) * C::EPI_WARP_FLOATS;
    int acc = 0;
    uint32_t acc_ph = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
      const int tile = w / p.splits;
      const int n_t = tile / p.num_m, m_t = tile - n_t * p.num_m;
      const int m0 = m_t * C::BM, n0 = n_t * BN;
      const int row_base = m0 + q * 32;
      mbar_wait(&tfull[acc], acc_ph);
      tc_fence_after();
      float run_max = -INFINITY, run_sum = 0.f;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        const int col0 = n0 + c * 32;
        if (col0 >= p.N) break;  // warp-uniform
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c * 32), r);
        tmem_ld_wait();
        if (p.lse_partial) {
          float cm = -INFINITY;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < p.N) cm = fmaxf(cm, __uint_as_float(r[j]));
          const float nm = fmaxf(run_max, cm);
          float add = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < p.N) add += __expf(__uint_as_float(r[j]) - nm);
          run_sum = run_sum * __expf(run_max - nm) + add;
          run_max = nm;
        }
        {  // phase 1: thread = row, 8 x STS.128 (pitch 36 floats -> conflict-free quarter-warps)
          float4* dst = reinterpret_cast<float4*>(epi + lane * C::EPI_PITCH);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            dst[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                 __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
        }
        __syncwarp();
        if (p.vec4) {
          // phase 2 (vector): lane -> (row sub-index rs, 4-column group cg); 8 iterations of
          // 4 rows x 32 columns, every global access is a full 16 B (fp32) / 8 B (bf16) vector.
          const int rs = lane >> 3, cg = lane & 7;
          const int col = col0 + cg * 4;
          const bool colok = col < p.N;  // N % 4 == 0 in vec mode
          float v[8][4];
          bool ok[8];
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.bias && colok) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rl = it * 4 + rs;
            const float4 t = *reinterpret_cast<const float4*>(epi + rl * C::EPI_PITCH + cg * 4);
            v[it][0] = t.x + b4.x; v[it][1] = t.y + b4.y; v[it][2] = t.z + b4.z; v[it][3] = t.w + b4.w;
            ok[it] = colok && (row_base + rl) < p.M;
          }
          const long long row0 = row_base + rs;
          if (p.out2) {
#pragma unroll
            for (int it = 0; it < 8; ++it)
              if (ok[it]) {
                __nv_bfloat162 lo = __floats2bfloat162_rn(v[it][0], v[it][1]);
                __nv_bfloat162 hi = __floats2bfloat162_rn(v[it][2], v[it][3]);
                uint2 pk = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
                *reinterpret_cast<uint2*>(p.out2 + (row0 + it * 4) * p.ldo2 + col) = pk;
              }
          }
          if (p.act == MMTG_ACT_TANH) {
#pragma unroll
            for (int it = 0; it < 8; ++it)
#pragma unroll
              for (int e = 0; e < 4; ++e) v[it][e] = tanhf(v[it][e]);
          } else if (p.act == MMTG_ACT_GELU_NEW) {
#pragma unroll
            for (int it = 0; it < 8; ++it)
#pragma unroll
              for (int e = 0; e < 4; ++e) v[it][e] = gelu_new_f(v[it][e]);
          }
          if (p.dgelu_src) {
            uint2 u[8];
#pragma unroll
            for (int it = 0; it < 8; ++it)
              u[it] = ok[it] ? __ldg(reinterpret_cast<const uint2*>(p.dgelu_src + (row0 + it * 4) * p.ldg + col))
                             : make_uint2(0u, 0u);
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const float2 a = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u[it].x));
              const float2 b = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u[it].y));
              if (p.dact_tanh_out) {
                v[it][0] *= 1.f - a.x * a.x; v[it][1] *= 1.f - a.y * a.y;
                v[it][2] *= 1.f - b.x * b.x; v[it][3] *= 1.f - b.y * b.y;
              } else {
                v[it][0] *= dgelu_new_f(a.x); v[it][1] *= dgelu_new_f(a.y);
                v[it][2] *= dgelu_new_f(b.x); v[it][3] *= dgelu_new_f(b.y);
              }
            }
          }
          if (p.residual) {
            float4 t[8];
#pragma unroll
            for (int it = 0; it < 8; ++it)
              t[it] = ok[it] ? __ldg(reinterpret_cast<const float4*>(p.residual + (row0 + it * 4) * p.ldr + col))
                             : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              v[it][0] += t[it].x; v[it][1] += t[it].y; v[it][2] += t[it].z; v[it][3] += t[it].w;
            }
          }
          if (p.rowtab0) {
#pragma unroll
            for (int it = 0; it < 8; ++it)
              if (ok[it]) {
                const long long row = row0 + it * 4;
                const long long ti = p.rowidx0 ? (long long)__ldg(p.rowidx0 + row) : (long long)(row % p.rowmod0);
                const float4 t = __ldg(reinterpret_cast<const float4*>(p.rowtab0 + ti * p.ldt0 + col));
                v[it][0] += t.x; v[it][1] += t.y; v[it][2] += t.z; v[it][3] += t.w;
              }
          }
          if (p.rowtab1) {
#pragma unroll
            for (int it = 0; it < 8; ++it)
              if (ok[it]) {
                const long long ti = (long long)__ldg(p.rowidx1 + row0 + it * 4);
                const float4 t = __ldg(reinterpret_cast<const float4*>(p.rowtab1 + ti * p.ldt1 + col));
                v[it][0] += t.x; v[it][1] += t.y; v[it][2] += t.z; v[it][3] += t.w;
              }
          }
          if (p.colsum) {
            float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int it = 0; it < 8; ++it)
              if (ok[it]) {
#pragma unroll
                for (int e = 0; e < 4; ++e) s4[e] += v[it][e];
              }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              s4[e] += __shfl_xor_sync(0xffffffffu, s4[e], 8);
              s4[e] += __shfl_xor_sync(0xffffffffu, s4[e], 16);
            }
            if (rs == 0 && colok) {
#pragma unroll
              for (int e = 0; e < 4; ++e) atomicAdd(p.colsum + col + e, s4[e]);
            }
          }
          if (p.atomic) {
#pragma unroll
            for (int it = 0; it < 8; ++it)
              if (ok[it]) {
                float* dstp = (float*)p.out + (row0 + it * 4) * p.ldo + col;
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dstp),
                             "f"(v[it][0]), "f"(v[it][1]), "f"(v[it][2]), "f"(v[it][3])
                             : "memory");
              }
          } else if (p.out_bf16) {
#pragma unroll
            for (int it = 0; it < 8; ++it)
              if (ok[it]) {
                __nv_bfloat162 lo = __floats2bfloat162_rn(v[it][0], v[it][1]);
                __nv_bfloat162 hi = __floats2bfloat162_rn(v[it][2], v[it][3]);
                uint2 pk = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
                *reinterpret_cast<uint2*>((bf16*)p.out + (row0 + it * 4) * p.ldo + col) = pk;
              }
          } else {
#pragma unroll
            for (int it = 0; it < 8; ++it)
              if (ok[it])
                *reinterpret_cast<float4*>((float*)p.out + (row0 + it * 4) * p.ldo + col) =
                    make_float4(v[it][0], v[it][1], v[it][2], v[it][3]);
          }
        } else {
          // phase 2 (scalar, unaligned pitches e.g. the contiguous [.., 13317] logits):
          // lane = column, coalesced 4-byte accesses; supports bias / act / residual / colsum.
          const int col = col0 + lane;
          const bool colok = col < p.N;
          const float bias_v = (p.bias && colok) ? __ldg(p.bias + col) : 0.f;
          float csum = 0.f;
          const int nrows = min(32, p.M - row_base);  // warp-uniform, may be <= 0
#pragma unroll 8
          for (int rr = 0; rr < nrows; ++rr) {
            const long long row = row_base + rr;
            float v = epi[rr * C::EPI_PITCH + lane] + bias_v;
            if (colok) {
              if (p.act == MMTG_ACT_TANH) v = tanhf(v);
              else if (p.act == MMTG_ACT_GELU_NEW) v = gelu_new_f(v);
              if (p.residual) v += __ldg(p.residual + row * p.ldr + col);
              csum += v;
              if (p.atomic) atomicAdd((float*)p.out + row * p.ldo + col, v);
              else if (p.out_bf16) ((bf16*)p.out)[row * p.ldo + col] = __float2bfloat16(v);
              else ((float*)p.out)[row * p.ldo + col] = v;
            }
          }
          if (p.colsum && colok && nrows > 0) atomicAdd(p.colsum + col, csum);
        }
        __syncwarp();
      }
      if (p.lse_partial) {
        const long long row = row_base + lane;
        if (row < p.M) {
          float* dst = p.lse_partial + ((long long)n_t * p.M + row) * 2;
          dst[0] = run_max;
          dst[1] = run_sum;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_ph ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// host side: tensor-map cache + launcher
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

struct TmapKey {
  const void* ptr;
  uint64_t inner, outer, ld;
  uint32_t box0, box1;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && inner == o.inner && outer == o.outer && ld == o.ld && box0 == o.box0 &&
           box1 == o.box1;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = (size_t)k.ptr;
    h = h * 1000003u ^ k.inner;
    h = h * 1000003u ^ k.outer;
    h = h * 1000003u ^ k.ld;
    h = h * 1000003u ^ (k.box0 * 1315423911u + k.box1);
    return h;
  }
};

// 2-D bf16 tensor map over a row-major [outer, inner] matrix with row pitch ld (elements).
int make_tmap_bf16_2d(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer,
                      uint64_t ld, uint32_t box0, uint32_t box1) {
  static std::mutex mu;
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  TmapKey key{ptr, inner, outer, ld, box0, box1};
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return 0;
    }
  }
  EncodeTiledFn enc = get_encode_fn();
  MMTG_CHECK_ARG(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  MMTG_CHECK_ARG(((uintptr_t)ptr & 15) == 0, "TMA operand base %p not 16-byte aligned", ptr);
  MMTG_CHECK_ARG((ld * 2) % 16 == 0, "TMA operand pitch %llu elements not a multiple of 8",
                 (unsigned long long)ld);
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box0, box1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)ptr, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MMTG_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) ptr=%p dims=%llu,%llu ld=%llu",
                 (int)r, ptr, (unsigned long long)inner, (unsigned long long)outer,
                 (unsigned long long)ld);
  {
    std::lock_guard<std::mutex> g(mu);
    if (cache.size() > 4096) cache.clear();
    cache.emplace(key, *out);
  }
  return 0;
}

void count_launch(int n = 1);

template <int BN>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p,
                       cudaStream_t st) {
  using C = GemmCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    MMTG_CUDA_OK(cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  const int total = p.num_m * p.num_n * p.splits;
  const int grid = total < num_sms() ? total : num_sms();
  gemm_bf16_tcgen05_kernel<BN><<<grid, C::THREADS, C::SMEM_BYTES, st>>>(tmA, tmB, p);
  MMTG_LAUNCH_OK();
  count_launch();
  return 0;
}

}  // namespace mmtg

using namespace mmtg;

extern "C" int mmtg_gemm_bf16(const mmtg_gemm_args* a, void* stream) {
  MMTG_CHECK_ARG(a != nullptr, "null args");
  MMTG_CHECK_ARG(a->A && a->B && a->out, "null A/B/out");
  MMTG_CHECK_ARG(a->M > 0 && a->N > 0 && a->K > 0, "bad shape M=%d N=%d K=%d", a->M, a->N, a->K);
  int BN = a->block_n;
  if (BN == 0) {
    // 256-wide tiles unless that leaves most of the machine idle
    const long long t256 = (long long)cdiv(a->M, 128) * cdiv(a->N, 256);
    BN = (a->N >= 256 && t256 >= 96) ? 256 : 128;
  }
  MMTG_CHECK_ARG(BN == 128 || BN == 256, "block_n must be 0, 128 or 256");
  const bool atomic = a->accumulate != 0 || a->split_k > 1;
  MMTG_CHECK_ARG(!(atomic && a->out_dtype != MMTG_F32), "accumulate/split_k need an fp32 out");
  MMTG_CHECK_ARG(!(a->rowtab0 && !a->rowidx0 && a->rowmod0 <= 0), "rowtab0 needs rowidx0 or rowmod0");
  MMTG_CHECK_ARG(!(a->rowtab1 && !a->rowidx1), "rowtab1 needs rowidx1");

  GemmParams p;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.num_m = cdiv(a->M, 128);
  p.num_n = cdiv(a->N, BN);
  p.num_kb = cdiv(a->K, 64);
  int splits = a->split_k > 1 ? a->split_k : 1;
  if (splits > p.num_kb) splits = p.num_kb;
  p.kb_per_split = cdiv(p.num_kb, splits);
  p.splits = cdiv(p.num_kb, p.kb_per_split);
  p.a_mn = a->a_mn_major ? 1 : 0;
  p.b_mn = a->b_mn_major ? 1 : 0;
  p.out = a->out; p.ldo = a->ldo; p.out_bf16 = a->out_dtype == MMTG_BF16; p.atomic = atomic;
  p.out2 = (bf16*)a->out2; p.ldo2 = a->ldo2;
  p.bias = a->bias; p.act = a->act;
  p.residual = a->residual; p.ldr = a->ldr;
  p.dgelu_src = (const bf16*)a->dgelu_src; p.ldg = a->ldg;
  p.rowtab0 = a->rowtab0; p.rowidx0 = a->rowidx0; p.ldt0 = a->ldt0; p.rowmod0 = a->rowmod0;
  p.rowtab1 = a->rowtab1; p.rowidx1 = a->rowidx1; p.ldt1 = a->ldt1;
  p.colsum = a->colsum;
  p.lse_partial = a->lse_partial;
  p.dact_tanh_out = a->dact_tanh_out;
  MMTG_CHECK_ARG(!(p.lse_partial && p.splits > 1), "lse_partial is incompatible with split_k");
  {
    // vector epilogue needs 16-B aligned fp32 rows / 8-B aligned bf16 rows for every operand
    auto al = [](const void* q, long long ld, int esz) {
      return q == nullptr || ((((uintptr_t)q) % (size_t)(4 * esz)) == 0 && ld % 4 == 0);
    };
    const int oesz = p.out_bf16 ? 2 : 4;
    p.vec4 = (a->N % 4 == 0) && al(p.out, p.ldo, oesz) && al(p.out2, p.ldo2, 2) &&
             al(p.bias, 4, 4) && al(p.residual, p.ldr, 4) && al(p.dgelu_src, p.ldg, 2) &&
             al(p.rowtab0, p.ldt0, 4) && al(p.rowtab1, p.ldt1, 4) && al(p.colsum, 4, 4);
    MMTG_CHECK_ARG(p.vec4 || !(p.out2 || p.dgelu_src || p.rowtab0 || p.rowtab1),
                   "out2/dgelu_src/rowtab epilogues need N %% 4 == 0 and 16-byte aligned operands");
  }

  CUtensorMap tmA, tmB;
  if (!p.a_mn) MMTG_TRY(make_tmap_bf16_2d(&tmA, a->A, a->K, a->M, a->lda, 64, 128));
  else         MMTG_TRY(make_tmap_bf16_2d(&tmA, a->A, a->M, a->K, a->lda, 64, 64));
  if (!p.b_mn) MMTG_TRY(make_tmap_bf16_2d(&tmB, a->B, a->K, a->N, a->ldb, 64, BN));
  else         MMTG_TRY(make_tmap_bf16_2d(&tmB, a->B, a->N, a->K, a->ldb, 64, 64));

  cudaStream_t st = (cudaStream_t)stream;
  if (BN == 256) return launch_gemm<256>(tmA, tmB, p, st);
  return launch_gemm<128>(tmA, tmB, p, st);
}
