// tcgen05 / TMEM / TMA GEMM for sm_100a:  out = epilogue(A · Bᵀ), bf16 operands, fp32 accumulate.
//
// Replaces the cuBLAS calls under nn.Linear / HF Conv1D on the MMTG hot path (forward, dgrad
// and wgrad) — see include/mmtg_b200.h for the reference call sites.
//
// Design (B200-first, not a translation of anything in the reference, which has no kernels):
//  * persistent CTAs (one per SM) in clusters of 2: the pair works on two vertically adjacent
//    128-row tiles of the same column block, so the B tile is fetched ONCE per pair — each CTA
//    loads half of it and TMA-multicasts it into both CTAs' shared memory (the mainloop is
//    L2-bandwidth bound otherwise: 48 KB per CTA per k-block -> 32 KB);
//  * static round-robin of the clusters over (tile pair, k-split) work units (default); opt-in:
//    cluster-launch-control dynamic scheduling, a stream-K split of the last partial round and
//    192-column tiles - all three measured, none faster (see the comments at their switches);
//  * warp-specialised: warp 0 = TMA producer, warp 1 = TMEM owner + single-thread tcgen05.mma
//    issuer, warps 2..9 = epilogue (two per TMEM lane quadrant, half the tile's columns each),
//    warp 10 = tile scheduler of the dynamic mode (idle otherwise);
//  * operands staged by TMA into a multi-stage SWIZZLE_128B shared-memory ring; both K-major
//    and MN-major operand layouts are consumed straight from HBM (no transposes for wgrad);
//  * fp32 accumulators double-buffered in TMEM (2 x BN columns) so the epilogue of tile i
//    overlaps the MMAs of tile i+1;
//  * epilogue: tcgen05.ld -> per-warp padded smem transpose -> coalesced global stores with
//    fused bias / tanh / gelu_new / dgelu / residual / row-gather adds / column sums / split-K
//    atomics / per-row log-sum-exp partials.
#include <mutex>
#include <stdlib.h>
#include <string.h>
#include <unordered_map>

#include "../../include/mmtg_b200.h"
#include "common.cuh"

namespace mmtg {

struct GemmParams {
  int M, N, K;
  int num_m, num_mp, num_n, num_kb, kb_per_split, splits;
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
  int out2_mode;
  const unsigned long long* drop_seed;
  uint32_t drop_site;
  float drop_p;
  int grid_mode;
  int clc;  // 1: work units come from cluster launch control (see TileSched), 0: static round-robin
  // tail split (F_TAIL): work units [0, tail_full) are whole tiles; each tile >= tail_full is cut into
  // tail_S k-slices of tail_kbs k-blocks (slice 0 owns the epilogue and adds the others' partials)
  int tail_full, tail_S, tail_kbs, tail_total;
  float* tail_ws;            // [tail tile][slice - 1][cta rank][128][BN] fp32 partial accumulators
  unsigned int* tail_flags;  // [2][tail tile]: arrivals of the partial slices / owner warps done
};

// TWO = cta_group::2: the pair's 256 x BN tile is ONE MMA; each CTA stages its own 128 A rows and
// only HALF of the B tile (the tensor core reads the other half from the peer SM), which halves
// the shared-memory write+read traffic per FLOP — the 1-SM kernel is smem-bandwidth bound at
// ~2/3 of the tensor peak (12 KB operand reads + 12 KB TMA writes per 128-cycle MMA).
template <int BN, bool TWO>
struct GemmCfg {
  static constexpr int BM = 128;
  static constexpr int BK = 64;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (TWO ? BN / 2 : BN) * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (232448 - 8 * 32 * 32 * 4 - 512 - 1024) / STAGE_BYTES;
  static constexpr int EPI_PITCH = 32;  // floats; XOR-swizzled 16-B chunks, no padding
  static constexpr int EPI_WARP_FLOATS = 32 * EPI_PITCH;
  static constexpr int EPI_BYTES = 8 * EPI_WARP_FLOATS * 4;
  static constexpr int BAR_BYTES = 512;  // ring / accumulator barriers [0,192), TMEM slot 192, scheduler 256..
  static_assert((2 * STAGES + 4) * 8 <= 192, "barrier block overflows into the TMEM slot");
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES + 1024;
  static constexpr int TMEM_COLS = BN == 192 ? 512 : 2 * BN;  // allocations are powers of two (192: 384 used)
  static constexpr int THREADS = 352;  // 10 working warps + the tile-scheduler warp
};

// epilogue feature bits (compile-time mask EPI)
enum : uint32_t {
  F_OUT2 = 1, F_ACT = 2, F_DACT = 4, F_RES = 8, F_ROWTAB = 16, F_COLSUM = 32, F_ATOMIC = 64,
  F_LSE = 128, F_SCALAR = 256, F_DROP = 512, F_TAIL = 1024
};

// One work item of a cluster: tile index, k-block range, and (tail split) which slice of how many.
struct WorkUnit {
  int tile, kb0, kb1, slice, nsl;
};
template <uint32_t EPI>
__device__ __forceinline__ WorkUnit decode_unit(const GemmParams& p, int w) {
  WorkUnit u;
  if constexpr (EPI & F_TAIL) {
    if (w < p.tail_full) {
      u.tile = w; u.kb0 = 0; u.kb1 = p.num_kb; u.slice = 0; u.nsl = 1;
    } else {
      const int t = w - p.tail_full;
      u.tile = p.tail_full + t / p.tail_S;
      u.slice = t - (u.tile - p.tail_full) * p.tail_S;
      u.kb0 = u.slice * p.tail_kbs;
      u.kb1 = min(p.num_kb, u.kb0 + p.tail_kbs);
      u.nsl = p.tail_S;
    }
  } else {
    u.tile = w / p.splits;
    u.slice = w - u.tile * p.splits;
    u.kb0 = u.slice * p.kb_per_split;
    u.kb1 = min(p.num_kb, u.kb0 + p.kb_per_split);
    u.nsl = 1;  // (uniform split-K combines through atomics, not through the fix-up)
  }
  return u;
}

// ---------------------------------------------------------------------------------------------
// Work distribution. Static mode: cluster c takes units c, c + #clusters, ... (a CTA pair that
// cannot become resident - its SMs are held by a kernel of another stream: a weight-gradient GEMM
// of the side stream, an NCCL all-reduce - keeps its units hostage until it starts). Dynamic mode
// (sm_100 cluster launch control): the grid has ONE cluster per work unit; a running cluster,
// when it finishes a unit, cancels a cluster that has not been launched yet and takes over its
// unit. Whatever SMs are available at any moment share the remaining work.
// The scheduler warp of the pair's CTA 0 issues `clusterlaunchcontrol.try_cancel` (multicast: the
// 16-byte response lands in both CTAs' shared memory and completes 16 bytes on both CTAs'
// `clc_full` barriers); the producer, MMA and epilogue warps of both CTAs read it and release
// the slot on CTA 0's `clc_empty` (20 arrivals).
// ---------------------------------------------------------------------------------------------
constexpr int CLC_STAGES = 3;
constexpr int CLC_CONSUMERS = 20;  // (producer + MMA + 8 epilogue warps) x 2 CTAs

__device__ __forceinline__ void clc_try_cancel(void* resp, uint64_t* bar) {
  asm volatile(
      "clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes.multicast::cluster::all.b128 [%0], [%1];"
      ::"r"(smem_u32(resp)), "r"(smem_u32(bar))
      : "memory");
}
// -> first CTA id (x) of the cancelled cluster, or -1 when nothing was left to cancel
__device__ __forceinline__ int clc_read(const void* resp) {
  uint32_t x, y, z, valid;
  asm volatile(
      "{\n"
      ".reg .pred p1;\n"
      ".reg .b128 r;\n"
      "ld.shared.b128 r, [%4];\n"
      "clusterlaunchcontrol.query_cancel.is_canceled.pred.b128 p1, r;\n"
      "selp.u32 %3, 1, 0, p1;\n"
      "mov.u32 %0, 0;\n"
      "mov.u32 %1, 0;\n"
      "mov.u32 %2, 0;\n"
      "@p1 clusterlaunchcontrol.query_cancel.get_first_ctaid.v4.b32.b128 {%0, %1, %2, _}, r;\n"
      "}\n"
      : "=r"(x), "=r"(y), "=r"(z), "=r"(valid)
      : "r"(smem_u32(resp))
      : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the slot is rewritten through the async proxy
  return valid ? (int)x : -1;
}
__device__ __forceinline__ void mbar_arrive_expect_tx_cluster(uint64_t* bar, uint32_t rank, uint32_t bytes) {
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, %1;\n"
      "mbarrier.arrive.expect_tx.shared::cluster.b64 _, [ra], %2;\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(rank), "r"(bytes)
      : "memory");
}
// per-warp cursor over the work units of this cluster
struct TileSched {
  int w;
  uint32_t stage, phase;
  int step, total, clc;
  uint64_t *full, *empty;
  const uint8_t* resp;
  __device__ __forceinline__ bool valid() const { return w >= 0 && w < total; }
  // whole warp calls; lane 0 releases the slot
  __device__ __forceinline__ void next() {
    if (!clc) {
      w += step;
      return;
    }
    mbar_wait<7>(&full[stage], phase);
    const int cta = clc_read(resp + stage * 16);
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive_cluster(&empty[stage], 0);
    if (++stage == CLC_STAGES) {
      stage = 0;
      phase ^= 1;
    }
    w = cta < 0 ? -1 : (cta >> 1);
  }
};

template <int BN, uint32_t EPI, bool TWO>
__global__ void __launch_bounds__(352, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA,
                         const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using C = GemmCfg<BN, TWO>;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B atoms need 1024-B alignment; offset (not cast) keeps the shared address space
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* epi_all = (float*)(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* full = (uint64_t*)(smem + C::STAGES * C::STAGE_BYTES + C::EPI_BYTES);
  uint64_t* empty = full + C::STAGES;
  uint64_t* tfull = empty + C::STAGES;
  uint64_t* tempty = tfull + 2;
  uint8_t* bar_blk = smem + C::STAGES * C::STAGE_BYTES + C::EPI_BYTES;
  uint32_t* tmem_slot = (uint32_t*)(bar_blk + 192);
  uint64_t* clc_full = (uint64_t*)(bar_blk + 256);  // [CLC_STAGES]
  uint64_t* clc_empty = (uint64_t*)(bar_blk + 288);
  uint8_t* clc_resp = bar_blk + 320;                // [CLC_STAGES][16]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // let the next PDL-launched kernel of the stream be scheduled as soon as SMs free up (it
  // blocks in its own griddepcontrol.wait until this grid has fully completed)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full[s], 1);
      // 1-SM mode: MMA commits of both CTAs of the pair (multicast arrivals);
      // 2-SM mode: the leader's single commit, multicast to both CTAs
      mbar_init(&empty[s], TWO ? 1 : 2);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], TWO ? 16 : 8);  // 2-SM: both CTAs' epilogue warps release the leader
    }
    for (int a = 0; a < CLC_STAGES; ++a) {
      mbar_init(&clc_full[a], 1);
      mbar_init(&clc_empty[a], CLC_CONSUMERS);
    }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    if (TWO) {
      tmem_alloc2(tmem_slot, C::TMEM_COLS);
      tmem_relinquish2();
    } else {
      tmem_alloc(tmem_slot, C::TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // peer barriers must be initialised before any multicast can land
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: this grid may have been scheduled while the previous kernel
  // of the stream was still draining (its prologue above overlapped that tail); nothing below
  // may touch global memory before the previous grid has completed and flushed.
  asm volatile("griddepcontrol.wait;" ::: "memory");

  const uint32_t crank = cluster_ctarank();       // 0 / 1 within the pair
  const int cluster_id = blockIdx.x >> 1;
  const int nclusters = gridDim.x >> 1;
  const int total = (EPI & F_TAIL) ? p.tail_total : p.num_mp * p.num_n * p.splits;  // work units of a PAIR
  TileSched ts{cluster_id, 0u, 0u, nclusters, total, p.clc, clc_full, clc_empty, clc_resp};

  if (warp == 10) {
    // ===================== tile scheduler (dynamic mode, CTA 0 of the pair) =====================
    if (p.clc && crank == 0 && lane == 0) {
      uint32_t stage = 0, phase = 0;
      while (true) {
        mbar_wait<8>(&clc_empty[stage], phase ^ 1);
        mbar_arrive_expect_tx_cluster(&clc_full[stage], 0, 16);
        mbar_arrive_expect_tx_cluster(&clc_full[stage], 1, 16);
        clc_try_cancel(clc_resp + stage * 16, &clc_full[stage]);
        mbar_wait<9>(&clc_full[stage], phase);
        const int cta = clc_read(clc_resp + stage * 16);
        if (++stage == CLC_STAGES) {
          stage = 0;
          phase ^= 1;
        }
        if (cta < 0) break;  // nothing left to cancel: further requests are not allowed
      }
    }
  } else if (warp == 0) {
    // ===================== TMA producer =====================
    int s = 0;
    uint32_t ph = 0;
    for (; ts.valid(); ts.next()) {
      const int w = ts.w;
      const WorkUnit wu = decode_unit<EPI>(p, w);
      const int tile = wu.tile;
      const int n_t = tile / p.num_mp, m_t = (tile - n_t * p.num_mp) * 2 + (int)crank;
      const int m0 = m_t * C::BM, n0 = n_t * BN;
      const int kb0 = wu.kb0, kb1 = wu.kb1;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait<1>(&empty[s], ph ^ 1);
        if (lane == 0) {
          uint8_t* sA = smem + s * C::STAGE_BYTES;
          uint8_t* sB = sA + C::A_BYTES;
          if (TWO) {
            // both CTAs' loads complete on the LEADER's barrier, which expects all four boxes
            if (crank == 0) mbar_arrive_expect_tx(&full[s], 2 * C::STAGE_BYTES);
            if (!p.a_mn) {
              tma_load_2d_2sm(sA, &tmA, &full[s], kb * C::BK, m0);
            } else {
#pragma unroll
              for (int j = 0; j < C::BM / 64; ++j)
                tma_load_2d_2sm(sA + j * 8192, &tmA, &full[s], m0 + 64 * j, kb * C::BK);
            }
            if (!p.b_mn) {  // this CTA's half of the B tile: rows [crank*BN/2, +BN/2)
              tma_load_2d_2sm(sB, &tmB, &full[s], kb * C::BK, n0 + (int)crank * (BN / 2));
            } else {
#pragma unroll
              for (int j = 0; j < BN / 128; ++j)
                tma_load_2d_2sm(sB + j * 8192, &tmB, &full[s], n0 + (int)crank * (BN / 2) + 64 * j,
                                kb * C::BK);
            }
          } else {
            // A tile (own) + the whole B tile: half from this CTA, half multicast by the peer
            mbar_arrive_expect_tx(&full[s], C::STAGE_BYTES);
            if (!p.a_mn) {
              tma_load_2d(sA, &tmA, &full[s], kb * C::BK, m0);
            } else {
#pragma unroll
              for (int j = 0; j < C::BM / 64; ++j)
                tma_load_2d(sA + j * 8192, &tmA, &full[s], m0 + 64 * j, kb * C::BK);
            }
            if (!p.b_mn) {  // rows [crank*BN/2, +BN/2) of the B tile -> both CTAs
              tma_load_2d_mc(sB + crank * (C::B_BYTES / 2), &tmB, &full[s], kb * C::BK,
                             n0 + (int)crank * (BN / 2), (uint16_t)3);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 128; ++j) {
                const int jj = (int)crank * (BN / 128) + j;
                tma_load_2d_mc(sB + jj * 8192, &tmB, &full[s], n0 + 64 * jj, kb * C::BK, (uint16_t)3);
              }
            }
          }
        }
        __syncwarp();
        if (++s == C::STAGES) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 1 && TWO && crank != 0) {
    // the peer's MMA warp has no work in 2-SM mode but still consumes the schedule (releases its slots)
    for (; ts.valid(); ts.next()) {
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (2-SM mode: leader CTA only) =====================
    const uint32_t idesc = umma_idesc_bf16(TWO ? 2 * C::BM : C::BM, BN, p.a_mn, p.b_mn);
    // K-major SW128: 8-row atoms 1024 B apart (SBO), UMMA_K=16 advances 32 B inside the row.
    // MN-major SW128: 64-element MN chunks 8192 B apart (LBO), 8-row K groups 1024 B apart
    // (SBO), UMMA_K=16 advances 16 rows = 2048 B.
    const uint32_t a_lbo = p.a_mn ? 8192u : 16u, b_lbo = p.b_mn ? 8192u : 16u;
    const uint32_t a_kstep = p.a_mn ? 2048u : 32u, b_kstep = p.b_mn ? 2048u : 32u;
    int s = 0;
    uint32_t ph = 0;
    int acc = 0;
    uint32_t acc_ph = 0;
    for (; ts.valid(); ts.next()) {
      const WorkUnit wu = decode_unit<EPI>(p, ts.w);
      const int kb0 = wu.kb0, kb1 = wu.kb1;
      mbar_wait<2>(&tempty[acc], acc_ph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait<3>(&full[s], ph);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_addr = smem_u32(smem + s * C::STAGE_BYTES);
          const uint32_t b_addr = a_addr + C::A_BYTES;
#pragma unroll
          for (int k = 0; k < C::BK / 16; ++k) {
            const uint64_t ad = umma_desc_sw128(a_addr + k * a_kstep, a_lbo, 1024u);
            const uint64_t bd = umma_desc_sw128(b_addr + k * b_kstep, b_lbo, 1024u);
            if (TWO) umma_bf16_2sm(d_tmem, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            else umma_bf16(d_tmem, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          // frees the smem slot in BOTH CTAs
          if (TWO) umma_commit2_mc(&empty[s], (uint16_t)3);
          else umma_commit_mc(&empty[s], (uint16_t)3);
        }
        __syncwarp();
        if (++s == C::STAGES) {
          s = 0;
          ph ^= 1;
        }
      }
      if (lane == 0) {  // accumulator complete -> epilogue (of both CTAs in 2-SM mode)
        if (TWO) umma_commit2_mc(&tfull[acc], (uint16_t)3);
        else umma_commit(&tfull[acc]);
      }
      __syncwarp();
      if (++acc == 2) {
        acc = 0;
        acc_ph ^= 1;
      }
    }
  } else if (warp >= 2 && warp < 10) {
    // ===================== epilogue warps (2..9) =====================
    // Two warps per TMEM lane quadrant: warps 2..5 own the left half of the tile's columns,
    // warps 6..9 the right half (a warp may only touch lanes 32*(warp%4) .. +31).
    // EPI is a compile-time feature mask: disabled features generate no code, which keeps the
    // per-chunk instruction footprint inside the instruction cache (the all-features epilogue
    // was measured fetch-bound: stall_no_inst 47 %).
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    constexpr int NCH = BN / 64;  // 32-column chunks per warp
    float* epi = epi_all + (warp - 2) * C::EPI_WARP_FLOATS;
    const int rs = lane >> 3, cg = lane & 7;  // vector mapping: row sub-index, 4-column group
    int acc = 0;
    uint32_t acc_ph = 0;
    DropKey dk = {0u, 0u, 0u, 1.f};
    if constexpr (EPI & F_DROP) dk = drop_key(p.drop_seed, p.drop_site, p.drop_p);
    for (; ts.valid(); ts.next()) {
      const WorkUnit wu = decode_unit<EPI>(p, ts.w);
      const int tile = wu.tile;
      const int n_t = tile / p.num_mp, m_t = (tile - n_t * p.num_mp) * 2 + (int)crank;
      const int m0 = m_t * C::BM, n0 = n_t * BN + half * (BN / 2);
      // tail split: slices > 0 park their raw accumulators in the workspace, slice 0 adds them in
      // slice order before its epilogue (fixed order: bit-reproducible)
      bool tail_part = false, tail_own = false;
      float* tail_blk = nullptr;  // this warp's 32 x (BN/2) block inside the partial tile of slice 1
      constexpr size_t TAIL_TILE = (size_t)2 * C::BM * BN;  // floats per (tile, slice): both CTAs of the pair
      if constexpr (EPI & F_TAIL) {
        if (wu.nsl > 1) {
          tail_part = wu.slice > 0;
          tail_own = !tail_part;
          const size_t tu = (size_t)(tile - p.tail_full);
          tail_blk = p.tail_ws + (tu * (size_t)(p.tail_S - 1) + (size_t)(tail_part ? wu.slice - 1 : 0)) * TAIL_TILE +
                     (size_t)crank * C::BM * BN + (size_t)(q * 32) * BN + half * (BN / 2);
        }
      }
      const int row_base = m0 + q * 32;
      const long long row0 = row_base + rs;
      // bias of the first chunk is fetched BEFORE waiting for the MMAs; later chunks prefetch
      float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
      float bs = 0.f;
      auto load_bias = [&](int c, float4& o4, float& o1) {
        o4 = make_float4(0.f, 0.f, 0.f, 0.f);
        o1 = 0.f;
        if (p.bias) {
          if constexpr (!(EPI & F_SCALAR)) {
            const int col = n0 + c * 32 + cg * 4;
            if (col < p.N) o4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
          } else {
            const int col = n0 + c * 32 + lane;
            if (col < p.N) o1 = __ldg(p.bias + col);
          }
        }
      };
      load_bias(0, b4, bs);
      mbar_wait<4>(&tfull[acc], acc_ph);
      tc_fence_after();
      if constexpr (EPI & F_TAIL) {
        if (tail_own) {  // all partial slices of this tile have landed (16 epilogue warps each)
          if (lane == 0) {
            const unsigned int* flag = p.tail_flags + (tile - p.tail_full);
            const unsigned int want = (unsigned int)(p.tail_S - 1) * 16u;
            unsigned int v, spins = 0;
            do {
              asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
              if (++spins > (1u << 26)) {
                printf("mmtg: GEMM tail-split wait timed out (tile %d, saw %u of %u)\n", tile, v, want);
                __trap();
              }
            } while (v < want);
          }
          __syncwarp();
        }
      }
      float run_max = -INFINITY, run_sum = 0.f;
      bool released = false;
#pragma unroll 1
      for (int c = 0; c < NCH; ++c) {
        const int col0 = n0 + c * 32;
        if (col0 >= p.N) break;  // warp-uniform
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) +
                          (uint32_t)(acc * BN + half * (BN / 2) + c * 32), r);
        // issue this chunk's global operand loads while the TMEM load is in flight (fetching
        // them one chunk ahead was measured SLOWER: +32..48 live registers spill at the 168
        // cap and the extra moves outweigh the hidden latency)
        const int col = col0 + cg * 4;
        const bool colok = col < p.N;  // N % 4 == 0 in vec mode
        bool ok[8];
        float4 res[8];
        uint2 dsrc[8];
        float4 b4n;
        float bsn;
        load_bias(c + 1 < NCH ? c + 1 : c, b4n, bsn);
        if constexpr (!(EPI & F_SCALAR)) {
#pragma unroll
          for (int it = 0; it < 8; ++it) ok[it] = colok && (row_base + it * 4 + rs) < p.M;
          if constexpr (EPI & F_RES) {
            if (p.residual && !tail_part) {
#pragma unroll
              for (int it = 0; it < 8; ++it)
                res[it] = ok[it] ? __ldg(reinterpret_cast<const float4*>(p.residual + (row0 + it * 4) * p.ldr + col))
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
          if constexpr (EPI & F_DACT) {
            if (p.dgelu_src) {
#pragma unroll
              for (int it = 0; it < 8; ++it)
                dsrc[it] = ok[it] ? __ldg(reinterpret_cast<const uint2*>(p.dgelu_src + (row0 + it * 4) * p.ldg + col))
                                  : make_uint2(0u, 0u);
            }
          }
        }
        tmem_ld_wait();
        if (c == NCH - 1 || col0 + 32 >= p.N) {
          // last TMEM read of this accumulator: hand it back to the MMA warp early
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (TWO) mbar_arrive_cluster(&tempty[acc], 0);  // the leader's MMA warp waits for both CTAs
            else mbar_arrive(&tempty[acc]);
          }
          released = true;
        }
        if constexpr (EPI & F_LSE) {
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
        }
        {  // phase 1: thread = row; 8 x STS.128, 16-byte chunk j of row r stored at j ^ (r & 7)
          float4* dst = reinterpret_cast<float4*>(epi + lane * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            dst[j ^ (lane & 7)] =
                make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                            __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
        }
        __syncwarp();
        if constexpr (EPI & F_TAIL) {
          if (tail_part) {  // raw accumulators -> workspace, 128 contiguous bytes per row
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int rl = it * 4 + rs;
              const float4 t = *reinterpret_cast<const float4*>(epi + rl * 32 + ((cg ^ (rl & 7)) << 2));
              __stcg(reinterpret_cast<float4*>(tail_blk + (size_t)rl * BN + c * 32 + cg * 4), t);
            }
            b4 = b4n;
            bs = bsn;
            __syncwarp();
            continue;
          }
        }
        if constexpr (!(EPI & F_SCALAR)) {
          // phase 2 (vector): 8 iterations of 4 rows x 32 columns; every global access is a full
          // 16-byte (fp32) / 8-byte (bf16) vector, 128 contiguous bytes per row.
          float v[8][4];
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rl = it * 4 + rs;
            const float4 t = *reinterpret_cast<const float4*>(epi + rl * 32 + ((cg ^ (rl & 7)) << 2));
            v[it][0] = t.x; v[it][1] = t.y; v[it][2] = t.z; v[it][3] = t.w;
          }
          if constexpr (EPI & F_TAIL) {
            if (tail_own) {
              for (int sl = 0; sl + 1 < p.tail_S; ++sl) {
                const float* src = tail_blk + (size_t)sl * TAIL_TILE + c * 32 + cg * 4;
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                  const float4 t = __ldcg(reinterpret_cast<const float4*>(src + (size_t)(it * 4 + rs) * BN));
                  v[it][0] += t.x; v[it][1] += t.y; v[it][2] += t.z; v[it][3] += t.w;
                }
              }
            }
          }
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            v[it][0] += b4.x; v[it][1] += b4.y; v[it][2] += b4.z; v[it][3] += b4.w;
          }
          bool act_done = false;
          if constexpr (EPI & F_OUT2) {
            if (p.out2) {
              const bool deriv = p.out2_mode == 1 && p.act == MMTG_ACT_GELU_NEW;
              if (deriv) {
                // gelu_new and its derivative from ONE tanh (c_fc forward: the activation goes to
                // the mlp c_proj, the derivative is kept for the backward multiply)
                act_done = true;
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                  float dv[4];
#pragma unroll
                  for (int e = 0; e < 4; ++e) gelu_new_both(v[it][e], v[it][e], dv[e]);
                  if (ok[it]) {
                    __nv_bfloat162 lo = __floats2bfloat162_rn(dv[0], dv[1]);
                    __nv_bfloat162 hi = __floats2bfloat162_rn(dv[2], dv[3]);
                    uint2 pk = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
                    *reinterpret_cast<uint2*>(p.out2 + (row0 + it * 4) * p.ldo2 + col) = pk;
                  }
                }
              } else {
#pragma unroll
                for (int it = 0; it < 8; ++it)
                  if (ok[it]) {
                    __nv_bfloat162 lo = __floats2bfloat162_rn(v[it][0], v[it][1]);
                    __nv_bfloat162 hi = __floats2bfloat162_rn(v[it][2], v[it][3]);
                    uint2 pk = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
                    *reinterpret_cast<uint2*>(p.out2 + (row0 + it * 4) * p.ldo2 + col) = pk;
                  }
              }
            }
          }
          if constexpr (EPI & F_ACT) {
            if (act_done) {
            } else if (p.act == MMTG_ACT_TANH) {
#pragma unroll
              for (int it = 0; it < 8; ++it)
#pragma unroll
                for (int e = 0; e < 4; ++e) v[it][e] = tanh_fast(v[it][e]);
            } else if (p.act == MMTG_ACT_GELU_NEW) {
#pragma unroll
              for (int it = 0; it < 8; ++it)
#pragma unroll
                for (int e = 0; e < 4; ++e) v[it][e] = gelu_new_fast(v[it][e]);
            }
          }
          if constexpr (EPI & F_DACT) {
            if (p.dgelu_src) {
#pragma unroll
              for (int it = 0; it < 8; ++it) {
                const float2 a = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&dsrc[it].x));
                const float2 b = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&dsrc[it].y));
                if (p.dact_tanh_out == 2) {
                  v[it][0] *= a.x; v[it][1] *= a.y; v[it][2] *= b.x; v[it][3] *= b.y;
                } else if (p.dact_tanh_out == 1) {
                  v[it][0] *= 1.f - a.x * a.x; v[it][1] *= 1.f - a.y * a.y;
                  v[it][2] *= 1.f - b.x * b.x; v[it][3] *= 1.f - b.y * b.y;
                } else {
                  v[it][0] *= dgelu_new_fast(a.x); v[it][1] *= dgelu_new_fast(a.y);
                  v[it][2] *= dgelu_new_fast(b.x); v[it][3] *= dgelu_new_fast(b.y);
                }
              }
            }
          }
          // inverted dropout of the value: before the residual add (resid_pdrop) ...
          auto apply_drop = [&]() {
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const uint32_t pair = (uint32_t)(((row0 + it * 4) * p.N + col) >> 1);
              const uint32_t b0 = drop_bits(dk, pair), b1 = drop_bits(dk, pair + 1);
              v[it][0] = drop_keep_lo(dk, b0) ? v[it][0] * dk.inv_keep : 0.f;
              v[it][1] = drop_keep_hi(dk, b0) ? v[it][1] * dk.inv_keep : 0.f;
              v[it][2] = drop_keep_lo(dk, b1) ? v[it][2] * dk.inv_keep : 0.f;
              v[it][3] = drop_keep_hi(dk, b1) ? v[it][3] * dk.inv_keep : 0.f;
            }
          };
          if constexpr ((EPI & F_DROP) && (EPI & F_RES)) {
            if (p.drop_p > 0.f) apply_drop();
          }
          if constexpr (EPI & F_RES) {
            if (p.residual) {
#pragma unroll
              for (int it = 0; it < 8; ++it) {
                v[it][0] += res[it].x; v[it][1] += res[it].y; v[it][2] += res[it].z; v[it][3] += res[it].w;
              }
            }
          }
          if constexpr (EPI & F_ROWTAB) {
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
          }
          // ... or of the embedding sum, after the row-table adds (embd_pdrop)
          if constexpr ((EPI & F_DROP) && !(EPI & F_RES)) {
            if (p.drop_p > 0.f) apply_drop();
          }
          if constexpr (EPI & F_COLSUM) {
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
          }
          bool do_atomic = false;
          if constexpr (EPI & F_ATOMIC) do_atomic = p.atomic != 0;
          if (do_atomic) {
            if constexpr (EPI & F_ATOMIC) {
#pragma unroll
              for (int it = 0; it < 8; ++it)
                if (ok[it]) {
                  float* dstp = (float*)p.out + (row0 + it * 4) * p.ldo + col;
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dstp),
                               "f"(v[it][0]), "f"(v[it][1]), "f"(v[it][2]), "f"(v[it][3])
                               : "memory");
                }
            }
          } else {
            if (p.out_bf16) {
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
          }
        } else {
          // phase 2 (scalar; unaligned pitches such as the contiguous [.., 13317] logits):
          // lane = column -> every store instruction writes 128 contiguous bytes of one row.
          const int scol = col0 + lane;
          const bool scolok = scol < p.N;
          const int nrows = min(32, p.M - row_base);  // warp-uniform, may be <= 0
          const int sw = lane >> 2, sl = lane & 3;    // swizzled position of this lane's column
          if constexpr (EPI & (F_ACT | F_RES | F_COLSUM | F_ATOMIC)) {
            float csum = 0.f;
            for (int rr = 0; rr < nrows; ++rr) {
              const long long row = row_base + rr;
              float v = epi[rr * 32 + ((sw ^ (rr & 7)) << 2) + sl] + bs;
              if (scolok) {
                if (p.act == MMTG_ACT_TANH) v = tanh_fast(v);
                else if (p.act == MMTG_ACT_GELU_NEW) v = gelu_new_fast(v);
                if (p.residual) v += __ldg(p.residual + row * p.ldr + scol);
                csum += v;
                if (p.atomic) atomicAdd((float*)p.out + row * p.ldo + scol, v);
                else if (p.out_bf16) ((bf16*)p.out)[row * p.ldo + scol] = __float2bfloat16(v);
                else ((float*)p.out)[row * p.ldo + scol] = v;
              }
            }
            if (p.colsum && scolok && nrows > 0) atomicAdd(p.colsum + scol, csum);
          } else {
            // lean path: store (+bias) only — the lm_head logits
            if (p.out_bf16) {
              bf16* orow = (bf16*)p.out + (long long)row_base * p.ldo + scol;
#pragma unroll
              for (int rr = 0; rr < 32; ++rr) {
                const float v = epi[rr * 32 + ((sw ^ (rr & 7)) << 2) + sl] + bs;
                if (rr < nrows && scolok) orow[(long long)rr * p.ldo] = __float2bfloat16(v);
              }
            } else {
              float* orow = (float*)p.out + (long long)row_base * p.ldo + scol;
#pragma unroll
              for (int rr = 0; rr < 32; ++rr) {
                const float v = epi[rr * 32 + ((sw ^ (rr & 7)) << 2) + sl] + bs;
                if (rr < nrows && scolok) orow[(long long)rr * p.ldo] = v;
              }
            }
          }
        }
        b4 = b4n;
        bs = bsn;
        __syncwarp();
      }
      if constexpr (EPI & F_TAIL) {
        if (tail_part) {  // this warp's share of the partial tile is written: release + count
          __syncwarp();
          if (lane == 0) {
            __threadfence();
            atomicAdd(p.tail_flags + (tile - p.tail_full), 1u);
          }
        } else if (tail_own) {  // the last of the pair's 16 owner warps re-arms the flags for the next launch
          __syncwarp();
          if (lane == 0) {
            unsigned int* done = p.tail_flags + 128 + (tile - p.tail_full);
            if (atomicAdd(done, 1u) == 15u) {
              *done = 0u;
              p.tail_flags[tile - p.tail_full] = 0u;
            }
          }
        }
      }
      if constexpr (EPI & F_LSE) {
        if (p.lse_partial) {
          // the two column halves of a 128/256-wide tile keep separate partial slots
          const long long row = row_base + lane;
          if (row < p.M) {
            float* dst = p.lse_partial + ((long long)(n_t * 2 + half) * p.M + row) * 2;
            dst[0] = run_max;
            dst[1] = run_sum;
          }
        }
      }
      if (!released) {
        // this half-tile lies entirely beyond N: nothing was read, still release the accumulator
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (TWO) mbar_arrive_cluster(&tempty[acc], 0);
          else mbar_arrive(&tempty[acc]);
        }
      }
      if (++acc == 2) {
        acc = 0;
        acc_ph ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer may still multicast into / arrive on this CTA until it is done
  if (warp == 1) {
    tc_fence_after();
    if (TWO) tmem_dealloc2(tmem_base, C::TMEM_COLS);
    else tmem_dealloc(tmem_base, C::TMEM_COLS);
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

// Tail-split workspace (partial accumulators + flags), one per (device, stream): GEMMs of one
// stream are serialised, so they can share it; launches on different streams never do.
struct TailWs {
  float* ws;
  unsigned int* flags;
};
constexpr size_t TAIL_WS_TILES = 80;                            // (tile, slice) partial tiles: <= clusters per grid
constexpr size_t TAIL_WS_BYTES = TAIL_WS_TILES * 2 * 128 * 256 * 4;  // 256 x 256 fp32 per pair tile
static bool tail_workspace(cudaStream_t st, TailWs* out) {
  static std::mutex mu;
  static std::unordered_map<unsigned long long, TailWs> map;
  const unsigned long long key = ((unsigned long long)(uintptr_t)st << 6) ^ (unsigned long long)current_device_index();
  std::lock_guard<std::mutex> g(mu);
  auto it = map.find(key);
  if (it != map.end()) {
    *out = it->second;
    return out->ws != nullptr;
  }
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
    cudaGetLastError();
    return false;  // no allocation inside a capture: this launch runs without the tail split
  }
  TailWs t{nullptr, nullptr};
  if (cudaMalloc(&t.ws, TAIL_WS_BYTES) != cudaSuccess || cudaMalloc(&t.flags, 256 * sizeof(unsigned int)) != cudaSuccess ||
      cudaMemset(t.flags, 0, 256 * sizeof(unsigned int)) != cudaSuccess) {
    cudaGetLastError();
    t.ws = nullptr;
  }
  map.emplace(key, t);
  *out = t;
  return t.ws != nullptr;
}

template <int BN, uint32_t EPI, bool TWO>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p,
                       cudaStream_t st) {
  using C = GemmCfg<BN, TWO>;
  MMTG_PER_DEVICE_FLAG(attr_set);
  if (!attr_set) {
    MMTG_CUDA_OK(cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN, EPI, TWO>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  const int total = (EPI & F_TAIL) ? p.tail_total : p.num_mp * p.num_n * p.splits;  // work units of a CTA pair
  // MMTG_GEMM_CTAS caps the persistent grid (data-parallel runs leave a few SMs to the NCCL
  // kernels: a persistent CTA pair that cannot become resident until an all-reduce kernel exits
  // delays its statically assigned tiles, and with them the whole GEMM)
  static const int cta_cap = []() {
    const char* e = getenv("MMTG_GEMM_CTAS");
    const int v = e ? atoi(e) : 0;
    return v >= 2 ? v : 0;
  }();
  int max_clusters = num_sms() / 2;
  if (cta_cap && cta_cap / 2 < max_clusters) max_clusters = cta_cap / 2;
  if (p.grid_mode == 1) max_clusters = total;  // one work unit per CTA pair
  // dynamic scheduling (cluster launch control): one cluster per work unit in the grid; the
  // resident clusters cancel and absorb the ones not launched yet (see TileSched)
  // MEASURED (round 2): correct, but not faster - train step 8.50 ms vs 8.35 ms static on 1 GPU,
  // 8.83 vs 8.73 ms on 2 GPUs (NCCL all-reduce kernels competing for SMs). Opt-in: MMTG_GEMM_CLC=1.
  static const bool clc_on = []() {
    const char* e = getenv("MMTG_GEMM_CLC");
    return e && e[0] == '1';
  }();
  GemmParams pl = p;
  pl.clc = (clc_on && p.grid_mode == 0 && !cta_cap && !(EPI & F_TAIL)) ? 1 : 0;
  if (pl.clc) max_clusters = total;
  const int grid = 2 * (total < max_clusters ? total : max_clusters);
  ProfScope prof(0, 2.0 * p.M * p.N * p.K,
                 2.0 * ((double)p.M * p.K + (double)p.N * p.K) + (p.out_bf16 ? 2.0 : 4.0) * p.M * p.N, st);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(C::THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = st;
  // Programmatic dependent launch (the next GEMM's CTAs are scheduled while this one drains and run
  // their TMEM / barrier prologue early). MEASURED in round 2 on the full train step: 8.25 ms with it,
  // 8.00-8.03 ms without (same box, two runs each) - the early CTAs hold whole SMs (227 KB of shared
  // memory) while they wait, which the side stream's weight-gradient GEMMs could have used. Opt-in:
  // MMTG_GEMM_PDL=1. (The griddepcontrol instructions in the kernel are no-ops without the attribute.)
  static const bool use_pdl = []() {
    const char* e = getenv("MMTG_GEMM_PDL");
    return e && e[0] == '1';
  }();
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl ? 2 : 1;
  MMTG_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_bf16_tcgen05_kernel<BN, EPI, TWO>, tmA, tmB, pl));
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
    if (a->N <= 128) BN = 128;
  }
  MMTG_CHECK_ARG(BN == 128 || BN == 192 || BN == 256, "block_n must be 0, 128, 192 or 256");
  const bool want192 = BN == 192;
  if (want192) BN = 256;  // decided below, once the epilogue features are known
  const bool atomic = a->accumulate != 0 || a->split_k > 1;
  MMTG_CHECK_ARG(!(atomic && a->out_dtype != MMTG_F32), "accumulate/split_k need an fp32 out");
  MMTG_CHECK_ARG(!(a->rowtab0 && !a->rowidx0 && a->rowmod0 <= 0), "rowtab0 needs rowidx0 or rowmod0");
  MMTG_CHECK_ARG(!(a->rowtab1 && !a->rowidx1), "rowtab1 needs rowidx1");

  GemmParams p;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.num_m = cdiv(a->M, 128);
  p.num_mp = (p.num_m + 1) / 2;
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
  p.drop_seed = (const unsigned long long*)a->drop_seed; p.drop_site = a->drop_site;
  p.drop_p = a->drop_seed ? a->drop_p : 0.f;
  p.grid_mode = a->grid_mode;
  p.out2_mode = a->out2_mode;
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

  static const bool two_sm = []() {  // MMTG_GEMM_2SM=0 selects the 1-SM (multicast) kernel
    const char* e = getenv("MMTG_GEMM_2SM");
    return !(e && e[0] == '0');
  }();
  {
    // 192-wide tiles for plain K-major-B GEMMs whose tile count quantises badly on the 74 CTA pairs:
    // the decoder's N = 768 dgrads are 90 pairs of 256 columns (two rounds for 1.22 rounds of work)
    // or 120 pairs of 192 columns (two rounds of 3/4 the size).
    // MEASURED (round 2): no gain - 7552 x 768 x {768, 2304, 3072}: 21.4 / 35.8 / 43.9 us against
    // 20.5 / 35.8 / 45.1 us with 256-wide tiles (cuBLAS: 17.4 / 27.6 / 33.6 us, stream-K): the
    // narrower pair tile needs 28 KB of operands per 384 MMA cycles (73 B/clk/SM, above the
    // ~64 B/clk the L2 delivers), so each of the smaller tiles runs slower. Explicit block_n = 192
    // or MMTG_GEMM_BN192=1 select it.
    static const bool bn192_on = []() {
      const char* e = getenv("MMTG_GEMM_BN192");
      return e && e[0] == '1';
    }();
    const bool plain = !p.out2 && p.act == MMTG_ACT_NONE && !p.dgelu_src && !p.residual && !p.rowtab0 && !p.rowtab1 &&
                       !p.colsum && !p.atomic && !p.lse_partial && p.vec4 && !(a->drop_seed && a->drop_p > 0.f) &&
                       p.splits == 1 && !p.b_mn && two_sm;
    MMTG_CHECK_ARG(!want192 || plain, "block_n 192 serves plain (no epilogue features) GEMMs with a K-major B operand");
    if (plain && BN == 256 && (want192 || (bn192_on && a->block_n == 0))) {
      const int Cn = num_sms() / 2;
      const long long c256 = (long long)cdiv(p.num_mp * cdiv(a->N, 256), Cn) * 256;
      const long long c192 = (long long)cdiv(p.num_mp * cdiv(a->N, 192), Cn) * 192;
      if (want192 || c192 < c256) {
        BN = 192;
        p.num_n = cdiv(a->N, 192);
      }
    }
  }
  CUtensorMap tmA, tmB;
  if (!p.a_mn) MMTG_TRY(make_tmap_bf16_2d(&tmA, a->A, a->K, a->M, a->lda, 64, 128));
  else         MMTG_TRY(make_tmap_bf16_2d(&tmA, a->A, a->M, a->K, a->lda, 64, 64));
  if (!p.b_mn) MMTG_TRY(make_tmap_bf16_2d(&tmB, a->B, a->K, a->N, a->ldb, 64, BN / 2));
  else         MMTG_TRY(make_tmap_bf16_2d(&tmB, a->B, a->N, a->K, a->ldb, 64, 64));

  cudaStream_t st = (cudaStream_t)stream;
  // pick the smallest instantiated epilogue that covers the requested features
  uint32_t need = 0;
  if (p.out2) need |= F_OUT2;
  if (p.act != MMTG_ACT_NONE) need |= F_ACT;
  if (p.dgelu_src) need |= F_DACT;
  if (p.residual) need |= F_RES;
  if (p.rowtab0 || p.rowtab1) need |= F_ROWTAB;
  if (p.colsum) need |= F_COLSUM;
  if (p.atomic) need |= F_ATOMIC;
  if (p.lse_partial) need |= F_LSE;
  if (!p.vec4) need |= F_SCALAR;
  if (p.drop_p > 0.f) {
    MMTG_CHECK_ARG(p.drop_p < 1.f && p.vec4 && a->N % 4 == 0 && !p.atomic && (p.residual || p.rowtab0 || p.rowtab1),
                   "dropout epilogue: needs p < 1, N %% 4 == 0, aligned operands, and a residual or row-table epilogue");
    need |= F_DROP;
  }
  // Tail split: with U tile pairs on C clusters the last, partial round leaves C - U % C clusters
  // idle for a whole tile time (the N = 768 GEMMs of the decoder: 90 pairs on 74 clusters = two
  // rounds for 1.22 rounds of work). The T = U % C tiles of that round are cut into S = C / T
  // k-slices, one per cluster; slice 0 adds the other slices' fp32 partials (parked in an L2-resident
  // workspace) in slice order and runs the epilogue. Deterministic, no atomics on the output.
  // MEASURED (round 2, B200): correct and bit-reproducible, but SLOWER than the plain two rounds -
  // 7552 x 768 x 3072 46.0 us vs 43.8 us, train step 8.69 ms vs 8.25 ms: slice 0 reads the three
  // partial tiles (384 KB per CTA) as 12 dependent L2 round trips per warp, which costs more than
  // the 0.75 tile time the split saves. Opt-in (MMTG_GEMM_TAIL=1) until the fix-up is pipelined.
  static const bool tail_on = []() {
    const char* e = getenv("MMTG_GEMM_TAIL");
    return e && e[0] == '1';
  }();
  p.tail_full = p.tail_S = p.tail_kbs = p.tail_total = 0;
  p.tail_ws = nullptr;
  p.tail_flags = nullptr;
  if (tail_on && two_sm && BN == 256 && !p.atomic && p.splits == 1 && p.vec4 && p.grid_mode == 0 &&
      (need & ~(uint32_t)(F_RES | F_DROP)) == 0 && !getenv("MMTG_GEMM_CTAS")) {
    const int U = p.num_mp * p.num_n, Cn = num_sms() / 2;
    const int T = U % Cn;
    const int S = T > 0 ? Cn / T : 0;
    if (U > Cn && S >= 2 && p.num_kb >= 2 * S && (size_t)T * (S - 1) <= TAIL_WS_TILES && T <= 128) {
      TailWs tw;
      if (tail_workspace(st, &tw)) {
        p.tail_full = U - T;
        p.tail_S = S;
        p.tail_kbs = cdiv(p.num_kb, S);
        p.tail_S = cdiv(p.num_kb, p.tail_kbs);  // (slices that would be empty are dropped)
        p.tail_total = p.tail_full + T * p.tail_S;
        p.tail_ws = tw.ws;
        p.tail_flags = tw.flags;
        need |= F_TAIL;
      }
    }
  }
#define MMTG_TRY_EPI(MASK)                                                        \
  if ((need & ~(uint32_t)(MASK)) == 0) {                                          \
    if (two_sm) {                                                                 \
      if (BN == 256) return launch_gemm<256, (MASK), true>(tmA, tmB, p, st);      \
      return launch_gemm<128, (MASK), true>(tmA, tmB, p, st);                     \
    }                                                                             \
    if (BN == 256) return launch_gemm<256, (MASK), false>(tmA, tmB, p, st);       \
    return launch_gemm<128, (MASK), false>(tmA, tmB, p, st);                      \
  }
  if (BN == 192) return launch_gemm<192, 0u, true>(tmA, tmB, p, st);
  if (need & F_TAIL) {
    if ((need & ~(uint32_t)F_TAIL) == 0) return launch_gemm<256, F_TAIL, true>(tmA, tmB, p, st);
    return launch_gemm<256, F_TAIL | F_RES | F_DROP, true>(tmA, tmB, p, st);
  }
  MMTG_TRY_EPI(0u)
  MMTG_TRY_EPI(F_ACT | F_OUT2)
  MMTG_TRY_EPI(F_RES)
  MMTG_TRY_EPI(F_RES | F_DROP)
  MMTG_TRY_EPI(F_DACT | F_COLSUM)
  MMTG_TRY_EPI(F_ATOMIC)
  MMTG_TRY_EPI(F_ROWTAB)
  MMTG_TRY_EPI(F_ROWTAB | F_DROP)
  MMTG_TRY_EPI(F_SCALAR | F_LSE)
  MMTG_TRY_EPI(F_SCALAR | F_ACT | F_RES | F_COLSUM | F_ATOMIC)
  MMTG_TRY_EPI(F_OUT2 | F_ACT | F_DACT | F_RES | F_ROWTAB | F_COLSUM | F_ATOMIC | F_LSE)
#undef MMTG_TRY_EPI
  set_last_error("no GEMM epilogue instantiation covers feature mask 0x%x", need);
  return -1;
}
