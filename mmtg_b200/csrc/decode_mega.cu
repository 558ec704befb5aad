// Persistent decode megakernel: all GPT-2 blocks + ln_f + lm_head of ONE decode position in a
// single launch (generation at batch <= 64 is latency-bound: the per-op path pays ~90 launches
// per position). One 256-thread CTA per SM stays resident (cooperative launch, thread-block
// clusters of 4) and a grid-wide barrier separates five phases per block:
//   A: qkv16 = bf16(LN1(h) W_attn + b)         (LayerNorm folded, see below; q pre-scaled by 1/8)
//   B: cached attention, one warp per (row, head): append K/V, softmax(q K^T) V on tensor cores
//      over TMA-staged 64-key boxes -> att16
//   C: h2 = h + att16 W_proj + b               (+ bf16 copy + LN2 row-statistics partials)
//   D: u16 = bf16(gelu(LN2(h2) W_fc + b))
//   E: h = h2 + u16 W_proj2 + b                (+ bf16 copy + LN1 / ln_f row-statistics partials)
// and finally F: logits = ln_f(h) wte^T.
//
// DETERMINISTIC split-K. Every GEMM phase is cut into (column chunk) x (4 k-slices): the four
// CTAs of a CLUSTER own the four k-slices of one column chunk, park their fp32 partial tiles in
// their own shared memory and, after a cluster barrier, each CTA reduces 16 of the 64 rows over
// the four partials through distributed shared memory IN A FIXED ORDER and applies the epilogue.
// There are no atomics and no partial sums in global memory: two runs on the same inputs give
// bit-identical logits (round 1 combined split-K partials with red.global.add: run-to-run
// differences of a few 1e-3, and ~3.5 us per phase spent draining the atomics at the barrier).
// Phase outputs are FINAL values, written once as the bf16 operand the next phase consumes
// (activations cross L2 as bf16, half of round 1's fp32 accumulators) plus the fp32 residual
// stream where one is needed.
//
// LayerNorm is FOLDED into the following linear layer (mmtg_decode_fold_weights): W' = g ⊙ W,
// cs = column sums of W', b' = bias + beta · W, so LN(x) W + bias = rstd (x W' - mean cs) + b'.
// The GEMM consumes raw bf16(x); the producer of x (phase C / E epilogue) emits per-row partial
// (sum, sum of squares) per column chunk, and the consumer's epilogue combines them in a fixed
// order into (mean, rstd).
//
// Weights stream by cp.async straight from the [K, N] matrices into padded (bank-conflict-free)
// shared-memory slots, issued one to four phases AHEAD of their use inside the barrier windows,
// and complete on mbarriers (cp.async.mbarrier.arrive), so the stream runs across phase barriers.
// mma.sync m16n8k16, ldmatrix on padded rows (row stride / 16 B odd).
#include <string.h>

#include "../../include/mmtg_b200.h"
#include "ops.h"
#include "mma_tiles.cuh"
#include "sampler.cuh"

namespace mmtg {

void count_launch(int n = 1);
int make_tmap_bf16_2d(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld,
                      uint32_t box0, uint32_t box1);

namespace {

constexpr int MG_THREADS = 256;
constexpr int CL = 4;          // cluster size = split-K ways
constexpr int ATT_WARPS = 6;   // attention: 2 x (8 KB K + 8 KB V) per warp = 192 KB
// columns per cluster (chunk width) of each GEMM phase; N / NC clusters take part
constexpr int NC_A = 72, NC_C = 24, NC_D = 96, NC_E = 32;
constexpr int KS_S = 192, KS_L = 768;  // k-slice per CTA: K / 4 for K = 768 and K = 3072
constexpr int NC_P = 24, KS_P = 128;   // projector layer 2 (K = 512 -> N = 768): 32 clusters

// padded shared-memory strides: (stride / 16) odd -> the 8 rows of an ldmatrix hit 8 distinct
// 16-byte bank groups
__host__ __device__ constexpr int w_stride(int nc) { return ((nc * 2 / 16) % 2 == 1) ? nc * 2 : nc * 2 + 16; }
__host__ __device__ constexpr int a_stride(int ksl) { return ksl * 2 + 16; }
// fp32 partial-tile pitch (floats): == 8 mod 32, so the 8 rows of a fragment store land in
// distinct 32-byte segments
__host__ __device__ constexpr int part_pitch(int nc) { return nc + ((8 - nc % 32) + 32) % 32; }

// shared-memory map. The attention boxes (6 x 32 KB) alias the activation / W_fc / W_proj2 slots,
// which are dead during phase B; the W_attn / W_proj slot lies above them and is live across B.
constexpr int OFF_ACTS = 0;                                  // [64][a_stride(768)]  (also the fp32 partial tile)
constexpr int SZ_ACTS = 64 * a_stride(KS_L);                 // 99,328
constexpr int OFF_WD = OFF_ACTS + SZ_ACTS;                   // [192][w_stride(96)]
constexpr int SZ_WD = KS_S * w_stride(NC_D);                 // 39,936
constexpr int OFF_WE = OFF_WD + SZ_WD;                       // [768][w_stride(32)]
constexpr int SZ_WE = KS_L * w_stride(NC_E);                 // 61,440
constexpr int OFF_WAC = OFF_WE + SZ_WE;                      // [192][w_stride(72)] / [192][w_stride(24)]
constexpr int SZ_WAC = KS_S * w_stride(NC_A);                // 27,648
constexpr int OFF_STATS = OFF_WAC + SZ_WAC;                  // [64][2] floats
constexpr int OFF_BARS = OFF_STATS + 512;
constexpr int MG_SMEM = OFF_BARS + 256;
static_assert(ATT_WARPS * 32768 <= OFF_WAC, "attention boxes must not reach the live W_attn / W_proj slot");
static_assert(MG_SMEM <= 232448, "shared memory budget (227 KB)");
static_assert(2 * 12 * 4096 <= OFF_WAC - OFF_WD, "lm_head weight buffers");
// full-step mode: the sampler's working logits + scratch and the projector weight slot share the
// activation region (dead between the lm_head phase and the next position's first GEMM)
constexpr int OFF_SAMP_SCR = 55296;  // >= 13317 * 4
constexpr int OFF_WP = 65536;        // [128][w_stride(24)] = 6,144 B
static_assert(OFF_SAMP_SCR + (int)sizeof(SamplerScratch) <= OFF_WP && OFF_WP + KS_P * w_stride(NC_P) <= SZ_ACTS, "sampler / projector slots");
static_assert(64 * a_stride(KS_P) <= OFF_SAMP_SCR, "projector activations");
constexpr int HEAD_BUF = 12 * 4096;  // one 32-row unit of wte': 12 k-tiles of [32][128 B]

enum { BAR_ACTS = 0, BAR_WAC = 1, BAR_WD = 2, BAR_WE = 3, BAR_H0 = 4, BAR_H1 = 5, BAR_WP = 6, BAR_ATT = 7 };

struct MegaParams {
  const float* P;
  const bf16* W;
  mmtg_param_offsets off;  // by value (~5 KB of kernel parameters)
  CUtensorMap tm_k, tm_v;  // [NL*B*NH*Lmax, 64] bf16, 64x64 boxes, 128B swizzle
  int B, E, NH, NL, V, Pl, Lmax;
  MegaBufs w;
  int* j_ptr;
  float* logits;
  unsigned long long* trace;  // optional: CTA 0 stamps globaltimer after every phase
  // full-step mode (mmtg_decode_steps_fused): embedding + projector prologue and the sampler run
  // inside the kernel, for n_steps consecutive positions in ONE launch
  int full, n_steps;
  int* gen;
  int gen_ld, sent_len, n_sent, S, He, table_rows;
  float temperature, top_p, rep_penalty;
  int top_k;
  const unsigned long long* seed_dev;
};

// Grid barrier split into arrive / wait: work that does not depend on the other CTAs' results of
// the finished phase (weight prefetch, K/V box prefetch) runs between the two, inside the
// barrier's latency.
__device__ __forceinline__ void barrier_arrive(unsigned int* counter) {
  __syncthreads();
  if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
}
__device__ __forceinline__ void barrier_wait(unsigned int* counter, unsigned int target) {
  if (threadIdx.x == 0) {
    unsigned int spins = 0;
    while (true) {
      unsigned int v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      if ((int)(v - target) >= 0) break;  // the counter runs on across launches (wrap-safe compare)
      if (++spins > (1u << 24)) {
        printf("mmtg: decode grid barrier timed out (block %d, target %u, saw %u)\n", blockIdx.x, target, v);
        __trap();
      }
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void substamp(unsigned long long* sub, int k) {
  if (sub && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    sub[k] = t;
  }
}

// arrive on `bar` once every cp.async this thread issued so far has landed
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ float2 ld_dsmem_f2(uint32_t local_addr, uint32_t rank) {
  uint32_t ra;
  float2 t;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(rank));
  asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(t.x), "=f"(t.y) : "r"(ra));
  return t;
}

// ---- weights: rows [rank*KSL, +KSL) x columns [chunk*NC, +NC) of a [K, ldw] bf16 matrix ----
// One bulk async copy (TMA, 1-D) per row segment, completing by byte count on the slot's own
// mbarrier: independent of the cp.async traffic of the activations, so a phase never waits for
// weights that were prefetched for a LATER phase.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
template <int KSL, int NC>
__device__ __forceinline__ void issue_w(const bf16* __restrict__ Wm, long long ldw, int chunk, int rank, uint8_t* slot,
                                        uint64_t* bar) {
  constexpr int SW = w_stride(NC);
  static_assert((NC * 2) % 16 == 0, "row segments must be multiples of 16 bytes");
  const bf16* src = Wm + (long long)rank * KSL * ldw + chunk * NC;
  if (threadIdx.x == 0) mbar_arrive_expect_tx(bar, (uint32_t)(KSL * NC * 2));
  fence_proxy_async();  // the slot's previous contents were read through the generic proxy
  for (int r = threadIdx.x; r < KSL; r += MG_THREADS) bulk_g2s(slot + r * SW, src + (long long)r * ldw, NC * 2, bar);
}
// ---- activations: rows 0..63 x columns [k0, k0+KSL) of a [*, lda] bf16 matrix (rows >= B zero) ----
template <int KSL>
__device__ __forceinline__ void issue_acts(const bf16* __restrict__ A, long long lda, int k0, int B, uint8_t* acts,
                                           uint64_t* bar) {
  constexpr int SA = a_stride(KSL), CH = KSL / 8;
  for (int idx = threadIdx.x; idx < 64 * CH; idx += MG_THREADS) {
    const int r = idx / CH, c = idx - r * CH;
    const bool ok = r < B;
    cp_async16(acts + r * SA + c * 16, A + (long long)(ok ? r : 0) * lda + k0 + c * 8, ok);
  }
  cp_async_arrive(bar);
}

// (mean, rstd) of `nrows` rows starting at row0 from the per-chunk partial sums the producer of
// the rows wrote: stats[row][chunk] = (sum, sum of squares), combined in chunk order
__device__ __forceinline__ void rows_mean_rstd(const float* __restrict__ part, int nch, int row0, int nrows, int E,
                                               float* out) {
  if ((int)threadIdx.x < nrows) {
    const int row = row0 + threadIdx.x;
    float s1 = 0.f, s2 = 0.f;
    for (int c = 0; c < nch; ++c) {
      const float2 v = __ldcg(reinterpret_cast<const float2*>(part + ((long long)row * 32 + c) * 2));
      s1 += v.x;
      s2 += v.y;
    }
    const float mean = s1 / (float)E;
    const float var = fmaxf(s2 / (float)E - mean * mean, 0.f);
    out[threadIdx.x * 2] = mean;
    out[threadIdx.x * 2 + 1] = rsqrtf(var + 1e-5f);
  }
}

// One GEMM phase for the CTA `rank` of cluster `chunk`: partial = acts[64 x KSL] * W[KSL x NC],
// cluster reduction of rows [16 rank, +16) in rank order, epilogue `epi(row, col, v0, v1)` per
// column pair and `epi.row_done(row, lane16)` once per row (16 lanes per row).
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// TAIL_SYNC: a second cluster barrier after the reduction. Only needed when this CTA's partial
// tile can be overwritten before the next GRID barrier completes (phase A: the K/V boxes of the
// attention phase are issued inside the barrier window and land on the same shared memory);
// everywhere else the grid barrier orders the peers' reads before the next write.
template <int KSL, int NC, bool TAIL_SYNC, class Epi, bool STAGED = false>
__device__ __forceinline__ void gemm_phase(const bf16* __restrict__ A, long long lda, int B, int rank, uint8_t* smem,
                                           const uint8_t* wslot, uint64_t* bars, int wbar, uint32_t& par, Epi& epi,
                                           unsigned long long* sub = nullptr) {
  constexpr int SA = a_stride(KSL), SW = w_stride(NC), NB = NC / 8, NB0 = (NB + 1) / 2, PITCH = part_pitch(NC);
  static_assert(64 * PITCH * 4 <= SZ_ACTS, "partial tile must fit the activation region");
  const int tid = threadIdx.x, warp = tid >> 5, l = lane_id();
  uint8_t* acts = smem + OFF_ACTS;
  substamp(sub, 0);
  if (STAGED) {  // the caller wrote the activation tile itself
    __syncthreads();
  } else {
    issue_acts<KSL>(A, lda, rank * KSL, B, acts, &bars[BAR_ACTS]);
    mbar_wait<21>(&bars[BAR_ACTS], (par >> BAR_ACTS) & 1u);
    par ^= 1u << BAR_ACTS;
  }
  substamp(sub, 1);
  mbar_wait<22>(&bars[wbar], (par >> wbar) & 1u);
  par ^= 1u << wbar;
  substamp(sub, 2);
  const int mt = warp & 3, nh = warp >> 2;
  const int nb_lo = nh ? NB0 : 0, nb_hi = nh ? NB : NB0;
  float acc[NB0][4];
#pragma unroll
  for (int i = 0; i < NB0; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const uint32_t a_base = smem_u32(acts) + (uint32_t)((mt * 16 + (l & 15)) * SA + (l >> 4) * 16);
  const uint32_t w_base = smem_u32(wslot) + (uint32_t)(((l & 7) + ((l >> 3) & 1) * 8) * SW);
#pragma unroll 4
  for (int kk = 0; kk < KSL / 16; ++kk) {
    uint32_t a0, a1, a2, a3;
    ldsm_x4(a_base + kk * 32, a0, a1, a2, a3);
    const uint32_t wk = w_base + (uint32_t)(kk * 16 * SW);
#pragma unroll
    for (int j = 0; j < NB0; j += 2) {
      const int nb = nb_lo + j;
      if (nb + 1 < nb_hi) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(wk + (uint32_t)((nb + (l >> 4)) * 16), b0, b1, b2, b3);
        mma16816(acc[j], a0, a1, a2, a3, b0, b1);
        if (j + 1 < NB0) mma16816(acc[j + 1], a0, a1, a2, a3, b2, b3);
      } else if (nb < nb_hi) {
        uint32_t b0, b1;
        ldsm_x2_t(wk + (uint32_t)(nb * 16), b0, b1);
        mma16816(acc[j], a0, a1, a2, a3, b0, b1);
      }
    }
  }
  __syncthreads();  // every warp is done reading the activations: the partial tile aliases them
  substamp(sub, 3);
  float* part = reinterpret_cast<float*>(acts);
#pragma unroll
  for (int j = 0; j < NB0; ++j) {
    const int nb = nb_lo + j;
    if (nb < nb_hi) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int row = mt * 16 + (l >> 2) + r * 8;
        *reinterpret_cast<float2*>(part + row * PITCH + nb * 8 + (l & 3) * 2) = make_float2(acc[j][2 * r], acc[j][2 * r + 1]);
      }
    }
  }
  cluster_arrive();
  // rows [16 rank, +16): two rows per warp, 16 lanes per row, column pairs lane16 + 16 i.
  // The epilogue's global operands are fetched while the cluster barrier completes.
  {
    constexpr int NI = (NC / 2 + 15) / 16;
    const int row = rank * 16 + warp * 2 + (l >> 4), l16 = l & 15;
    const uint32_t pbase = smem_u32(part) + (uint32_t)(row * PITCH * 4);
    typename Epi::Pre pre[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int pp = l16 + 16 * i;
      if (row < B && pp < NC / 2) pre[i] = epi.load(row, 2 * pp);
    }
    cluster_wait();
    substamp(sub, 4);
    if (row < B) {
      float2 v[NI][CL];
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int pp = l16 + 16 * i;
        if (pp < NC / 2) {
#pragma unroll
          for (int rk = 0; rk < CL; ++rk) v[i][rk] = ld_dsmem_f2(pbase + pp * 8, rk);
        }
      }
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int pp = l16 + 16 * i;
        if (pp < NC / 2) {
          float2 a = v[i][0];
#pragma unroll
          for (int rk = 1; rk < CL; ++rk) {  // fixed order: rank 0, 1, 2, 3
            a.x += v[i][rk].x;
            a.y += v[i][rk].y;
          }
          epi.apply(row, 2 * pp, a.x, a.y, pre[i]);
        }
      }
    }
    epi.row_done(row, l16, row < B);
  }
  substamp(sub, 5);
  if (TAIL_SYNC) cluster_sync_all();  // peers may still be reading this CTA's partial tile
  substamp(sub, 6);
}

// ---- epilogues (col = column inside the chunk; n0 = first column of the chunk) ----
struct Pre2 {
  float2 a, b;
};
struct EpiQKV {  // qkv16 = bf16(rstd (acc - mean cs) + b'), q pre-scaled by 1/8
  typedef Pre2 Pre;
  const float *cs, *bq, *stats;  // stats: smem (mean, rstd) of this CTA's 16 rows
  bf16* out;
  int n0, E, row0;
  __device__ __forceinline__ Pre load(int, int col) const {
    const int n = n0 + col;
    return Pre{__ldg(reinterpret_cast<const float2*>(cs + n)), __ldg(reinterpret_cast<const float2*>(bq + n))};
  }
  __device__ __forceinline__ void apply(int row, int col, float v0, float v1, const Pre& q) {
    const int n = n0 + col;
    const float mean = stats[(row - row0) * 2], rstd = stats[(row - row0) * 2 + 1];
    const float sc = n < E ? 0.125f : 1.f;
    const float o0 = (rstd * (v0 - mean * q.a.x) + q.b.x) * sc, o1 = (rstd * (v1 - mean * q.a.y) + q.b.y) * sc;
    *reinterpret_cast<uint32_t*>(out + (long long)row * 3 * E + n) = pack_bf16(o0, o1);
  }
  __device__ __forceinline__ void row_done(int, int, bool) {}
};
struct EpiResid {  // x = resid + bias + acc -> fp32 + bf16 copies, per-row (sum, sumsq) partial of this chunk
  typedef Pre2 Pre;
  const float *resid, *bias;
  float* out32;
  bf16* out16;
  float* stats_part;  // [64][32][2]
  int n0, E, chunk;
  float s1, s2;
  __device__ __forceinline__ Pre load(int row, int col) const {
    const int n = n0 + col;
    return Pre{__ldcg(reinterpret_cast<const float2*>(resid + (long long)row * E + n)),
               __ldg(reinterpret_cast<const float2*>(bias + n))};
  }
  __device__ __forceinline__ void apply(int row, int col, float v0, float v1, const Pre& q) {
    const int n = n0 + col;
    const float o0 = q.a.x + q.b.x + v0, o1 = q.a.y + q.b.y + v1;
    *reinterpret_cast<float2*>(out32 + (long long)row * E + n) = make_float2(o0, o1);
    *reinterpret_cast<uint32_t*>(out16 + (long long)row * E + n) = pack_bf16(o0, o1);
    s1 += o0 + o1;
    s2 += o0 * o0 + o1 * o1;
  }
  __device__ __forceinline__ void row_done(int row, int l16, bool valid) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {  // fixed tree over the 16 lanes of the row
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (valid && l16 == 0) *reinterpret_cast<float2*>(stats_part + ((long long)row * 32 + chunk) * 2) = make_float2(s1, s2);
    s1 = s2 = 0.f;
  }
};
struct Pre3 {
  float2 a, b, c;
};
struct EpiEmbed {  // h = acc + b2 + wpe[pos] + wte[type(row)] -> fp32 + bf16 + LN1 (sum, sumsq) partial
  typedef Pre3 Pre;
  const float *bias, *wpe_row, *wte;
  const int* types;  // smem: token-type id of each of the 64 rows
  float* out32;
  bf16* out16;
  float* stats_part;
  int n0, E, chunk;
  float s1, s2;
  __device__ __forceinline__ Pre load(int row, int col) const {
    const int n = n0 + col;
    return Pre{__ldg(reinterpret_cast<const float2*>(bias + n)), __ldg(reinterpret_cast<const float2*>(wpe_row + n)),
               __ldg(reinterpret_cast<const float2*>(wte + (long long)types[row] * E + n))};
  }
  __device__ __forceinline__ void apply(int row, int col, float v0, float v1, const Pre& q) {
    const int n = n0 + col;
    const float o0 = v0 + q.a.x + q.b.x + q.c.x, o1 = v1 + q.a.y + q.b.y + q.c.y;
    *reinterpret_cast<float2*>(out32 + (long long)row * E + n) = make_float2(o0, o1);
    *reinterpret_cast<uint32_t*>(out16 + (long long)row * E + n) = pack_bf16(o0, o1);
    s1 += o0 + o1;
    s2 += o0 * o0 + o1 * o1;
  }
  __device__ __forceinline__ void row_done(int row, int l16, bool valid) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (valid && l16 == 0) *reinterpret_cast<float2*>(stats_part + ((long long)row * 32 + chunk) * 2) = make_float2(s1, s2);
    s1 = s2 = 0.f;
  }
};
struct EpiGelu {  // u16 = bf16(gelu_new(rstd (acc - mean cs) + b'))
  typedef Pre2 Pre;
  const float *cs, *bq, *stats;
  bf16* out;
  int n0, N, row0;
  __device__ __forceinline__ Pre load(int, int col) const {
    const int n = n0 + col;
    return Pre{__ldg(reinterpret_cast<const float2*>(cs + n)), __ldg(reinterpret_cast<const float2*>(bq + n))};
  }
  __device__ __forceinline__ void apply(int row, int col, float v0, float v1, const Pre& q) {
    const int n = n0 + col;
    const float mean = stats[(row - row0) * 2], rstd = stats[(row - row0) * 2 + 1];
    const float o0 = gelu_new_fast(rstd * (v0 - mean * q.a.x) + q.b.x), o1 = gelu_new_fast(rstd * (v1 - mean * q.a.y) + q.b.y);
    *reinterpret_cast<uint32_t*>(out + (long long)row * N + n) = pack_bf16(o0, o1);
  }
  __device__ __forceinline__ void row_done(int, int, bool) {}
};

// ---------------------------------------------------------------------------------------------
// Cached attention: one warp per (row, head). 64-key boxes of the K and V cache arrive through
// TMA (128B swizzle, two buffers per warp, the first two issued one phase ahead) so ~190 KB per
// SM is in flight: at late positions the KV read is the largest HBM stream of the step. The
// single query row is row 0 of an m16n8k16 A operand (lanes 0-3 hold it, other rows are zero):
// 32 + 32 tensor-core instructions per box instead of ~1000 scalar ones. Online softmax across
// boxes; keys past `pos` and padded keys are masked by selection, never by arithmetic.
// ---------------------------------------------------------------------------------------------
// generic-proxy writes (cache rows appended by earlier positions of this launch, shared memory read
// by ldmatrix) ordered before the async-proxy (TMA) accesses that follow, for every state space
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void attn_issue(const MegaParams& p, int layer, int u, int c, uint8_t* wbuf, uint64_t* bars) {
  const int row = ((layer * p.B * p.NH + u) * p.Lmax) + c * 64;
  uint8_t* buf = wbuf + (c & 1) * 16384;
  mbar_arrive_expect_tx(&bars[c & 1], 16384u);
  tma_load_2d(buf, &p.tm_k, &bars[c & 1], 0, row);
  tma_load_2d(buf + 8192, &p.tm_v, &bars[c & 1], 0, row);
}

// lane 0: pull the K/V boxes 2.. of a unit (the ones that do not fit the two staged buffers)
// into L2 right behind the TMA loads of boxes 0 and 1
__device__ __forceinline__ void attn_prefetch_l2(const MegaParams& p, int layer, int u, int pos) {
  const int rows = min((pos / 64 + 1) * 64, p.Lmax) - 128;
  if (rows <= 0) return;
  const size_t base = ((size_t)layer * p.B * p.NH + u) * p.Lmax * 64 + 128 * 64;
  const uint32_t bytes = (uint32_t)rows * 128u;
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.w.kcache + base), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.w.vcache + base), "r"(bytes) : "memory");
}

__device__ __forceinline__ void attn_warp(const MegaParams& p, int layer, int u, int pos,
                                          uint8_t* wbuf, uint64_t* bars, uint32_t& par, bool prefetched) {
  const int l = lane_id(), g4 = l & 3;
  const bool act = l < 4;  // lanes holding row 0 of the MMA fragments
  const int E = p.E, NH = p.NH;
  const int b = u / NH, h = u - b * NH;
  const int nch = pos / 64 + 1;
  if (!prefetched && l == 0) {
    fence_proxy_async_all();
    attn_issue(p, layer, u, 0, wbuf, bars);
    if (nch > 1) attn_issue(p, layer, u, 1, wbuf, bars);
  }
  // q / k / v of this position are FINAL bf16 values in qkv16 (phase A's epilogue applied the
  // folded LayerNorm, the bias and the 1/8 query scale): q as A-operand fragments, lane g4 holds
  // columns kk*16 + 2 g4 + {0,1} (a0) and +8 (a2)
  const bf16* qrow = p.w.qkv16 + (long long)b * 3 * E + h * 64;
  uint32_t qa[4][2];
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const uint32_t v = __ldcg(reinterpret_cast<const unsigned int*>(qrow + kk * 16 + hh * 8 + 2 * g4));
      qa[kk][hh] = act ? v : 0u;
    }
  }
  // this position's K (lanes 0-7) and V (lanes 8-15) rows, 8 dims per lane: to the cache and,
  // below, into the staged box
  const int ks = l >> 3, part = l & 7;
  uint4 newrow = make_uint4(0, 0, 0, 0);
  const size_t cbase = ((size_t)layer * p.B * NH + u) * p.Lmax * 64;
  if (ks < 2) {
    newrow = __ldcg(reinterpret_cast<const uint4*>(qrow + (ks + 1) * E + part * 8));
    bf16* dst = (ks == 0 ? p.w.kcache : p.w.vcache) + cbase + (size_t)pos * 64 + part * 8;
    *reinterpret_cast<uint4*>(dst) = newrow;
  }
  float m = -INFINITY, lsum = 0.f;
  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  // key-padding mask bits of keys 0..pos, one load batch up front: lane i keeps word i
  const int* km = p.w.keymask + (long long)b * p.Lmax;
  uint32_t mword = 0;
  {
    const int nw = (pos + 32) / 32;  // <= 32 words (Lmax <= 1024)
    int kv[8];
#pragma unroll 1
    for (int w0 = 0; w0 < nw; w0 += 8) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int key = (w0 + i) * 32 + l;
        kv[i] = (w0 + i < nw && key <= pos) ? km[key] : 0;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t bits = __ballot_sync(0xffffffffu, kv[i] != 0);
        if (l == w0 + i) mword = bits;
      }
    }
  }
#pragma unroll 1
  for (int c = 0; c < nch; ++c) {
    uint8_t* kb = wbuf + (c & 1) * 16384;
    uint8_t* vb = kb + 8192;
    const uint32_t mlo = __shfl_sync(0xffffffffu, mword, (2 * c) & 31);
    const uint32_t mhi = __shfl_sync(0xffffffffu, mword, (2 * c + 1) & 31);
    mbar_wait<20>(&bars[c & 1], (par >> (c & 1)) & 1u);
    par ^= 1u << (c & 1);
    if (c == nch - 1) {
      const int rl = pos - c * 64;
      if (ks < 2) *reinterpret_cast<uint4*>((ks == 0 ? kb : vb) + rl * 128 + ((part ^ (rl & 7)) << 4)) = newrow;
      __syncwarp();
    }
    // S = q K^T (row 0 of the 16 x 64 product)
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
    {
      const bf16* tile = reinterpret_cast<const bf16*>(kb);
#pragma unroll
      for (int nb = 0; nb < 8; nb += 2) {
        const int r = nb * 8 + (l & 7) + ((l >> 4) & 1) * 8;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4(tile_addr(tile, r, kk * 2 + ((l >> 3) & 1)), b0, b1, b2, b3);
          mma16816(s[nb], qa[kk][0], 0u, qa[kk][1], 0u, b0, b1);
          mma16816(s[nb + 1], qa[kk][0], 0u, qa[kk][1], 0u, b2, b3);
        }
      }
    }
    // mask + online softmax on lanes 0-3 (keys nb*8 + 2 g4 + {0,1}); other lanes hold zero rows
    float cmax = -INFINITY;
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int kl = nb * 8 + 2 * g4 + e;
        const bool ok = kl < 32 ? ((mlo >> kl) & 1u) : ((mhi >> (kl - 32)) & 1u);
        s[nb][e] = ok ? s[nb][e] : -INFINITY;
        cmax = fmaxf(cmax, s[nb][e]);
      }
    }
    cmax = fmaxf(cmax, __shfl_xor_sync(0xffffffffu, cmax, 1));
    cmax = fmaxf(cmax, __shfl_xor_sync(0xffffffffu, cmax, 2));
    const float m_new = __shfl_sync(0xffffffffu, fmaxf(m, cmax), 0);
    if (m_new != -INFINITY) {  // warp-uniform
      const float scale = (m == -INFINITY) ? 0.f : __expf(m - m_new);
      lsum *= scale;
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        o[nb][0] *= scale;
        o[nb][1] *= scale;
      }
      uint32_t pa[4][2];
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        const float p0 = (s[nb][0] == -INFINITY) ? 0.f : __expf(s[nb][0] - m_new);
        const float p1 = (s[nb][1] == -INFINITY) ? 0.f : __expf(s[nb][1] - m_new);
        lsum += p0 + p1;
        pa[nb >> 1][nb & 1] = act ? pack_bf16(p0, p1) : 0u;
      }
      const bf16* tile = reinterpret_cast<const bf16*>(vb);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int r = kk * 16 + (l & 7) + ((l >> 3) & 1) * 8;
#pragma unroll
        for (int nb = 0; nb < 8; nb += 2) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4_t(tile_addr(tile, r, nb + (l >> 4)), b0, b1, b2, b3);
          mma16816(o[nb], pa[kk][0], 0u, pa[kk][1], 0u, b0, b1);
          mma16816(o[nb + 1], pa[kk][0], 0u, pa[kk][1], 0u, b2, b3);
        }
      }
      m = m_new;
    }
    __syncwarp();
    if (c + 2 < nch && l == 0) {
      fence_proxy_async_all();
      attn_issue(p, layer, u, c + 2, wbuf, bars);
    }
  }
  lsum += __shfl_xor_sync(0xffffffffu, lsum, 1);
  lsum += __shfl_xor_sync(0xffffffffu, lsum, 2);
  const float inv = lsum > 0.f ? 1.f / lsum : 0.f;
  if (act) {
    bf16* dst = p.w.att16 + (long long)b * E + h * 64 + 2 * g4;
#pragma unroll
    for (int nb = 0; nb < 8; ++nb)
      *reinterpret_cast<uint32_t*>(dst + nb * 8) = pack_bf16(o[nb][0] * inv, o[nb][1] * inv);
  }
}


// ---- lm_head: 32-vocab-row units over the full K = 768 (wte' is [V, K]), direct stores ----
__device__ __forceinline__ void issue_w_head(const bf16* __restrict__ wte, int V, int E, int n0, uint8_t* buf, uint64_t* bar) {
  const int tid = threadIdx.x;
  const int kt = E / 64;
  for (int idx = tid; idx < kt * 256; idx += MG_THREADS) {
    const int t = idx >> 8, r = (idx >> 3) & 31, c = idx & 7;
    const bool ok = (n0 + r) < V;
    cp_async16(buf + t * 4096 + r * 128 + ((c ^ (r & 7)) << 4), wte + (long long)(ok ? n0 + r : 0) * E + t * 64 + c * 8,
               ok);
  }
  cp_async_arrive(bar);
}
// logits[0:B, n0:n0+32] = rstd (acts wte'^T - mean cs) + b'; acts: [64][a_stride(768)] bf16 in smem
__device__ __forceinline__ void head_unit(const uint8_t* acts, const uint8_t* buf, const float* stats,
                                          const float* __restrict__ cs, const float* __restrict__ bh, int B, int V,
                                          int E, int n0, float* __restrict__ logits) {
  constexpr int SA = a_stride(KS_L);
  const int warp = threadIdx.x >> 5, l = lane_id();
  const int m = warp & 3, nq = warp >> 2;
  float acc[2][4], csv[2][2], bhv[2][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll
  for (int nb = 0; nb < 2; ++nb)  // epilogue vectors: loads in flight during the product
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int col = n0 + nq * 16 + nb * 8 + (l & 3) * 2 + e;
      csv[nb][e] = col < V ? __ldg(cs + col) : 0.f;
      bhv[nb][e] = col < V ? __ldg(bh + col) : 0.f;
    }
  const int kt = E / 64;
  const uint32_t a_base = smem_u32(acts) + (uint32_t)((m * 16 + (l & 15)) * SA + (l >> 4) * 16);
#pragma unroll 4
  for (int t = 0; t < kt; ++t) {
    const bf16* tile = reinterpret_cast<const bf16*>(buf + t * 4096);
    const int r = nq * 16 + (l & 7) + ((l >> 4) & 1) * 8;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t a0, a1, a2, a3, b0, b1, b2, b3;
      ldsm_x4(a_base + (uint32_t)((t * 4 + kk) * 32), a0, a1, a2, a3);
      ldsm_x4(tile_addr(tile, r, kk * 2 + ((l >> 3) & 1)), b0, b1, b2, b3);
      mma16816(acc[0], a0, a1, a2, a3, b0, b1);
      mma16816(acc[1], a0, a1, a2, a3, b2, b3);
    }
  }
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    const int row = m * 16 + (l >> 2) + rr * 8;
    if (row >= B) continue;
    const float mean = stats[row * 2], rstd = stats[row * 2 + 1];
#pragma unroll
    for (int nb = 0; nb < 2; ++nb) {
      const int col = n0 + nq * 16 + nb * 8 + (l & 3) * 2;
      float* dst = logits + (long long)row * V + col;
      if (col < V) dst[0] = rstd * (acc[nb][2 * rr] - mean * csv[nb][0]) + bhv[nb][0];
      if (col + 1 < V) dst[1] = rstd * (acc[nb][2 * rr + 1] - mean * csv[nb][1]) + bhv[nb][1];
    }
  }
}

__global__ void __launch_bounds__(MG_THREADS, 1)
decode_mega_kernel(const __grid_constant__ MegaParams p) {
  extern __shared__ __align__(1024) uint8_t mg_smem[];
  float* stats = reinterpret_cast<float*>(mg_smem + OFF_STATS);  // [64][2]
  int* row_types = reinterpret_cast<int*>(mg_smem + OFF_STATS);  // [64] (projector phase only)
  uint64_t* bars = reinterpret_cast<uint64_t*>(mg_smem + OFF_BARS);
  const int G = gridDim.x, cta = blockIdx.x, tid = threadIdx.x, warp = tid >> 5;
  const int chunk = cta / CL, rank = (int)cluster_ctarank();
  const int B = p.B, E = p.E, NH = p.NH;
  const int j0 = *p.j_ptr;
  // the grid-barrier counter runs on from launch to launch: word 1 holds the value it had when
  // this launch started (written by the previous launch's CTA 0 after its last barrier)
  const unsigned int bar_base = __ldcg(p.w.barrier + 1);
  unsigned int phase = 0;
  int tr = 0;
  auto stamp = [&]() {
    if (p.trace && cta == 0 && tid == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      p.trace[tr++] = t;
    }
  };
  auto grid_wait = [&]() { barrier_wait(p.w.barrier, bar_base + (++phase) * (unsigned int)G); };
  stamp();
  if (tid == 0) {
    for (int i = 0; i < BAR_ATT; ++i)  // weight slots complete by byte count, cp.async ones by 256 thread arrivals
      mbar_init(&bars[i], (i == BAR_WAC || i == BAR_WD || i == BAR_WE || i == BAR_WP) ? 1 : MG_THREADS);
    for (int i = 0; i < 2 * ATT_WARPS; ++i) mbar_init(&bars[BAR_ATT + i], 1);
    fence_barrier_init();
    tma_prefetch_desc(&p.tm_k);
    tma_prefetch_desc(&p.tm_v);
  }
  __syncthreads();
  uint32_t par = 0;      // parity bits of the cp.async / bulk-copy barriers
  uint32_t att_par = 0;  // parity bits of this warp's two K/V box barriers
  uint8_t* wbuf = mg_smem + warp * 32768;
  uint64_t* wbars = bars + BAR_ATT + 2 * warp;
  uint8_t* slotAC = mg_smem + OFF_WAC;
  uint8_t* slotD = mg_smem + OFF_WD;
  uint8_t* slotE = mg_smem + OFF_WE;
  uint8_t* slotP = mg_smem + OFF_ACTS + OFF_WP;
  const bool inA = chunk < 3 * E / NC_A, inC = chunk < E / NC_C, inD = chunk < 4 * E / NC_D, inE = chunk < E / NC_E;
  const bool inP = p.full && chunk < E / NC_P;
  const size_t wA = (size_t)E * 3 * E, wD = (size_t)E * 4 * E;
  const int nsteps = p.full ? p.n_steps : 1;
  if (inA) issue_w<KS_S, NC_A>(p.w.f_attn, 3 * E, chunk, rank, slotAC, &bars[BAR_WAC]);
  if (inP) issue_w<KS_P, NC_P>(p.w.w2t, E, chunk, rank, slotP, &bars[BAR_WP]);

  for (int step = 0; step < nsteps; ++step) {
    const int j = j0 + step;
    const int pos = p.Pl + j;
    const bool last_step = step + 1 == nsteps;
    int nch1;  // chunks of the LN1 statistics partials the block-0 input comes with
    if (p.full) {
      // ---- P: h = tanh(T1[tok] + C1[pair] + b1) W2^T + b2 + wpe[pos] + wte[type]  (src/model.py:316-318,
      //      HF GPT2Model embeddings; decode_prep + the two projector GEMMs of the per-op step).
      //      T1 = table W1^T and C1 = ctx W1^T are precomputed per generation call (projector layer 1
      //      is linear in table[tok] + ctx). Inference-branch type id / key mask: src/model.py:300-312.
      const int sl = p.sent_len;
      if (tid < 64) {
        int ty = 0;
        if (tid < B) {
          const int tok = __ldcg(p.gen + (long long)tid * p.gen_ld + j);
          if ((unsigned)tok >= (unsigned)p.table_rows) {
            printf("mmtg: token id %d (row %d, step %d) is outside the token table [0, %d)\n", tok, tid, j, p.table_rows);
            __trap();
          }
          const int r = (j + 1) % sl;
          if (!(r == 0 || r == 1 || tok == 0)) {
            const int sidx = j / sl;
            ty = sidx < p.n_sent ? sidx + 1 : 1;
          }
          if (cta == 0) p.w.keymask[(long long)tid * p.Lmax + pos] = tok != 0 ? 1 : 0;
        }
        row_types[tid] = ty;
      }
      if (inP) {
        constexpr int SA = a_stride(KS_P);
        uint8_t* acts = mg_smem + OFF_ACTS;
        const int He = p.He, kctx = j / (2 * sl);
        const float* b1 = p.P + p.off.proj1_b;
        for (int idx = tid; idx < 64 * (KS_P / 2); idx += MG_THREADS) {
          const int r = idx / (KS_P / 2), c2 = idx - r * (KS_P / 2);
          const int k = rank * KS_P + 2 * c2;
          uint32_t packed = 0u;
          if (r < B) {
            const int tok = __ldcg(p.gen + (long long)r * p.gen_ld + j);
            float2 v = __ldg(reinterpret_cast<const float2*>(p.w.T1 + (long long)tok * He + k));
            if (kctx < p.S) {
              const float2 c = __ldg(reinterpret_cast<const float2*>(p.w.C1 + ((long long)kctx * B + r) * He + k));
              v.x += c.x;
              v.y += c.y;
            }
            const float2 bb = __ldg(reinterpret_cast<const float2*>(b1 + k));
            packed = pack_bf16(tanhf(v.x + bb.x), tanhf(v.y + bb.y));
          }
          *reinterpret_cast<uint32_t*>(acts + r * SA + c2 * 4) = packed;
        }
        EpiEmbed epi{p.P + p.off.proj2_b, p.P + p.off.wpe + (long long)pos * E, p.P + p.off.wte, row_types, p.w.h, p.w.h16,
                     p.w.stats1, chunk * NC_P, E, chunk, 0.f, 0.f};
        gemm_phase<KS_P, NC_P, false, EpiEmbed, true>(nullptr, 0, B, rank, mg_smem, slotP, bars, BAR_WP, par, epi);
      }
      nch1 = E / NC_P;
    } else {
      // ---- step-only mode: bf16 copy and LN1 row statistics of the block-0 input the caller wrote ----
      const int row = cta * 8 + warp;
      if (row < B) {
        const int l = lane_id();
        const float4* xr = reinterpret_cast<const float4*>(p.w.h + (long long)row * E);
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const float4 v = __ldcg(xr + l + i * 32);
          s1 += v.x + v.y + v.z + v.w;
          s2 += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
          *reinterpret_cast<uint2*>(p.w.h16 + (long long)row * E + (l + i * 32) * 4) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
        }
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
        if (l == 0) *reinterpret_cast<float2*>(p.w.stats1 + (long long)row * 32 * 2) = make_float2(s1, s2);
      }
      nch1 = 1;
    }
    barrier_arrive(p.w.barrier);
    grid_wait();
    stamp();

    for (int layer = 0; layer < p.NL; ++layer) {
      const mmtg_layer_offsets lo = p.off.layer[layer];
      const bool last = layer + 1 == p.NL;
      const float* fv = p.w.f_vec + (size_t)layer * 14 * E;  // cs_attn[3E] b_attn[3E] cs_fc[4E] b_fc[4E]
      auto tsub = [&](int o) -> unsigned long long* { return (p.trace && last && last_step && cta == 0) ? p.trace + o : nullptr; };
      // ---- A: qkv16 = bf16(rstd (bf16(h) W'_attn - mean cs) + b') ----
      if (inA) {
        rows_mean_rstd(p.w.stats1, nch1, rank * 16, 16, E, stats);
        EpiQKV epi{fv, fv + 3 * E, stats, p.w.qkv16, chunk * NC_A, E, rank * 16};
        gemm_phase<KS_S, NC_A, true>(p.w.h16, E, B, rank, mg_smem, slotAC, bars, BAR_WAC, par, epi, tsub(80));
      }
      barrier_arrive(p.w.barrier);
      substamp(tsub(80), 7);
      // window: W_proj of this block; cached K/V boxes of this CTA's first attention units
      if (inC) issue_w<KS_S, NC_C>(p.W + lo.proj_w, E, chunk, rank, slotAC, &bars[BAR_WAC]);
      if (warp < ATT_WARPS && lane_id() == 0) {
        const int u = cta + G * warp;
        if (u < B * NH) {
          fence_proxy_async_all();
          attn_issue(p, layer, u, 0, wbuf, wbars);
          if (pos >= 64) attn_issue(p, layer, u, 1, wbuf, wbars);
        }
      }
      substamp(tsub(80), 8);
      grid_wait();
      substamp(tsub(80), 9);
      stamp();
      // ---- B: cached attention, one warp per (row, head) ----
      substamp(tsub(120), 0);
      if (warp < ATT_WARPS) {
        bool first = true;
        for (int u = cta + G * warp; u < B * NH; u += G * ATT_WARPS) {
          attn_warp(p, layer, u, pos, wbuf, wbars, att_par, first);
          first = false;
          __syncwarp();
        }
      }
      substamp(tsub(120), 1);
      barrier_arrive(p.w.barrier);
      substamp(tsub(120), 2);
      // window: W'_fc and W_proj2 of this block (their slots were under the attention boxes)
      if (inD) issue_w<KS_S, NC_D>(p.w.f_fc + layer * wD, 4 * E, chunk, rank, slotD, &bars[BAR_WD]);
      if (inE) issue_w<KS_L, NC_E>(p.W + lo.proj2_w, E, chunk, rank, slotE, &bars[BAR_WE]);
      grid_wait();
      stamp();
      // ---- C: h2 = h + b_proj + att16 W_proj ----
      if (inC) {
        EpiResid epi{p.w.h, p.P + lo.proj_b, p.w.h2, p.w.h2_16, p.w.stats2, chunk * NC_C, E, chunk, 0.f, 0.f};
        gemm_phase<KS_S, NC_C, false>(p.w.att16, E, B, rank, mg_smem, slotAC, bars, BAR_WAC, par, epi, tsub(90));
      }
      barrier_arrive(p.w.barrier);
      // window: W'_attn of the next block (of block 0 of the NEXT position after the last block:
      // weights do not depend on the position); at the last block pull this CTA's lm_head rows into L2
      if (inA && (!last || !last_step))
        issue_w<KS_S, NC_A>(p.w.f_attn + (last ? 0 : layer + 1) * wA, 3 * E, chunk, rank, slotAC, &bars[BAR_WAC]);
      if (last && tid == 0) {
        for (int u = cta; u * 32 < p.V; u += G) {
          const uint32_t bytes = (uint32_t)min(32, p.V - u * 32) * (uint32_t)E * 2u;
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.w.f_wte + (size_t)u * 32 * E), "r"(bytes)
                       : "memory");
        }
      }
      grid_wait();
      stamp();
      // ---- D: u16 = bf16(gelu(rstd (bf16(h2) W'_fc - mean cs) + b')) ----
      if (inD) {
        rows_mean_rstd(p.w.stats2, E / NC_C, rank * 16, 16, E, stats);
        EpiGelu epi{fv + 6 * E, fv + 10 * E, stats, p.w.u16, chunk * NC_D, 4 * E, rank * 16};
        gemm_phase<KS_S, NC_D, false>(p.w.h2_16, E, B, rank, mg_smem, slotD, bars, BAR_WD, par, epi, tsub(100));
      }
      barrier_arrive(p.w.barrier);
      substamp(tsub(100), 7);
      grid_wait();
      substamp(tsub(100), 9);
      stamp();
      // ---- E: h = h2 + b_proj2 + u16 W_proj2 ----
      if (inE) {
        EpiResid epi{p.w.h2, p.P + lo.proj2_b, p.w.h, p.w.h16, p.w.stats1, chunk * NC_E, E, chunk, 0.f, 0.f};
        gemm_phase<KS_L, NC_E, false>(p.w.u16, 4 * E, B, rank, mg_smem, slotE, bars, BAR_WE, par, epi, tsub(110));
      }
      nch1 = E / NC_E;
      barrier_arrive(p.w.barrier);
      // window at the last block: the first two lm_head units
      if (last) {
        issue_w_head(p.w.f_wte, p.V, E, cta * 32, mg_smem + OFF_WD, &bars[BAR_H0]);
        issue_w_head(p.w.f_wte, p.V, E, (cta + G) * 32, mg_smem + OFF_WD + HEAD_BUF, &bars[BAR_H1]);
      }
      grid_wait();
      stamp();
    }
    // ---- F: logits = rstd (bf16(h) wte'^T - mean cs) + b' (tied lm_head with ln_f folded in) ----
    {
      const int nun = cdiv(p.V, 32);
      uint8_t* hb = mg_smem + OFF_WD;
      const float* hv = p.w.f_vec + (size_t)p.NL * 14 * E;  // cs_head[V] b_head[V]
      issue_acts<KS_L>(p.w.h16, E, 0, B, mg_smem + OFF_ACTS, &bars[BAR_ACTS]);
      rows_mean_rstd(p.w.stats1, nch1, 0, 64 < B ? 64 : B, E, stats);
      mbar_wait<23>(&bars[BAR_ACTS], (par >> BAR_ACTS) & 1u);
      par ^= 1u << BAR_ACTS;
      __syncthreads();  // stats
      int k = 0;
      for (int u = cta; u < nun; u += G, ++k) {
        const int hb_i = BAR_H0 + (k & 1);
        mbar_wait<24>(&bars[hb_i], (par >> hb_i) & 1u);
        par ^= 1u << hb_i;
        head_unit(mg_smem + OFF_ACTS, hb + (k & 1) * HEAD_BUF, stats, hv, hv + p.V, B, p.V, E, u * 32, p.logits);
        __syncthreads();
        if (u + 2 * G < nun) issue_w_head(p.w.f_wte, p.V, E, (u + 2 * G) * 32, hb + (k & 1) * HEAD_BUF, &bars[hb_i]);
      }
      // a CTA with fewer than two units still owns the completion of the prefetches it issued
      for (; k < 2; ++k) {
        const int hb_i = BAR_H0 + (k & 1);
        mbar_wait<25>(&bars[hb_i], (par >> hb_i) & 1u);
        par ^= 1u << hb_i;
      }
    }
    stamp();
    if (p.full) {
      // ---- S: sampler (src/generate.py:118-142), CTA b decides gen[b][j + 1] ----
      barrier_arrive(p.w.barrier);
      if (inP && !last_step) issue_w<KS_P, NC_P>(p.w.w2t, E, chunk, rank, slotP, &bars[BAR_WP]);
      grid_wait();
      if (cta < B) {
        sample_row(p.logits + (long long)cta * p.V, p.gen + (long long)cta * p.gen_ld, cta, j, 1, p.V, p.sent_len,
                   p.temperature, p.top_k, p.top_p, p.rep_penalty, p.seed_dev ? p.seed_dev[0] : 0ull, nullptr,
                   reinterpret_cast<float*>(mg_smem + OFF_ACTS), reinterpret_cast<SamplerScratch*>(mg_smem + OFF_ACTS + OFF_SAMP_SCR));
      }
      barrier_arrive(p.w.barrier);
      grid_wait();
      stamp();
    }
  }
  if (cta == 0 && tid == 0) {
    p.w.barrier[1] = bar_base + phase * (unsigned int)G;  // where the next launch starts counting
    if (p.full) *p.j_ptr = j0 + nsteps;
  }
}

unsigned long long* g_mega_trace = nullptr;

// ---- LayerNorm folding (once per set of weights) ----
// W'[k,n] = bf16(g[k] W[k,n]); cs[n] = sum_k W'[k,n]; bout[n] = bias[n] + sum_k beta[k] W[k,n]
__global__ void __launch_bounds__(256)
fold_kn_kernel(const float* __restrict__ Wsrc, const float* __restrict__ g, const float* __restrict__ beta,
               const float* __restrict__ bias, int K, int N, bf16* __restrict__ Wf, float* __restrict__ cs,
               float* __restrict__ bout) {
  // block = 32 columns x 8 row groups; row groups interleave over k, reduced through smem
  __shared__ float red[2][8][32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + tx;
  float s = 0.f, bb = 0.f;
  if (n < N) {
#pragma unroll 4
    for (int k = ty; k < K; k += 8) {
      const float w = Wsrc[(long long)k * N + n];
      const bf16 wf = __float2bfloat16(g[k] * w);
      Wf[(long long)k * N + n] = wf;
      s += __bfloat162float(wf);
      bb += beta[k] * w;
    }
  }
  red[0][ty][tx] = s;
  red[1][ty][tx] = bb;
  __syncthreads();
  if (ty == 0 && n < N) {
#pragma unroll
    for (int j = 1; j < 8; ++j) {
      s += red[0][j][tx];
      bb += red[1][j][tx];
    }
    cs[n] = s;
    bout[n] = bias[n] + bb;
  }
}
// tied lm_head: wte'[n,k] = bf16(g[k] wte[n,k]) (warp per vocabulary row)
__global__ void __launch_bounds__(256)
fold_head_kernel(const float* __restrict__ wte, const float* __restrict__ g, const float* __restrict__ beta, int V,
                 int E, bf16* __restrict__ Wf, float* __restrict__ cs, float* __restrict__ bout) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= V) return;
  const int l = lane_id();
  float s = 0.f, bb = 0.f;
  for (int k = l; k < E; k += 32) {
    const float w = wte[(long long)n * E + k];
    const bf16 wf = __float2bfloat16(g[k] * w);
    Wf[(long long)n * E + k] = wf;
    s += __bfloat162float(wf);
    bb += beta[k] * w;
  }
  s = warp_sum(s);
  bb = warp_sum(bb);
  if (l == 0) {
    cs[n] = s;
    bout[n] = bb;
  }
}

}  // namespace

int decode_fold_weights(const mmtg_model* m, const MegaBufs& w, cudaStream_t st) {
  const mmtg_dims& d = m->dims;
  const int E = d.E;
  const float* P = m->params;
  for (int l = 0; l < d.NL; ++l) {
    const mmtg_layer_offsets& lo = m->off.layer[l];
    float* fv = w.f_vec + (size_t)l * 14 * E;
    fold_kn_kernel<<<cdiv(3 * E, 32), 256, 0, st>>>(P + lo.attn_w, P + lo.ln1_w, P + lo.ln1_b, P + lo.attn_b, E, 3 * E,
                                                     w.f_attn + (size_t)l * E * 3 * E, fv, fv + 3 * E);
    MMTG_LAUNCH_OK();
    fold_kn_kernel<<<cdiv(4 * E, 32), 256, 0, st>>>(P + lo.fc_w, P + lo.ln2_w, P + lo.ln2_b, P + lo.fc_b, E, 4 * E,
                                                     w.f_fc + (size_t)l * E * 4 * E, fv + 6 * E, fv + 10 * E);
    MMTG_LAUNCH_OK();
  }
  float* hv = w.f_vec + (size_t)d.NL * 14 * E;
  fold_head_kernel<<<cdiv(d.V, 8), 256, 0, st>>>(P + m->off.wte, P + m->off.lnf_w, P + m->off.lnf_b, d.V, E, w.f_wte,
                                                 hv, hv + d.V);
  MMTG_LAUNCH_OK();
  count_launch(2 * d.NL + 1);
  return 0;
}

// Launch: all blocks + lm_head of one decode position. bufs.h holds the block-0 input (projector
// output + wpe + wte[type]); bufs.barrier must be zero at launch (decode_prep resets it).
int decode_mega_launch(const mmtg_model* m, int Lmax, const MegaBufs& bufs, int* j_ptr, float* logits,
                       const MegaStepArgs* full, cudaStream_t st) {
  const mmtg_dims& d = m->dims;
  MMTG_CHECK_ARG(d.B <= 64 && d.E == 768 && d.NH * 64 == d.E && Lmax <= 1024,
                 "decode megakernel: unsupported shape (B=%d E=%d Lmax=%d)", d.B, d.E, Lmax);
  MMTG_PER_DEVICE_FLAG(attr_set);
  static int grid_by_dev[64] = {0};
  int& grid = grid_by_dev[current_device_index()];
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeCooperative;
  attr[1].val.cooperative = 1;
  cfg.blockDim = dim3(MG_THREADS);
  cfg.dynamicSmemBytes = MG_SMEM;
  cfg.stream = st;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  if (!attr_set) {
    MMTG_CUDA_OK(cudaFuncSetAttribute(decode_mega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MG_SMEM));
    // the grid barrier needs every CTA co-resident: as many clusters of 4 as the device can hold
    // at once with one CTA per SM (33 on a B200: cluster size 4 strands 16 of the 148 SMs)
    cfg.gridDim = dim3(num_sms() / CL * CL);
    int ncl = 0;
    MMTG_CUDA_OK(cudaOccupancyMaxActiveClusters(&ncl, decode_mega_kernel, &cfg));
    MMTG_CHECK_ARG(ncl >= 32, "decode megakernel: only %d clusters of %d CTAs fit at once (32 needed)", ncl, CL);
    grid = (ncl < num_sms() / CL ? ncl : num_sms() / CL) * CL;
    attr_set = true;
  }
  cfg.gridDim = dim3(grid);
  MegaParams p;
  p.P = m->params;
  p.W = (const bf16*)m->params_bf16;
  p.off = m->off;
  const uint64_t rows = (uint64_t)d.NL * d.B * d.NH * Lmax;
  MMTG_TRY(make_tmap_bf16_2d(&p.tm_k, bufs.kcache, 64, rows, 64, 64, 64));
  MMTG_TRY(make_tmap_bf16_2d(&p.tm_v, bufs.vcache, 64, rows, 64, 64, 64));
  p.B = d.B; p.E = d.E; p.NH = d.NH; p.NL = d.NL; p.V = d.V; p.Pl = d.P; p.Lmax = Lmax;
  p.w = bufs;
  p.j_ptr = j_ptr;
  p.logits = logits;
  p.trace = g_mega_trace;
  p.full = full ? 1 : 0;
  p.n_steps = full ? full->n_steps : 1;
  p.gen = full ? full->gen : nullptr;
  p.gen_ld = full ? full->gen_ld : 0;
  p.sent_len = full ? full->sent_len : 1;
  p.n_sent = full ? full->n_sent : 0;
  p.S = d.S; p.He = d.He; p.table_rows = m->table_rows < d.V ? m->table_rows : d.V;
  p.temperature = full ? full->temperature : 1.f;
  p.top_k = full ? full->top_k : 1;
  p.top_p = full ? full->top_p : 0.f;
  p.rep_penalty = full ? full->rep_penalty : 1.f;
  p.seed_dev = full ? full->seed_dev : nullptr;
  if (full) {
    MMTG_CHECK_ARG(full->n_steps >= 1 && full->gen && full->temperature > 0.f && full->top_k >= 0 && full->top_k <= MAX_SURV &&
                       d.He == 4 * KS_P && d.V * 4 <= OFF_SAMP_SCR && d.V > 102,
                   "decode megakernel (full-step mode): unsupported arguments");
  }
  MMTG_CUDA_OK(cudaLaunchKernelEx(&cfg, decode_mega_kernel, p));
  count_launch();
  return 0;
}

}  // namespace mmtg

// Debug: CTA 0 writes a globaltimer stamp (ns) at kernel start, after the prologue barrier, after
// every phase barrier (5 per block) and at its end into `dev_buf` (>= 80 entries); NULL disables.
extern "C" int mmtg_decode_set_trace(uint64_t* dev_buf) {
  mmtg::g_mega_trace = (unsigned long long*)dev_buf;
  return 0;
}
