// mma.sync (HMMA m16n8k16 bf16) tile helpers on 64x64 bf16 shared-memory tiles with XOR-swizzled
// 16-byte chunks: cp.async tile loads, ldmatrix fragment loads, NT / NN tile products.
// Shared by the attention backward kernels and the skinny (M <= 64) decode GEMM.
#pragma once
#include "common.cuh"

namespace mmtg {
namespace {

constexpr int ATT_THREADS = 128;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool pred) {
  const uint32_t d = smem_u32(smem_dst);
  const int sz = pred ? 16 : 0;  // zero-fill when out of range
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N));
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                        uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                          uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2,
                                         uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
      "{%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

// A [64 rows][64 cols] bf16 tile in smem, 128 B per row, 16-B chunk c of row r stored at c^(r&7).
__device__ __forceinline__ uint32_t tile_addr(const bf16* tile, int row, int chunk) {
  return smem_u32(tile) + (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4));
}
// Load rows [row0, row0+64) x 64 columns (col0..col0+63) of a [*, ld] bf16 matrix; rows >= nrows
// are zero-filled. 128 threads, 4 x 16-B chunks each.
__device__ __forceinline__ void load_tile_async(bf16* tile, const bf16* g, long long ld, int row0,
                                                int nrows_valid, int col0) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = threadIdx.x + i * ATT_THREADS;  // 0..511
    const int r = idx >> 3, c = idx & 7;
    const bool ok = (row0 + r) < nrows_valid;
    const bf16* src = g + (long long)(ok ? (row0 + r) : 0) * ld + col0 + c * 8;
    cp_async16(reinterpret_cast<uint8_t*>(tile) + r * 128 + ((c ^ (r & 7)) << 4), src, ok);
  }
}

// A-operand fragments (16 rows x 64 cols = 4 k-steps) of rows [row0, row0+16) of a tile.
__device__ __forceinline__ void load_a_frags(const bf16* tile, int row0, uint32_t (&a)[4][4]) {
  const int l = lane_id();
  const int r = row0 + (l & 15), cg = l >> 4;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk)
    ldsm_x4(tile_addr(tile, r, kk * 2 + cg), a[kk][0], a[kk][1], a[kk][2], a[kk][3]);
}

// C[16 x 64] += A_frags(16 x 64) * Tile^T where Tile is [64 n][64 k] (k contiguous): "NT".
__device__ __forceinline__ void mma_nt(float (&c)[8][4], const uint32_t (&a)[4][4],
                                       const bf16* tile) {
  const int l = lane_id();
#pragma unroll
  for (int nb = 0; nb < 8; nb += 2) {
    const int r = nb * 8 + (l & 7) + ((l >> 4) & 1) * 8;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4(tile_addr(tile, r, kk * 2 + ((l >> 3) & 1)), b0, b1, b2, b3);
      mma16816(c[nb], a[kk][0], a[kk][1], a[kk][2], a[kk][3], b0, b1);
      mma16816(c[nb + 1], a[kk][0], a[kk][1], a[kk][2], a[kk][3], b2, b3);
    }
  }
}
// C[16 x 64] += P(16 x 64, fp32 C-fragments converted to bf16) * Tile where Tile is [64 k][64 n]
// (n contiguous): "NN" via ldmatrix.trans.
__device__ __forceinline__ void mma_nn(float (&c)[8][4], const float (&p)[8][4], const bf16* tile) {
  const int l = lane_id();
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    const uint32_t a0 = pack_bf16(p[2 * kk][0], p[2 * kk][1]);
    const uint32_t a1 = pack_bf16(p[2 * kk][2], p[2 * kk][3]);
    const uint32_t a2 = pack_bf16(p[2 * kk + 1][0], p[2 * kk + 1][1]);
    const uint32_t a3 = pack_bf16(p[2 * kk + 1][2], p[2 * kk + 1][3]);
    const int r = kk * 16 + (l & 7) + ((l >> 3) & 1) * 8;
#pragma unroll
    for (int nb = 0; nb < 8; nb += 2) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(tile_addr(tile, r, nb + (l >> 4)), b0, b1, b2, b3);
      mma16816(c[nb], a0, a1, a2, a3, b0, b1);
      mma16816(c[nb + 1], a0, a1, a2, a3, b2, b3);
    }
  }
}


// C[16 x 64] += A_frags(16 x 64, from smem) * Tile where Tile is [64 k][64 n] (n contiguous).
__device__ __forceinline__ void mma_nn_a(float (&c)[8][4], const uint32_t (&a)[4][4], const bf16* tile) {
  const int l = lane_id();
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    const int r = kk * 16 + (l & 7) + ((l >> 3) & 1) * 8;
#pragma unroll
    for (int nb = 0; nb < 8; nb += 2) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(tile_addr(tile, r, nb + (l >> 4)), b0, b1, b2, b3);
      mma16816(c[nb], a[kk][0], a[kk][1], a[kk][2], a[kk][3], b0, b1);
      mma16816(c[nb + 1], a[kk][0], a[kk][1], a[kk][2], a[kk][3], b2, b3);
    }
  }
}

}  // namespace
}  // namespace mmtg
