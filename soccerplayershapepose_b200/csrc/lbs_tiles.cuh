// Pieces shared by the skinning kernels (lbs.cu) and the fused blend + skinning forward (fused_fwd.cu): item
// shapes, plan record access, the transposing flush of a staging tile into (B, V, 3) row segments and the
// packed-pair slot cache.
#pragma once

#include "skin_common.cuh"

namespace b200smpl {

template <int HV>
struct ItemShape {
  static constexpr int ROWS = HV * 3;             // blend rows (floats per body) of one item
  static constexpr int HROW = ROWS + 4;           // staging row pitch: 16-byte aligned, HROW / 4 odd -> a quarter-warp
  static constexpr int TILE_WORDS = 32 * HROW;    // of 128-bit row accesses (lane = row) covers all banks
  static constexpr int NCH4 = ROWS / 4;           // float4 chunks of v_posed per body
  static constexpr int NCH8 = ROWS / 8;           // 8-row chunks of dvp per body
  static constexpr int PAIRS = ROWS / 2;          // 8-byte pieces per body row
  static constexpr int VP_WORDS = NCH4 * 128;     // dense v_posed block [NCH4][32] float4
  static constexpr int PLAN_WORDS = (HV / 8) * 40;  // plan records of one item
  static constexpr uint32_t TX_BYTES = (VP_WORDS + PLAN_WORDS) * 4;
  static_assert((HROW / 4) % 2 == 1 && ROWS % 8 == 0 && HV % 8 == 0, "item shape");
};

// plan words of vertices [4u, 4u+4) of the item inside the stash (records of 8 vertices: 8 float4 + 8 words)
__device__ __forceinline__ const float4* plan_wts(const uint32_t* stash, int u) {
  return reinterpret_cast<const float4*>(stash + (u >> 1) * 40 + (u & 1) * 16);
}
__device__ __forceinline__ const uint32_t* plan_meta(const uint32_t* stash, int u) {
  return stash + (u >> 1) * 40 + 32 + (u & 1) * 4;
}

// ---- rows of the staging tile <-> row segments of a (B, V, 3) tensor, in 8-byte pieces.  The 32 * PAIRS pieces
// are taken 32 per warp instruction: piece = k * 32 + lane -> row = piece / PAIRS, column pair = piece % PAIRS.
// Since 96 = RPP * PAIRS the (row offset, column) of a lane repeats every 3 instructions, RPP rows further down,
// so a lane keeps 3 global pointers and adds a constant stride.
template <int HV>
__device__ __forceinline__ void tile_to_global_full(const float* tile, float* dst0, size_t row_stride, int lane) {
  using SH = ItemShape<HV>;
  constexpr int RPP = 96 / SH::PAIRS, J = 32 / RPP;
  float* gp[3];
  const float* sp[3];
#pragma unroll
  for (int kk = 0; kk < 3; ++kk) {
    const int piece = kk * 32 + lane, rr = piece / SH::PAIRS, c = (piece - rr * SH::PAIRS) * 2;
    gp[kk] = dst0 + (size_t)rr * row_stride + c;
    sp[kk] = tile + rr * SH::HROW + c;
  }
  const size_t step = (size_t)RPP * row_stride;
#pragma unroll
  for (int j = 0; j < J; ++j)
#pragma unroll
    for (int kk = 0; kk < 3; ++kk) {
      st_stream2(gp[kk], *reinterpret_cast<const float2*>(sp[kk] + j * RPP * SH::HROW));
      gp[kk] += step;
    }
}
template <int HV>
__device__ __forceinline__ void tile_to_global(const float* tile, float* dst0, size_t row_stride, int nrows,
                                               int ncols, int lane) {
  using SH = ItemShape<HV>;
  if (nrows == 32 && ncols == SH::ROWS) {
    tile_to_global_full<HV>(tile, dst0, row_stride, lane);
    return;
  }
#pragma unroll 1
  for (int k = 0; k < SH::PAIRS; ++k) {
    const int piece = k * 32 + lane;
    const int r = piece / SH::PAIRS;
    const int c = (piece - r * SH::PAIRS) * 2;
    const float2 v = *reinterpret_cast<const float2*>(tile + r * SH::HROW + c);
    float* dst = dst0 + (size_t)r * row_stride + c;
    if (r < nrows) {
      if (c + 1 < ncols) st_stream2(dst, v);
      else if (c < ncols) st_stream(dst, v.x);
    }
  }
}

// element-wise fallbacks (odd V or a base pointer that is not 8-byte aligned)
template <int HV>
__device__ __forceinline__ void tile_to_global_scalar(const float* tile, float* dst0, size_t row_stride, int nrows,
                                                      int ncols, int lane) {
  for (int r = 0; r < nrows; ++r)
    for (int c = lane; c < ncols; c += 32) dst0[(size_t)r * row_stride + c] = tile[r * ItemShape<HV>::HROW + c];
}

// The four cached transforms ("slots") as packed pairs: x/y rows of a slot are (r0c, r1c) pairs, the z rows of
// two slots share pairs (lo = slots 0 / 2, hi = slots 1 / 3), so a vertex costs 25 FFMA2-class instructions
// instead of 48 scalar FFMA.
struct SlotXY {
  f2 c0, c1, c2, t;          // (r00 r10) (r01 r11) (r02 r12) (t0 t1)
};
struct SlotZ2 {
  f2 r20, r21, r22, t2;      // z row of two slots
};
struct Slots {
  SlotXY s0, s1, s2, s3;
  SlotZ2 zA, zB;
};

// gradient rows: global -> staging tile with 8-byte cp.async (fully asynchronous, no registers held)
__device__ __forceinline__ void cp_async8_zfill(void* dst_smem, const void* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global.L2::256B [%0], [%1], 8, %2;" ::"r"(smem_addr(dst_smem)), "l"(src), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
template <int HV>
__device__ __forceinline__ void issue_rows(float* tile, const float* __restrict__ src0, size_t row_stride, int nrows,
                                           int ncols, int lane) {
  using SH = ItemShape<HV>;
  constexpr int RPP = 96 / SH::PAIRS, J = 32 / RPP;
  if (nrows == 32 && ncols == SH::ROWS) {
    const float* gp[3];
    float* sp[3];
#pragma unroll
    for (int kk = 0; kk < 3; ++kk) {
      const int piece = kk * 32 + lane, rr = piece / SH::PAIRS, c = (piece - rr * SH::PAIRS) * 2;
      gp[kk] = src0 + (size_t)rr * row_stride + c;
      sp[kk] = tile + rr * SH::HROW + c;
    }
    const size_t step = (size_t)RPP * row_stride;
#pragma unroll
    for (int j = 0; j < J; ++j)
#pragma unroll
      for (int kk = 0; kk < 3; ++kk) {
        cp_async8_zfill(sp[kk] + j * RPP * SH::HROW, gp[kk], 8);
        gp[kk] += step;
      }
    return;
  }
#pragma unroll 1
  for (int k = 0; k < SH::PAIRS; ++k) {
    const int piece = k * 32 + lane;
    const int r = piece / SH::PAIRS;
    const int c = (piece - r * SH::PAIRS) * 2;
    const bool ok = r < nrows && c < ncols;
    const int bytes = ok ? (c + 1 < ncols ? 8 : 4) : 0;
    cp_async8_zfill(tile + r * SH::HROW + c, ok ? src0 + (size_t)r * row_stride + c : src0, bytes);
  }
}

// Slot pairs (lo = slots 0 / 2, hi = slots 1 / 3): rotation entries and gradient accumulators as packed pairs, so
// the per-vertex arithmetic is 54 FFMA2-class instructions instead of 111 scalar ones.
struct SlotPair {
  f2 R[9];                   // R[r * 3 + c]
  f2 D[AELEMS];              // dL/dA accumulators, D[r * 4 + c]
};
struct BwdState {
  SlotPair A, B;
  uint32_t prev;             // plan word of the previous vertex (joint ids of the slots)
  float sx, sy, sz;
};
// close one slot: its accumulator halves -> fp32 RED into the group's dA rows [joint * 12 + e][32]; clear them
template <bool HI>
__device__ __forceinline__ void flush_half(SlotPair& P, float* dA_g, int joint, int lane) {
  float* p = dA_g + (size_t)joint * AELEMS * 32 + lane;
#pragma unroll
  for (int e = 0; e < AELEMS; ++e) {
    red_add(p + e * 32, get_half<HI>(P.D[e]));
    P.D[e] = set_half<HI>(P.D[e], 0.f);
  }
}

// close the accumulators of the current group: 4 slots -> dA, translation sums -> dtransl
__device__ __forceinline__ void bwd_close_group(BwdState& s, float* dA_g, float* dtr_g, int lane) {
  flush_half<false>(s.A, dA_g, s.prev & 31, lane);
  flush_half<true>(s.A, dA_g, (s.prev >> 5) & 31, lane);
  flush_half<false>(s.B, dA_g, (s.prev >> 10) & 31, lane);
  flush_half<true>(s.B, dA_g, (s.prev >> 15) & 31, lane);
  red_add(dtr_g + lane, s.sx);
  red_add(dtr_g + 32 + lane, s.sy);
  red_add(dtr_g + 64 + lane, s.sz);
  s.sx = s.sy = s.sz = 0.f;
}


}  // namespace b200smpl
