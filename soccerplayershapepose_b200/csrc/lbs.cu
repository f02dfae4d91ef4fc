// Linear blend skinning of the 6890 vertices, forward and backward (SURVEY.md section 8 row a9;
// smplx.lbs.lbs tail: T = W.A ; v = T.[v_posed;1]).  HBM-roofline kernels.
//
// Mapping: lane = body.  The work is the flat list of (body group, 32-vertex tile) items; every WARP
// owns one contiguous run of that list (no CTA-level coupling, no barriers), so the load is balanced to
// one tile per warp whatever the batch is.  Everything per-vertex (4 joint ids, 4 weights) is
// warp-uniform; everything per-body lives in registers: the 4 cached joint transforms ("slots") are
// (re)loaded -- three coalesced 512-byte float4 loads from A_blk, L1/L2 resident -- only where the packed
// plan says a slot's joint changes (8 times per 32 vertices on the SMPL mesh order).
//
// Data movement: v_posed comes from the blend GEMM as [n/4][S][4] (a warp reads 512 contiguous bytes per
// float4) straight into registers, 8 vertices ahead of the arithmetic.  The tensors whose layout the
// caller fixes -- (B, 6890, 3) vertices / vertex gradients -- are transposed through a per-warp
// shared-memory tile [32 bodies][100]: the lane = body side uses 128-bit accesses on its own row, the
// global side moves each 384-byte row segment as 8-byte accesses (two rows per three warp
// instructions), every byte read or written exactly once with streaming accesses.  The backward writes
// dv_posed as bf16 hi/lo 16-byte chunks [n/8][S][8] (512 contiguous bytes per warp store, the operand
// layout of the gradient GEMM) and adds dA / dtransl into the slab accumulators with fp32 REDs.
#include "skin_common.cuh"

namespace b200smpl {

struct WarpRun {
  int item, end, g, t;
};
// contiguous run of (group, tile) items of warp `gw` out of `nw`
__device__ __forceinline__ WarpRun make_run(int ngroups, int ntiles, int gw, int nw) {
  const long long total = (long long)ngroups * ntiles;
  WarpRun r;
  r.item = (int)(total * gw / nw);
  r.end = (int)(total * (gw + 1) / nw);
  r.g = r.item / ntiles;
  r.t = r.item - r.g * ntiles;
  return r;
}

// the 24 v_posed rows (8 vertices) of one body: 6 float4, chunk stride S
__device__ __forceinline__ void fetch_vp8(float (&P)[24], const float4* __restrict__ p, size_t S) {
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const float4 v = ld_stream4(p + i * S);
    P[i * 4 + 0] = v.x; P[i * 4 + 1] = v.y; P[i * 4 + 2] = v.z; P[i * 4 + 3] = v.w;
  }
}
__device__ __forceinline__ void fetch_vp4(float (&P)[12], const float4* __restrict__ p, size_t S) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float4 v = ld_stream4(p + i * S);
    P[i * 4 + 0] = v.x; P[i * 4 + 1] = v.y; P[i * 4 + 2] = v.z; P[i * 4 + 3] = v.w;
  }
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
constexpr int FWD_WARPS = 12;
constexpr int FWD_THREADS = FWD_WARPS * 32;
constexpr size_t FWD_SMEM = (size_t)FWD_WARPS * TTILE_WORDS * 4;

struct Slots {
  float a0[AELEMS], a1[AELEMS], a2[AELEMS], a3[AELEMS];
};

// 8 vertices: P -> skinned coordinates -> 6 float4 stores into the lane's own row of the staging tile
__device__ __forceinline__ void skin_fwd8(const float (&P)[24], Slots& s, const float4* __restrict__ A_g, int lane,
                                          const uint32_t* __restrict__ meta, const float4* __restrict__ wts,
                                          uint32_t force, float tx, float ty, float tz, float* row_out) {
  uint32_t mts[8];
  {
    const uint4 m0 = __ldg(reinterpret_cast<const uint4*>(meta));
    const uint4 m1 = __ldg(reinterpret_cast<const uint4*>(meta) + 1);
    mts[0] = m0.x | force; mts[1] = m0.y; mts[2] = m0.z; mts[3] = m0.w;
    mts[4] = m1.x; mts[5] = m1.y; mts[6] = m1.z; mts[7] = m1.w;
  }
  float o[24];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint32_t mt = mts[i];
    const float4 w = __ldg(wts + i);
    if (mt & (0xFu << 20)) {
      if (mt & (1u << 20)) load_slot_g(s.a0, A_g, mt & 31, lane);
      if (mt & (1u << 21)) load_slot_g(s.a1, A_g, (mt >> 5) & 31, lane);
      if (mt & (1u << 22)) load_slot_g(s.a2, A_g, (mt >> 10) & 31, lane);
      if (mt & (1u << 23)) load_slot_g(s.a3, A_g, (mt >> 15) & 31, lane);
    }
    const float px = P[i * 3], py = P[i * 3 + 1], pz = P[i * 3 + 2];
    float ox = tx, oy = ty, oz = tz;
#define B200_SKIN(a, wk)                                                        \
  ox = fmaf(wk, fmaf(a[0], px, fmaf(a[1], py, fmaf(a[2], pz, a[3]))), ox);      \
  oy = fmaf(wk, fmaf(a[4], px, fmaf(a[5], py, fmaf(a[6], pz, a[7]))), oy);      \
  oz = fmaf(wk, fmaf(a[8], px, fmaf(a[9], py, fmaf(a[10], pz, a[11]))), oz);
    B200_SKIN(s.a0, w.x)
    B200_SKIN(s.a1, w.y)
    B200_SKIN(s.a2, w.z)
    B200_SKIN(s.a3, w.w)
#undef B200_SKIN
    o[i * 3] = ox; o[i * 3 + 1] = oy; o[i * 3 + 2] = oz;
  }
#pragma unroll
  for (int i = 0; i < 6; ++i)
    *reinterpret_cast<float4*>(row_out + i * 4) = make_float4(o[i * 4], o[i * 4 + 1], o[i * 4 + 2], o[i * 4 + 3]);
}

// rows of the staging tile <-> 384-byte row segments of a (B, V, 3) tensor.  Three warp instructions move
// two rows: slot = k * 32 + lane in [0, 96) -> row = slot / 48, column pair = slot % 48.
template <bool FULL>
__device__ __forceinline__ void tile_to_global(const float* tile, float* dst0, size_t row_stride, int nrows,
                                               int ncols, int lane) {
#pragma unroll 4
  for (int r2 = 0; r2 < 32; r2 += 2) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int slot = k * 32 + lane;
      const int rr = slot >= 48 ? 1 : 0;
      const int c = (slot - 48 * rr) * 2;
      const int r = r2 + rr;
      const float2 v = *reinterpret_cast<const float2*>(tile + r * TROW + c);
      float* dst = dst0 + (size_t)r * row_stride + c;
      if (FULL) {
        st_stream2(dst, v);
      } else if (r < nrows) {
        if (c + 1 < ncols) st_stream2(dst, v);
        else if (c < ncols) st_stream(dst, v.x);
      }
    }
  }
}
template <bool FULL>
__device__ __forceinline__ void global_to_tile(float* tile, const float* src0, size_t row_stride, int nrows, int ncols,
                                               int lane) {
#pragma unroll 1
  for (int r8 = 0; r8 < 32; r8 += 8) {   // 12 eight-byte loads in flight per lane, then 12 shared stores
    float2 v[12];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int slot = k * 32 + lane;
        const int rr = slot >= 48 ? 1 : 0;
        const int c = (slot - 48 * rr) * 2;
        const int r = r8 + q * 2 + rr;
        const float* src = src0 + (size_t)r * row_stride + c;
        if (FULL) {
          v[q * 3 + k] = ld_stream2(src);
        } else {
          float2 x = make_float2(0.f, 0.f);
          if (r < nrows) {
            if (c + 1 < ncols) x = ld_stream2(src);
            else if (c < ncols) x.x = ld_stream(src);
          }
          v[q * 3 + k] = x;
        }
      }
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int slot = k * 32 + lane;
        const int rr = slot >= 48 ? 1 : 0;
        const int c = (slot - 48 * rr) * 2;
        *reinterpret_cast<float2*>(tile + (r8 + q * 2 + rr) * TROW + c) = v[q * 3 + k];
      }
  }
}
// element-wise fallbacks (odd V or a base pointer that is not 8-byte aligned)
__device__ __forceinline__ void tile_to_global_scalar(const float* tile, float* dst0, size_t row_stride, int nrows,
                                                      int ncols, int lane) {
  for (int r = 0; r < nrows; ++r)
    for (int c = lane; c < ncols; c += 32) dst0[(size_t)r * row_stride + c] = tile[r * TROW + c];
}
__device__ __forceinline__ void global_to_tile_scalar(float* tile, const float* src0, size_t row_stride, int nrows,
                                                      int ncols, int lane) {
  for (int r = 0; r < 32; ++r)
    for (int c = lane; c < 96; c += 32)
      tile[r * TROW + c] = (r < nrows && c < ncols) ? src0[(size_t)r * row_stride + c] : 0.f;
}

__global__ void __launch_bounds__(FWD_THREADS, 1)
lbs_fwd_kernel(const float4* __restrict__ vpB, int S, const float4* __restrict__ A_blk, int b0, int nb, int ngroups,
               const float* __restrict__ transl, float* __restrict__ verts, int V, int ntiles, int vec_ok,
               const uint32_t* __restrict__ vmeta, const float4* __restrict__ vwts) {
  extern __shared__ __align__(128) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* tile = smem + warp * TTILE_WORDS;
  float* my_row = tile + lane * TROW;
  WarpRun run = make_run(ngroups, ntiles, blockIdx.x * FWD_WARPS + warp, gridDim.x * FWD_WARPS);
  if (run.item >= run.end) return;

  float P[24], Q[24];
  int g = run.g, t = run.t;
  const float4* vp_lane = vpB + (size_t)g * 32 + lane;            // + chunk * S
  fetch_vp8(P, vp_lane + (size_t)(t * 24) * S, S);
  Slots sl;
  int cur_g = -1;
  const float4* A_g = nullptr;
  float tx = 0.f, ty = 0.f, tz = 0.f;
  for (int item = run.item; item < run.end; ++item) {
    uint32_t force = 0u;
    if (g != cur_g) {                                             // new body group: translation + all four slots
      cur_g = g;
      A_g = A_blk + (size_t)g * (AG_WORDS / 4);
      tx = ty = tz = 0.f;
      if (transl != nullptr && g * 32 + lane < nb) {
        const float* tp = transl + (size_t)(b0 + g * 32 + lane) * 3;
        tx = tp[0]; ty = tp[1]; tz = tp[2];
      }
      force = 0xFu << 20;
    } else if (item == run.item) {
      force = 0xFu << 20;
    }
    const int vbase = t * TILE_V;
    const uint32_t* meta = vmeta + vbase;
    const float4* wts = vwts + vbase;
    const float4* vp_t = vp_lane + (size_t)(t * 24) * S;
    // next item (for the prefetch that crosses the tile boundary)
    int ng = g, nt = t + 1;
    if (nt == ntiles) { nt = 0; ++ng; }
    fetch_vp8(Q, vp_t + 6 * S, S);
    skin_fwd8(P, sl, A_g, lane, meta, wts, force, tx, ty, tz, my_row);
    fetch_vp8(P, vp_t + 12 * S, S);
    skin_fwd8(Q, sl, A_g, lane, meta + 8, wts + 8, 0u, tx, ty, tz, my_row + 24);
    fetch_vp8(Q, vp_t + 18 * S, S);
    skin_fwd8(P, sl, A_g, lane, meta + 16, wts + 16, 0u, tx, ty, tz, my_row + 48);
    const float4* vp_next = vpB + (size_t)ng * 32 + lane;
    if (item + 1 < run.end) fetch_vp8(P, vp_next + (size_t)(nt * 24) * S, S);
    skin_fwd8(Q, sl, A_g, lane, meta + 24, wts + 24, 0u, tx, ty, tz, my_row + 72);
    __syncwarp();
    // flush: each body row of the tile is 96 contiguous floats of the (B, V, 3) output
    {
      const int nrows = min(32, nb - g * 32);
      const int ncols = min(TILE_V, V - vbase) * 3;
      float* dst0 = verts + ((size_t)(b0 + g * 32) * V + vbase) * 3;
      if (!vec_ok) tile_to_global_scalar(tile, dst0, (size_t)V * 3, nrows, ncols, lane);
      else if (nrows == 32 && ncols == 96) tile_to_global<true>(tile, dst0, (size_t)V * 3, 32, 96, lane);
      else tile_to_global<false>(tile, dst0, (size_t)V * 3, nrows, ncols, lane);
    }
    __syncwarp();
    if (ng != g) vp_lane = vp_next;
    g = ng; t = nt;
  }
}

// ---------------------------------------------------------------------------------------------
// backward:  given dV, v_posed and A:
//   dv_posed = sum_s w_s R_s^T dV            -> bf16 hi (+lo) chunks of dvp (GEMM operand)
//   dA_j    += w_s [dV (x) p | dV]           -> per-slot register accumulators -> fp32 RED into dA_acc
//   dtransl += dV                            -> fp32 RED into dtr_acc
// ---------------------------------------------------------------------------------------------
constexpr int BWD_WARPS = 12;
constexpr int BWD_THREADS = BWD_WARPS * 32;
constexpr size_t BWD_SMEM = (size_t)BWD_WARPS * TTILE_WORDS * 4;

struct BwdState {
  float a0[9], a1[9], a2[9], a3[9];                      // rotation parts of the 4 cached transforms
  float d0[AELEMS], d1[AELEMS], d2[AELEMS], d3[AELEMS];  // their gradient accumulators
  uint32_t prev;                                         // plan word of the previous vertex (joint ids of the slots)
  float sx, sy, sz;
};

// 4 vertices: P (v_posed), staged dV from the lane's row -> q[12] (dv_posed rows), accumulators updated
__device__ __forceinline__ void skin_bwd4(const float (&P)[12], BwdState& s, const float4* __restrict__ A_g,
                                          float* __restrict__ dA_g, int lane, const uint32_t* __restrict__ meta,
                                          const float4* __restrict__ wts, uint32_t force, const float* row_in,
                                          float (&q)[12]) {
  const uint4 m4 = __ldg(reinterpret_cast<const uint4*>(meta));
  const uint32_t mts[4] = {m4.x | force, m4.y, m4.z, m4.w};
  float G[12];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float4 v = *reinterpret_cast<const float4*>(row_in + i * 4);
    G[i * 4] = v.x; G[i * 4 + 1] = v.y; G[i * 4 + 2] = v.z; G[i * 4 + 3] = v.w;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t mt = mts[i];
    const float4 w = __ldg(wts + i);
    if (mt & (0xFu << 20)) {
      const uint32_t pv = s.prev;
      if (mt & (1u << 20)) { flush_slot_g(s.d0, dA_g, pv & 31, lane); load_rot_g(s.a0, A_g, mt & 31, lane); }
      if (mt & (1u << 21)) { flush_slot_g(s.d1, dA_g, (pv >> 5) & 31, lane); load_rot_g(s.a1, A_g, (mt >> 5) & 31, lane); }
      if (mt & (1u << 22)) { flush_slot_g(s.d2, dA_g, (pv >> 10) & 31, lane); load_rot_g(s.a2, A_g, (mt >> 10) & 31, lane); }
      if (mt & (1u << 23)) { flush_slot_g(s.d3, dA_g, (pv >> 15) & 31, lane); load_rot_g(s.a3, A_g, (mt >> 15) & 31, lane); }
    }
    s.prev = mt;
    const float px = P[i * 3], py = P[i * 3 + 1], pz = P[i * 3 + 2];
    const float gx = G[i * 3], gy = G[i * 3 + 1], gz = G[i * 3 + 2];
    s.sx += gx; s.sy += gy; s.sz += gz;
    float qx = 0.f, qy = 0.f, qz = 0.f;
#define B200_SKIN_BWD(a, d, wk)                                                   \
  {                                                                               \
    const float hx = wk * gx, hy = wk * gy, hz = wk * gz;                         \
    qx = fmaf(a[0], hx, fmaf(a[3], hy, fmaf(a[6], hz, qx)));                      \
    qy = fmaf(a[1], hx, fmaf(a[4], hy, fmaf(a[7], hz, qy)));                      \
    qz = fmaf(a[2], hx, fmaf(a[5], hy, fmaf(a[8], hz, qz)));                      \
    d[0] = fmaf(hx, px, d[0]); d[1] = fmaf(hx, py, d[1]); d[2] = fmaf(hx, pz, d[2]); d[3] += hx;    \
    d[4] = fmaf(hy, px, d[4]); d[5] = fmaf(hy, py, d[5]); d[6] = fmaf(hy, pz, d[6]); d[7] += hy;    \
    d[8] = fmaf(hz, px, d[8]); d[9] = fmaf(hz, py, d[9]); d[10] = fmaf(hz, pz, d[10]); d[11] += hz; \
  }
    B200_SKIN_BWD(s.a0, s.d0, w.x)
    B200_SKIN_BWD(s.a1, s.d1, w.y)
    B200_SKIN_BWD(s.a2, s.d2, w.z)
    B200_SKIN_BWD(s.a3, s.d3, w.w)
#undef B200_SKIN_BWD
    q[i * 3] = qx; q[i * 3 + 1] = qy; q[i * 3 + 2] = qz;
  }
}

// close the accumulators of the current group: 4 slots -> dA, translation sums -> dtransl
__device__ __forceinline__ void bwd_close_group(BwdState& s, float* dA_g, float* dtr_g, int lane) {
  flush_slot_g(s.d0, dA_g, s.prev & 31, lane);
  flush_slot_g(s.d1, dA_g, (s.prev >> 5) & 31, lane);
  flush_slot_g(s.d2, dA_g, (s.prev >> 10) & 31, lane);
  flush_slot_g(s.d3, dA_g, (s.prev >> 15) & 31, lane);
  red_add(dtr_g + lane, s.sx);
  red_add(dtr_g + 32 + lane, s.sy);
  red_add(dtr_g + 64 + lane, s.sz);
  s.sx = s.sy = s.sz = 0.f;
}

__global__ void __launch_bounds__(BWD_THREADS, 1)
lbs_bwd_kernel(const float4* __restrict__ vpB, int S, const float4* __restrict__ A_blk, int b0, int nb, int ngroups,
               const float* __restrict__ grad_verts, int V, int ntiles, int vec_ok,
               const uint32_t* __restrict__ vmeta, const float4* __restrict__ vwts,
               __nv_bfloat16* __restrict__ dvp_hi, __nv_bfloat16* __restrict__ dvp_lo,
               float* __restrict__ dA_acc, float* __restrict__ dtr_acc) {
  extern __shared__ __align__(128) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* tile = smem + warp * TTILE_WORDS;
  const float* my_row = tile + lane * TROW;
  WarpRun run = make_run(ngroups, ntiles, blockIdx.x * BWD_WARPS + warp, gridDim.x * BWD_WARPS);
  if (run.item >= run.end) return;

  float P[12], Q[12];
  int g = run.g, t = run.t;
  BwdState st;
#pragma unroll
  for (int e = 0; e < AELEMS; ++e) st.d0[e] = st.d1[e] = st.d2[e] = st.d3[e] = 0.f;
  st.prev = 0u;
  st.sx = st.sy = st.sz = 0.f;
  int cur_g = -1;
  const float4* A_g = nullptr;
  float* dA_g = nullptr;
  const float4* vp_lane = nullptr;
  for (int item = run.item; item < run.end; ++item) {
    uint32_t force = 0u;
    if (g != cur_g) {
      if (cur_g >= 0) bwd_close_group(st, dA_g, dtr_acc + (size_t)cur_g * 96, lane);
      cur_g = g;
      A_g = A_blk + (size_t)g * (AG_WORDS / 4);
      dA_g = dA_acc + (size_t)g * AG_WORDS;
      vp_lane = vpB + (size_t)g * 32 + lane;
      force = 0xFu << 20;     // the accumulators are zero here, so the flushes this triggers add nothing
    }
    const int vbase = t * TILE_V;
    const uint32_t* meta = vmeta + vbase;
    const float4* wts = vwts + vbase;
    const float4* vp_t = vp_lane + (size_t)(t * 24) * S;
    fetch_vp4(P, vp_t, S);
    // ---- stage dV rows of this tile (384 contiguous bytes per body row) ----
    {
      const int nrows = min(32, nb - g * 32);
      const int ncols = min(TILE_V, V - vbase) * 3;
      const float* src0 = grad_verts + ((size_t)(b0 + g * 32) * V + vbase) * 3;
      if (!vec_ok) global_to_tile_scalar(tile, src0, (size_t)V * 3, nrows, ncols, lane);
      else if (nrows == 32 && ncols == 96) global_to_tile<true>(tile, src0, (size_t)V * 3, 32, 96, lane);
      else global_to_tile<false>(tile, src0, (size_t)V * 3, max(nrows, 0), ncols, lane);
    }
    __syncwarp();
    // ---- 4 x (4 + 4 vertices): v_posed prefetched one unit ahead; 8 vertices = 24 rows = 3 chunks of dvp ----
    const size_t chunk_stride = (size_t)S * 8;                                  // elements between 8-row chunks
    __nv_bfloat16* hi_p = dvp_hi + ((size_t)(t * 12) * S + (size_t)g * 32 + lane) * 8;
    __nv_bfloat16* lo_p = dvp_lo ? dvp_lo + ((size_t)(t * 12) * S + (size_t)g * 32 + lane) * 8 : nullptr;
#pragma unroll 1
    for (int u = 0; u < 4; ++u) {
      float qa[12], qb[12];
      fetch_vp4(Q, vp_t + (size_t)(u * 6 + 3) * S, S);
      skin_bwd4(P, st, A_g, dA_g, lane, meta + u * 8, wts + u * 8, u == 0 ? force : 0u, my_row + u * 24, qa);
      if (u < 3) fetch_vp4(P, vp_t + (size_t)(u * 6 + 6) * S, S);
      skin_bwd4(Q, st, A_g, dA_g, lane, meta + u * 8 + 4, wts + u * 8 + 4, 0u, my_row + u * 24 + 12, qb);
      const float c0[8] = {qa[0], qa[1], qa[2], qa[3], qa[4], qa[5], qa[6], qa[7]};
      const float c1[8] = {qa[8], qa[9], qa[10], qa[11], qb[0], qb[1], qb[2], qb[3]};
      const float c2[8] = {qb[4], qb[5], qb[6], qb[7], qb[8], qb[9], qb[10], qb[11]};
      store_dvp_chunk(c0, hi_p + (size_t)(u * 3) * chunk_stride, lo_p ? lo_p + (size_t)(u * 3) * chunk_stride : nullptr);
      store_dvp_chunk(c1, hi_p + (size_t)(u * 3 + 1) * chunk_stride, lo_p ? lo_p + (size_t)(u * 3 + 1) * chunk_stride : nullptr);
      store_dvp_chunk(c2, hi_p + (size_t)(u * 3 + 2) * chunk_stride, lo_p ? lo_p + (size_t)(u * 3 + 2) * chunk_stride : nullptr);
    }
    __syncwarp();
    if (++t == ntiles) { t = 0; ++g; }
  }
  bwd_close_group(st, dA_g, dtr_acc + (size_t)cur_g * 96, lane);
}

// ---------------------------------------------------------------------------------------------
static int run_grid(int ngroups, int ntiles, int warps, int num_sms) {
  const long long total = (long long)ngroups * ntiles;
  return (int)std::max<long long>(1, std::min<long long>(num_sms, (total + warps - 1) / warps));
}

int launch_lbs_fwd(const DevModel& m, const float* vpB, int S, const float* A_blk, int b0, int nb,
                   const float* transl, float* verts, int num_sms, cudaStream_t st) {
  if (nb <= 0) return 0;
  const int groups = (nb + 31) / 32;
  const int vec_ok = ((m.V & 1) == 0 && (reinterpret_cast<uintptr_t>(verts) & 7) == 0) ? 1 : 0;
  B200_CUDA_TRY(cudaFuncSetAttribute(lbs_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM));
  LaunchTimer _timer("lbs_fwd", st);
  lbs_fwd_kernel<<<run_grid(groups, m.ntiles, FWD_WARPS, num_sms), FWD_THREADS, FWD_SMEM, st>>>(
      reinterpret_cast<const float4*>(vpB), S, reinterpret_cast<const float4*>(A_blk), b0, nb, groups, transl, verts,
      m.V, m.ntiles, vec_ok, m.vmeta, m.vwts);
  B200_LAUNCH_CHECK("lbs_fwd");
  return 0;
}

// Sw = active slab width (multiple of 32, >= nb): rows of absent bodies get zero dvp
// dA_acc [S/32][288][32] and dtr_acc [S/32][3][32] must be zeroed by the caller; warps add into them
int launch_lbs_bwd(const DevModel& m, const float* vpB, int S, int Sw, const float* A_blk, int b0, int nb,
                   const float* grad_verts, __nv_bfloat16* dvp_hi, __nv_bfloat16* dvp_lo, float* dA_acc,
                   float* dtr_acc, int num_sms, cudaStream_t st) {
  const int groups = Sw / 32;
  const int vec_ok = ((m.V & 1) == 0 && (reinterpret_cast<uintptr_t>(grad_verts) & 7) == 0) ? 1 : 0;
  B200_CUDA_TRY(cudaFuncSetAttribute(lbs_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM));
  LaunchTimer _timer("lbs_bwd", st);
  lbs_bwd_kernel<<<run_grid(groups, m.ntiles, BWD_WARPS, num_sms), BWD_THREADS, BWD_SMEM, st>>>(
      reinterpret_cast<const float4*>(vpB), S, reinterpret_cast<const float4*>(A_blk), b0, nb, groups, grad_verts, m.V,
      m.ntiles, vec_ok, m.vmeta, m.vwts, dvp_hi, dvp_lo, dA_acc, dtr_acc);
  B200_LAUNCH_CHECK("lbs_bwd");
  return 0;
}

}  // namespace b200smpl
