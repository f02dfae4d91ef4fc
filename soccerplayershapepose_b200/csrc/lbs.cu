// Linear blend skinning of the 6890 vertices, forward and backward (SURVEY.md section 8 row a9;
// smplx.lbs.lbs tail: T = W.A ; v = T.[v_posed;1]).  HBM-roofline kernels.
//
// Mapping: lane = body.  The work is the flat list of (body group, HV-vertex item) units; every CTA owns a
// contiguous range of it, walks it one body group at a time and gives each warp a contiguous run of the
// group's items, so the load is balanced to one item per warp whatever the batch is.  Everything per-vertex (4 joint ids, 4 weights) is
// warp-uniform; everything per-body lives in registers: the 4 cached joint transforms ("slots") are
// (re)loaded -- three coalesced 512-byte float4 loads from A_blk, L1/L2 resident -- only where the packed
// plan says a slot's joint changes (8 times per 32 vertices on the SMPL mesh order).
//
// Data movement: v_posed comes from the blend GEMM group-blocked, [S/32][n/4][32][4]: a warp's run is one
// contiguous stream (512 bytes per float4 and chunk).  The tensors whose layout the
// caller fixes -- (B, 6890, 3) vertices / vertex gradients -- are transposed through a per-warp
// shared-memory tile [32 bodies][100]: the lane = body side uses 128-bit accesses on its own row, the
// global side moves each 384-byte row segment as 8-byte accesses (two rows per three warp
// instructions), every byte read or written exactly once with streaming accesses.  The backward writes
// dv_posed as bf16 hi/lo 16-byte chunks [S/128][n/8][128][8] (512 contiguous bytes per warp store, the operand
// layout of the gradient GEMM) and adds dA / dtransl into the slab accumulators with fp32 REDs.
#include "skin_common.cuh"

namespace b200smpl {

// Pipeline: every warp keeps TWO staging buffers and fills the one of item k+1 with cp.async (v_posed: 16
// bytes per lane and chunk into the lane's own row; vertex gradients: 8-byte pieces of the (B, V, 3) rows,
// zero-filled outside the batch / mesh) while it works on item k, so a whole item per warp (3-6 KB) is in
// flight at any time without holding registers.
template <int HV>
struct ItemShape {
  static constexpr int ROWS = HV * 3;             // blend rows (floats per body) of one item
  static constexpr int HROW = ROWS + 4;           // staging row pitch: 16-byte aligned, HROW / 4 odd -> a quarter-warp
  static constexpr int TILE_WORDS = 32 * HROW;    // of 128-bit row accesses (lane = row) covers all banks
  static constexpr int NCH4 = ROWS / 4;           // float4 chunks of v_posed per body
  static constexpr int NCH8 = ROWS / 8;           // 8-row chunks of dvp per body
  static constexpr int PAIRS = ROWS / 2;          // 8-byte pieces per body row
  static constexpr int STASH_WORDS = HV * 5;      // plan words of one item: HV metas + HV float4 weights
  static_assert((HROW / 4) % 2 == 1 && ROWS % 8 == 0, "item shape");
};

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global.L2::256B [%0], [%1], 16;" ::"r"(smem_addr(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8_zfill(void* dst_smem, const void* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global.L2::256B [%0], [%1], 8, %2;" ::"r"(smem_addr(dst_smem)), "l"(src), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// contiguous range of the flat (group, item) list owned by this CTA; it is walked one group at a time
struct CtaRange {
  int begin, end;
};
__device__ __forceinline__ CtaRange make_cta_range(int ngroups, int nitems) {
  const long long total = (long long)ngroups * nitems;
  CtaRange r;
  r.begin = (int)(total * blockIdx.x / gridDim.x);
  r.end = (int)(total * (blockIdx.x + 1) / gridDim.x);
  return r;
}

// plan words of one item, held by lanes 0..HV-1 between the global load and the stash store
struct PlanRegs {
  uint32_t m;
  float4 w;
};
template <int HV>
__device__ __forceinline__ PlanRegs plan_load(const uint32_t* __restrict__ vmeta, const float4* __restrict__ vwts,
                                              int vbase, int lane) {
  PlanRegs r;
  r.m = 0u;
  r.w = make_float4(0.f, 0.f, 0.f, 0.f);
  if (lane < HV) {
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r.m) : "l"(vmeta + vbase + lane));
    r.w = ld_stream4(vwts + vbase + lane);
  }
  return r;
}
template <int HV>
__device__ __forceinline__ void plan_store(uint32_t* stash, const PlanRegs& r, int lane) {
  if (lane < HV) {
    stash[lane] = r.m;
    reinterpret_cast<float4*>(stash + HV)[lane] = r.w;
  }
}

// v_posed of one item: NCH4 float4 per body, straight into the lane's own row (no cross-lane traffic)
template <int HV>
__device__ __forceinline__ void issue_vp_rows(float* my_row, const float4* __restrict__ vp_item_lane) {
#pragma unroll
  for (int c = 0; c < ItemShape<HV>::NCH4; ++c) cp_async16(my_row + c * 4, vp_item_lane + c * 32);
}
// ... or into a dense [chunk][lane] float4 array
template <int HV>
__device__ __forceinline__ void issue_vp_dense(float4* vbuf, const float4* __restrict__ vp_item_lane, int lane) {
#pragma unroll
  for (int c = 0; c < ItemShape<HV>::NCH4; ++c) cp_async16(vbuf + c * 32 + lane, vp_item_lane + c * 32);
}
// 32 row segments of a (B, V, 3) tensor <-> staging tile in 8-byte pieces.  The 32 * PAIRS pieces are taken
// 32 per warp instruction: piece = k * 32 + lane -> row = piece / PAIRS, column pair = piece % PAIRS.  Since
// 96 = RPP * PAIRS the (row offset, column) of a lane repeats every 3 instructions, RPP rows further down, so
// a lane keeps 3 global pointers and adds a constant stride.
template <int HV>
__device__ __forceinline__ void issue_rows_full(float* tile, const float* __restrict__ src0, size_t row_stride, int lane) {
  using SH = ItemShape<HV>;
  constexpr int RPP = 96 / SH::PAIRS, J = 32 / RPP;
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
}
// general version: zero-filled outside [nrows) x [ncols)
template <int HV>
__device__ __forceinline__ void issue_rows(float* tile, const float* __restrict__ src0, size_t row_stride, int nrows,
                                           int ncols, int lane) {
  using SH = ItemShape<HV>;
  if (nrows == 32 && ncols == SH::ROWS) {
    issue_rows_full<HV>(tile, src0, row_stride, lane);
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

// rows of the staging tile -> row segments of a (B, V, 3) tensor (8-byte stores)
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
template <int HV>
__device__ __forceinline__ void global_to_tile_scalar(float* tile, const float* src0, size_t row_stride, int nrows,
                                                      int ncols, int lane) {
  for (int r = 0; r < 32; ++r)
    for (int c = lane; c < ItemShape<HV>::ROWS; c += 32)
      tile[r * ItemShape<HV>::HROW + c] = (r < nrows && c < ncols) ? src0[(size_t)r * row_stride + c] : 0.f;
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
#ifndef B200_FWD_HV
#define B200_FWD_HV 16
#endif
#ifndef B200_FWD_WARPS
#define B200_FWD_WARPS 12
#endif
constexpr int FWD_HV = B200_FWD_HV;
constexpr int FWD_WARPS = B200_FWD_WARPS;
constexpr int FWD_THREADS = FWD_WARPS * 32;
constexpr int FWD_WARP_WORDS = 2 * ItemShape<FWD_HV>::TILE_WORDS + 2 * ItemShape<FWD_HV>::STASH_WORDS;
constexpr size_t FWD_SMEM = (size_t)(AG_WORDS + FWD_WARPS * FWD_WARP_WORDS) * 4 + 16;

struct Slots {
  float a0[AELEMS], a1[AELEMS], a2[AELEMS], a3[AELEMS];
};

// 4 vertices, in place on the lane's own row: v_posed -> skinned coordinates
__device__ __forceinline__ void skin_fwd4(Slots& s, const float* A_s, int lane, const uint32_t* meta_s,
                                          const float4* wts_s, uint32_t force, float tx, float ty, float tz,
                                          float* row_io) {
  const uint4 m4 = *reinterpret_cast<const uint4*>(meta_s);
  const uint32_t mts[4] = {m4.x | force, m4.y, m4.z, m4.w};
  float4 ws[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) ws[i] = wts_s[i];
  float P[12], o[12];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float4 v = *reinterpret_cast<const float4*>(row_io + i * 4);
    P[i * 4] = v.x; P[i * 4 + 1] = v.y; P[i * 4 + 2] = v.z; P[i * 4 + 3] = v.w;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t mt = mts[i];
    const float4 w = ws[i];
    if (mt & (0xFu << 20)) {
      if (mt & (1u << 20)) load_slot(s.a0, A_s, mt & 31, lane);
      if (mt & (1u << 21)) load_slot(s.a1, A_s, (mt >> 5) & 31, lane);
      if (mt & (1u << 22)) load_slot(s.a2, A_s, (mt >> 10) & 31, lane);
      if (mt & (1u << 23)) load_slot(s.a3, A_s, (mt >> 15) & 31, lane);
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
  for (int i = 0; i < 3; ++i)
    *reinterpret_cast<float4*>(row_io + i * 4) = make_float4(o[i * 4], o[i * 4 + 1], o[i * 4 + 2], o[i * 4 + 3]);
}

__global__ void __launch_bounds__(FWD_THREADS, 1)
lbs_fwd_kernel(const float4* __restrict__ vpB, int nc4, const float4* __restrict__ A_blk, int b0, int nb, int ngroups,
               const float* __restrict__ transl, float* __restrict__ verts, int V, int nitems, int vec_ok,
               const uint32_t* __restrict__ vmeta, const float4* __restrict__ vwts) {
  constexpr int HV = FWD_HV;
  using SH = ItemShape<HV>;
  extern __shared__ __align__(128) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* A_s = smem;                                               // [24][3][32] float4: the group's transforms
  float* wbase = smem + AG_WORDS + warp * FWD_WARP_WORDS;
  float* tiles = wbase;                                            // [2][32][HROW]
  uint32_t* stash = reinterpret_cast<uint32_t*>(wbase + 2 * SH::TILE_WORDS);   // [2][STASH_WORDS]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + AG_WORDS + FWD_WARPS * FWD_WARP_WORDS);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t a_phase = 0;
  const CtaRange cta = make_cta_range(ngroups, nitems);
  for (int seg0 = cta.begin; seg0 < cta.end;) {
    // ---- one body group at a time per CTA: its 36 KB of transforms come in with one TMA bulk copy ----
    const int g = seg0 / nitems;
    if (threadIdx.x == 0) {
      mbar_expect_tx(bar, AG_WORDS * 4);
      bulk_g2s(A_s, A_blk + (size_t)g * (AG_WORDS / 4), AG_WORDS * 4, bar);
    }
    const int seg1 = min(cta.end, (g + 1) * nitems);
    const int len = seg1 - seg0;
    const int it0 = seg0 - g * nitems + len * warp / FWD_WARPS;    // this warp's contiguous items of the group
    const int it1 = seg0 - g * nitems + len * (warp + 1) / FWD_WARPS;
    if (it0 < it1) {
      const float4* vp_g = vpB + (size_t)g * nc4 * 32 + lane;
      float tx = 0.f, ty = 0.f, tz = 0.f;
      if (transl != nullptr && g * 32 + lane < nb) {
        const float* tp = transl + (size_t)(b0 + g * 32 + lane) * 3;
        tx = tp[0]; ty = tp[1]; tz = tp[2];
      }
      const int nrows = min(32, nb - g * 32);
      float* v_g = verts + (size_t)(b0 + g * 32) * V * 3;
      plan_store<HV>(stash, plan_load<HV>(vmeta, vwts, it0 * HV, lane), lane);
      issue_vp_rows<HV>(tiles + lane * SH::HROW, vp_g + (size_t)(it0 * SH::NCH4) * 32);
      cp_async_commit();
      Slots sl;
      int buf = 0;
      __syncwarp();
      mbar_wait(bar, a_phase);
      for (int t = it0; t < it1; ++t) {
        const bool more = t + 1 < it1;
        float* tile = tiles + buf * SH::TILE_WORDS;
        float* my_row = tile + lane * SH::HROW;
        PlanRegs nplan;
        if (more) {                                                // next item: v_posed into the other buffer
          nplan = plan_load<HV>(vmeta, vwts, (t + 1) * HV, lane);
          issue_vp_rows<HV>(tiles + (buf ^ 1) * SH::TILE_WORDS + lane * SH::HROW, vp_g + (size_t)((t + 1) * SH::NCH4) * 32);
        }
        cp_async_commit();
        cp_async_wait<1>();                                        // this item's rows have landed (own row only)
        const uint32_t* meta_s = stash + buf * SH::STASH_WORDS;
        const float4* wts_s = reinterpret_cast<const float4*>(meta_s + HV);
#pragma unroll 1
        for (int u = 0; u < HV / 4; ++u)
          skin_fwd4(sl, A_s, lane, meta_s + u * 4, wts_s + u * 4, (u == 0 && t == it0) ? (0xFu << 20) : 0u, tx, ty, tz,
                    my_row + u * 12);
        if (more) plan_store<HV>(stash + (buf ^ 1) * SH::STASH_WORDS, nplan, lane);
        __syncwarp();
        // flush: each body row of the item is ROWS contiguous floats of the (B, V, 3) output
        {
          const int vbase = t * HV;
          const int ncols = max(0, min(HV, V - vbase)) * 3;
          float* dst0 = v_g + (size_t)vbase * 3;
          if (!vec_ok) tile_to_global_scalar<HV>(tile, dst0, (size_t)V * 3, nrows, ncols, lane);
          else tile_to_global<HV>(tile, dst0, (size_t)V * 3, nrows, ncols, lane);
        }
        __syncwarp();                                              // rows are re-filled two items later
        buf ^= 1;
      }
    }
    seg0 = seg1;
    a_phase ^= 1;
    if (seg0 < cta.end) __syncthreads();                           // group switch is CTA-wide: A_s is re-filled
  }
}

// ---------------------------------------------------------------------------------------------
// backward:  given dV, v_posed and A:
//   dv_posed = sum_s w_s R_s^T dV            -> bf16 hi (+lo) chunks of dvp (GEMM operand)
//   dA_j    += w_s [dV (x) p | dV]           -> per-slot register accumulators -> fp32 RED into dA_acc
//   dtransl += dV                            -> fp32 RED into dtr_acc
// ---------------------------------------------------------------------------------------------
#ifndef B200_BWD_HV
#define B200_BWD_HV 8
#endif
#ifndef B200_BWD_WARPS
#define B200_BWD_WARPS 12
#endif
constexpr int BWD_HV = B200_BWD_HV;
constexpr int BWD_WARPS = B200_BWD_WARPS;
constexpr int BWD_THREADS = BWD_WARPS * 32;
constexpr int BWD_WARP_WORDS = 2 * ItemShape<BWD_HV>::TILE_WORDS + 2 * ItemShape<BWD_HV>::NCH4 * 128 +
                               2 * ItemShape<BWD_HV>::STASH_WORDS;
constexpr size_t BWD_SMEM = (size_t)(AG_WORDS + BWD_WARPS * BWD_WARP_WORDS) * 4 + 16;

struct BwdState {
  float a0[9], a1[9], a2[9], a3[9];                      // rotation parts of the 4 cached transforms
  float d0[AELEMS], d1[AELEMS], d2[AELEMS], d3[AELEMS];  // their gradient accumulators
  uint32_t prev;                                         // plan word of the previous vertex (joint ids of the slots)
  float sx, sy, sz;
};

// 4 vertices: v_posed from the dense buffer, staged dV of the lane's row -> dv_posed written back in place
// (fp32), accumulators updated
__device__ __forceinline__ void skin_bwd4(BwdState& s, const float* A_s, float* __restrict__ dA_g,
                                          int lane, const uint32_t* meta_s, const float4* wts_s, uint32_t force,
                                          const float4* vp_s, float* row_io) {
  const uint4 m4 = *reinterpret_cast<const uint4*>(meta_s);
  const uint32_t mts[4] = {m4.x | force, m4.y, m4.z, m4.w};
  float4 ws[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) ws[i] = wts_s[i];
  float P[12], G[12], q[12];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float4 v = vp_s[i * 32];
    P[i * 4] = v.x; P[i * 4 + 1] = v.y; P[i * 4 + 2] = v.z; P[i * 4 + 3] = v.w;
    const float4 h = *reinterpret_cast<const float4*>(row_io + i * 4);
    G[i * 4] = h.x; G[i * 4 + 1] = h.y; G[i * 4 + 2] = h.z; G[i * 4 + 3] = h.w;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t mt = mts[i];
    const float4 w = ws[i];
    if (mt & (0xFu << 20)) {
      const uint32_t pv = s.prev;
      if (mt & (1u << 20)) { flush_slot_g(s.d0, dA_g, pv & 31, lane); load_rot(s.a0, A_s, mt & 31, lane); }
      if (mt & (1u << 21)) { flush_slot_g(s.d1, dA_g, (pv >> 5) & 31, lane); load_rot(s.a1, A_s, (mt >> 5) & 31, lane); }
      if (mt & (1u << 22)) { flush_slot_g(s.d2, dA_g, (pv >> 10) & 31, lane); load_rot(s.a2, A_s, (mt >> 10) & 31, lane); }
      if (mt & (1u << 23)) { flush_slot_g(s.d3, dA_g, (pv >> 15) & 31, lane); load_rot(s.a3, A_s, (mt >> 15) & 31, lane); }
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
#pragma unroll
  for (int i = 0; i < 3; ++i)
    *reinterpret_cast<float4*>(row_io + i * 4) = make_float4(q[i * 4], q[i * 4 + 1], q[i * 4 + 2], q[i * 4 + 3]);
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
lbs_bwd_kernel(const float4* __restrict__ vpB, int nc4, const float4* __restrict__ A_blk, int b0, int nb, int ngroups,
               const float* __restrict__ grad_verts, int V, int nitems, int vec_ok,
               const uint32_t* __restrict__ vmeta, const float4* __restrict__ vwts,
               __nv_bfloat16* __restrict__ dvp_hi, __nv_bfloat16* __restrict__ dvp_lo,
               float* __restrict__ dA_acc, float* __restrict__ dtr_acc) {
  constexpr int HV = BWD_HV;
  using SH = ItemShape<HV>;
  extern __shared__ __align__(128) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* A_s = smem;                                                           // the group's transforms
  float* wbase = smem + AG_WORDS + warp * BWD_WARP_WORDS;
  float* tiles = wbase;                                                        // [2][32][HROW]   dV in / dv_posed out
  float4* vbufs = reinterpret_cast<float4*>(wbase + 2 * SH::TILE_WORDS);       // [2][NCH4][32]   v_posed
  uint32_t* stash = reinterpret_cast<uint32_t*>(wbase + 2 * SH::TILE_WORDS + 2 * SH::NCH4 * 128);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + AG_WORDS + BWD_WARPS * BWD_WARP_WORDS);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t a_phase = 0;
  const size_t row_stride = (size_t)V * 3;
  const CtaRange cta = make_cta_range(ngroups, nitems);
  for (int seg0 = cta.begin; seg0 < cta.end;) {
    const int g = seg0 / nitems;
    if (threadIdx.x == 0) {
      mbar_expect_tx(bar, AG_WORDS * 4);
      bulk_g2s(A_s, A_blk + (size_t)g * (AG_WORDS / 4), AG_WORDS * 4, bar);
    }
    const int seg1 = min(cta.end, (g + 1) * nitems);
    const int len = seg1 - seg0;
    const int it0 = seg0 - g * nitems + len * warp / BWD_WARPS;
    const int it1 = seg0 - g * nitems + len * (warp + 1) / BWD_WARPS;
    if (it0 < it1) {
      float* dA_g = dA_acc + (size_t)g * AG_WORDS;
      const float4* vp_g = vpB + (size_t)g * nc4 * 32 + lane;
      const float* dv_g = grad_verts + (size_t)(b0 + g * 32) * V * 3;
      const int nrows = max(0, min(32, nb - g * 32));
      plan_store<HV>(stash, plan_load<HV>(vmeta, vwts, it0 * HV, lane), lane);
      issue_vp_dense<HV>(vbufs, vp_g + (size_t)(it0 * SH::NCH4) * 32, lane);
      if (vec_ok) issue_rows<HV>(tiles, dv_g + (size_t)it0 * HV * 3, row_stride, nrows, max(0, min(HV, V - it0 * HV)) * 3, lane);
      cp_async_commit();
      BwdState st;
#pragma unroll
      for (int e = 0; e < AELEMS; ++e) st.d0[e] = st.d1[e] = st.d2[e] = st.d3[e] = 0.f;
      st.prev = 0u;
      st.sx = st.sy = st.sz = 0.f;
      int buf = 0;
      mbar_wait(bar, a_phase);
      for (int t = it0; t < it1; ++t) {
        const bool more = t + 1 < it1;
        float* tile = tiles + buf * SH::TILE_WORDS;
        float* my_row = tile + lane * SH::HROW;
        const float4* vp_s = vbufs + buf * (SH::NCH4 * 32) + lane;
        PlanRegs nplan;
        if (more) {
          nplan = plan_load<HV>(vmeta, vwts, (t + 1) * HV, lane);
          issue_vp_dense<HV>(vbufs + (buf ^ 1) * (SH::NCH4 * 32), vp_g + (size_t)((t + 1) * SH::NCH4) * 32, lane);
          if (vec_ok)
            issue_rows<HV>(tiles + (buf ^ 1) * SH::TILE_WORDS, dv_g + (size_t)(t + 1) * HV * 3, row_stride, nrows,
                           max(0, min(HV, V - (t + 1) * HV)) * 3, lane);
        }
        cp_async_commit();
        cp_async_wait<1>();
        if (!vec_ok)
          global_to_tile_scalar<HV>(tile, dv_g + (size_t)t * HV * 3, row_stride, nrows, max(0, min(HV, V - t * HV)) * 3, lane);
        __syncwarp();                                              // the dV rows were written by other lanes
        const uint32_t* meta_s = stash + buf * SH::STASH_WORDS;
        const float4* wts_s = reinterpret_cast<const float4*>(meta_s + HV);
        // the first vertex of the run (re)loads all four slots; the accumulators are zero there
#pragma unroll 1
        for (int u = 0; u < HV / 4; ++u)
          skin_bwd4(st, A_s, dA_g, lane, meta_s + u * 4, wts_s + u * 4, (u == 0 && t == it0) ? (0xFu << 20) : 0u,
                    vp_s + u * 96, my_row + u * 12);
        if (more) plan_store<HV>(stash + (buf ^ 1) * SH::STASH_WORDS, nplan, lane);
        // ---- the lane's gradient rows -> bf16 hi/lo chunks of dvp (512 contiguous bytes per warp store) ----
        {
          // dvp [S/128][n_pad/8][128][8]: chunk stride 1024 elements, this body at ((g & 3) * 32 + lane) * 8
          const size_t o0 = (((size_t)(g >> 2) * (nc4 >> 1) + t * SH::NCH8) * 128 + (g & 3) * 32 + lane) * 8;
#pragma unroll
          for (int c = 0; c < SH::NCH8; ++c) {
            const float4 x = *reinterpret_cast<const float4*>(my_row + c * 8);
            const float4 y = *reinterpret_cast<const float4*>(my_row + c * 8 + 4);
            const float ch[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
            store_dvp_chunk(ch, dvp_hi + o0 + c * 1024, dvp_lo ? dvp_lo + o0 + c * 1024 : nullptr);
          }
        }
        __syncwarp();                                              // buffers are re-filled two items later
        buf ^= 1;
      }
      bwd_close_group(st, dA_g, dtr_acc + (size_t)g * 96, lane);
    }
    seg0 = seg1;
    a_phase ^= 1;
    if (seg0 < cta.end) __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
static int run_grid(int ngroups, int nitems, int warps, int num_sms) {
  const long long total = (long long)ngroups * nitems;
  return (int)std::max<long long>(1, std::min<long long>(num_sms, (total + warps - 1) / warps));
}

int launch_lbs_fwd(const DevModel& m, const float* vpB, int S, const float* A_blk, int b0, int nb,
                   const float* transl, float* verts, int num_sms, cudaStream_t st) {
  if (nb <= 0) return 0;
  const int groups = (nb + 31) / 32;
  const int nitems = m.ntiles * (TILE_V / FWD_HV);
  const int vec_ok = ((m.V & 1) == 0 && (reinterpret_cast<uintptr_t>(verts) & 7) == 0) ? 1 : 0;
  B200_CUDA_TRY(cudaFuncSetAttribute(lbs_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM));
  LaunchTimer _timer("lbs_fwd", st);
  lbs_fwd_kernel<<<run_grid(groups, nitems, FWD_WARPS, num_sms), FWD_THREADS, FWD_SMEM, st>>>(
      reinterpret_cast<const float4*>(vpB), m.n_pad / 4, reinterpret_cast<const float4*>(A_blk), b0, nb, groups, transl, verts,
      m.V, nitems, vec_ok, m.vmeta, m.vwts);
  B200_LAUNCH_CHECK("lbs_fwd");
  return 0;
}

// Sw = active slab width (multiple of 32, >= nb): rows of absent bodies get zero dvp
// dA_acc [S/32][288][32] and dtr_acc [S/32][3][32] must be zeroed by the caller; warps add into them
int launch_lbs_bwd(const DevModel& m, const float* vpB, int S, int Sw, const float* A_blk, int b0, int nb,
                   const float* grad_verts, __nv_bfloat16* dvp_hi, __nv_bfloat16* dvp_lo, float* dA_acc,
                   float* dtr_acc, int num_sms, cudaStream_t st) {
  const int groups = Sw / 32;
  const int nitems = m.ntiles * (TILE_V / BWD_HV);
  const int vec_ok = ((m.V & 1) == 0 && (reinterpret_cast<uintptr_t>(grad_verts) & 7) == 0) ? 1 : 0;
  B200_CUDA_TRY(cudaFuncSetAttribute(lbs_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM));
  LaunchTimer _timer("lbs_bwd", st);
  lbs_bwd_kernel<<<run_grid(groups, nitems, BWD_WARPS, num_sms), BWD_THREADS, BWD_SMEM, st>>>(
      reinterpret_cast<const float4*>(vpB), m.n_pad / 4, reinterpret_cast<const float4*>(A_blk), b0, nb, groups, grad_verts, m.V,
      nitems, vec_ok, m.vmeta, m.vwts, dvp_hi, dvp_lo, dA_acc, dtr_acc);
  B200_LAUNCH_CHECK("lbs_bwd");
  return 0;
}

}  // namespace b200smpl
