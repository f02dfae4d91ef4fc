// Linear blend skinning of the 6890 vertices, forward and backward (SURVEY.md section 8 row a9;
// smplx.lbs.lbs tail: T = W.A ; v = T.[v_posed;1]).  HBM-roofline kernels.
//
// Mapping: lane = body, a warp owns a group of 32 bodies and walks 32-vertex tiles.  Everything
// per-vertex (4 joint ids, 4 weights) is warp-uniform; everything per-body (the 4 cached joint
// transforms) lives in registers and is re-loaded from the CTA's shared-memory copy of the group's
// A[24][12][32] (one 36 KB TMA bulk copy per CTA) only when the packed plan says a slot's joint
// changes.
//
// Data movement: v_posed comes from the blend GEMM in a group-blocked layout -- one contiguous
// 12 KB chunk [96 rows][32 bodies] per (tile, group) -- and is read straight into registers, 8
// vertices (24 coalesced 128 B lines per warp) ahead of the arithmetic.  The tensors whose layout
// the caller fixes ((B, 6890, 3) vertices / vertex gradients, K-major bf16 dv_posed rows for the
// gradient GEMM) are transposed through a per-warp shared-memory tile [96][33] (conflict-free for
// lane = body and for lane = column), so global accesses are contiguous 384 B (192 B bf16) row
// segments, each written exactly once with streaming stores.  12-14 warps per SM (1 CTA / SM).
#include "skin_common.cuh"

namespace b200smpl {

constexpr int PF = 8;   // vertices per register prefetch unit (24 loads in flight per lane)

// fetch the 3 v_posed rows of PF consecutive plan entries (processing order) from the chunk
__device__ __forceinline__ void fetch_vp(float (&P)[PF * 3], const float* __restrict__ chunk_lane,
                                         const uint32_t* __restrict__ meta) {
#pragma unroll
  for (int i = 0; i < PF; ++i) {
    const int c = ((__ldg(meta + i) >> 24) & 31) * 3;
    P[i * 3 + 0] = ld_stream(chunk_lane + c * 32);
    P[i * 3 + 1] = ld_stream(chunk_lane + c * 32 + 32);
    P[i * 3 + 2] = ld_stream(chunk_lane + c * 32 + 64);
  }
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
#ifndef B200_FWD_WARPS
#define B200_FWD_WARPS 12
#endif
constexpr int FWD_WARPS = B200_FWD_WARPS;
constexpr int FWD_THREADS = FWD_WARPS * 32;
constexpr size_t FWD_SMEM = (size_t)(AG_WORDS + FWD_WARPS * TTILE_WORDS) * 4 + 16;

struct Slots {
  float a0[AELEMS], a1[AELEMS], a2[AELEMS], a3[AELEMS];
};

__device__ __forceinline__ void skin_fwd(const float (&P)[PF * 3], Slots& s, const float* A_s, int lane,
                                         const uint32_t* __restrict__ meta, const float4* __restrict__ wts,
                                         uint32_t force, float tx, float ty, float tz, float* out_lane) {
  uint32_t mts[PF];
  float4 ws[PF];
#pragma unroll
  for (int i = 0; i < PF; ++i) {               // warp-uniform plan words of the whole unit, issued together
    mts[i] = __ldg(meta + i);
    ws[i] = __ldg(wts + i);
  }
#pragma unroll
  for (int i = 0; i < PF; ++i) {
    const uint32_t mt = mts[i] | (i == 0 ? force : 0u);
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
    float* o = out_lane + ((mt >> 24) & 31) * (3 * TPITCH);
    o[0] = ox;
    o[TPITCH] = oy;
    o[2 * TPITCH] = oz;
  }
}

__global__ void __launch_bounds__(FWD_THREADS, 1)
lbs_fwd_kernel(const float* __restrict__ vpB, int G, const float* __restrict__ A_blk, int b0, int nb,
               const float* __restrict__ transl, float* __restrict__ verts, int V, int ntiles,
               int tiles_per_split, const uint32_t* __restrict__ vmeta, const float4* __restrict__ vwts) {
  extern __shared__ __align__(128) float smem[];
  float* A_s = smem;                                         // [288][32]
  float* tiles = smem + AG_WORDS;                            // [FWD_WARPS][96][33]
  uint64_t* bar = reinterpret_cast<uint64_t*>(tiles + FWD_WARPS * TTILE_WORDS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.x;                                  // body group inside the slab
  const int col0 = g * 32;
  const int gb0 = b0 + col0;                                 // global index of the group's first body
  const int tile_begin = blockIdx.y * tiles_per_split;
  const int tile_end = min(ntiles, tile_begin + tiles_per_split);
  if (threadIdx.x == 0) fetch_group_transforms(A_s, A_blk, g, bar);

  float P[PF * 3], Q[PF * 3];
  int tile = tile_begin + warp;
  const float* chunk0 = vpB + (size_t)g * CHUNK_WORDS + lane;
  const size_t tstride = (size_t)G * CHUNK_WORDS;
  if (tile < tile_end) fetch_vp(P, chunk0 + tile * tstride, vmeta + tile * TILE_V);
  float tx = 0.f, ty = 0.f, tz = 0.f;
  if (transl != nullptr && col0 + lane < nb) {
    const float* t = transl + (size_t)(gb0 + lane) * 3;
    tx = t[0]; ty = t[1]; tz = t[2];
  }
  __syncthreads();                                           // barrier init visible to every waiter
  mbar_wait(bar, 0);

  float* out_s = tiles + warp * TTILE_WORDS;
  float* out_lane = out_s + lane;
  const int nrows_valid = min(32, nb - col0);                // bodies of this group that exist
  Slots sl;
  for (; tile < tile_end; tile += FWD_WARPS) {
    const int vbase = tile * TILE_V;
    const uint32_t* meta = vmeta + vbase;
    const float4* wts = vwts + vbase;
    const float* chunk = chunk0 + tile * tstride;
    fetch_vp(Q, chunk, meta + PF);
    skin_fwd(P, sl, A_s, lane, meta, wts, 0xFu << 20, tx, ty, tz, out_lane);
    fetch_vp(P, chunk, meta + 2 * PF);
    skin_fwd(Q, sl, A_s, lane, meta + PF, wts + PF, 0u, tx, ty, tz, out_lane);
    fetch_vp(Q, chunk, meta + 3 * PF);
    skin_fwd(P, sl, A_s, lane, meta + 2 * PF, wts + 2 * PF, 0u, tx, ty, tz, out_lane);
    const int next = tile + FWD_WARPS;
    if (next < tile_end) fetch_vp(P, chunk0 + next * tstride, vmeta + next * TILE_V);
    skin_fwd(Q, sl, A_s, lane, meta + 3 * PF, wts + 3 * PF, 0u, tx, ty, tz, out_lane);
    __syncwarp();
    // flush: each body row of the tile is 96 contiguous floats of the (B, V, 3) output
    const int ncols = min(TILE_V, V - vbase) * 3;
    float* dst0 = verts + ((size_t)gb0 * V + vbase) * 3 + lane;
    const float* src = out_s + lane * TPITCH;
    if (ncols == TILE_V * 3) {
#pragma unroll 4
      for (int r = 0; r < nrows_valid; ++r) {
        float* dst = dst0 + (size_t)r * V * 3;
        st_stream(dst, src[r]);
        st_stream(dst + 32, src[32 * TPITCH + r]);
        st_stream(dst + 64, src[64 * TPITCH + r]);
      }
    } else {
      for (int r = 0; r < nrows_valid; ++r)
        for (int c = lane; c < ncols; c += 32) dst0[(size_t)r * V * 3 + c - lane] = out_s[c * TPITCH + r];
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// backward:  given dV, recomputed v_posed and A:
//   dv_posed = sum_s w_s R_s^T dV            -> bf16 hi (+lo) rows dvp[body][3v+k]  (GEMM operand)
//   dA_j    += w_s [dV (x) p | dV]           -> per-CTA shared accumulators -> dA_part[split]
//   dtransl += dV
// ---------------------------------------------------------------------------------------------
constexpr int BWD_WARPS = 12;
constexpr int BWD_THREADS = BWD_WARPS * 32;
constexpr size_t BWD_SMEM = (size_t)(2 * AG_WORDS + 96 + BWD_WARPS * TTILE_WORDS) * 4 + 16;

constexpr int BPF = 4;   // backward prefetch unit: 4 vertices (12 loads in flight per lane) keeps the loop body in the I-cache

struct BwdState {
  float a0[9], a1[9], a2[9], a3[9];                      // rotation parts of the 4 cached transforms
  float d0[AELEMS], d1[AELEMS], d2[AELEMS], d3[AELEMS];  // their gradient accumulators
  int j0, j1, j2, j3;
  float sx, sy, sz;
};

__device__ __forceinline__ void fetch_vp4(float (&P)[BPF * 3], const float* __restrict__ chunk_lane,
                                          const uint32_t* __restrict__ meta) {
#pragma unroll
  for (int i = 0; i < BPF; ++i) {
    const int c = ((__ldg(meta + i) >> 24) & 31) * 3;
    P[i * 3 + 0] = ld_stream(chunk_lane + c * 32);
    P[i * 3 + 1] = ld_stream(chunk_lane + c * 32 + 32);
    P[i * 3 + 2] = ld_stream(chunk_lane + c * 32 + 64);
  }
}

__device__ __forceinline__ void skin_bwd(const float (&P)[BPF * 3], BwdState& s, const float* A_s, float* dA_s,
                                         int lane, const uint32_t* __restrict__ meta,
                                         const float4* __restrict__ wts, bool first, float* g_lane) {
  uint32_t* g_lane_u = reinterpret_cast<uint32_t*>(g_lane);
#pragma unroll
  for (int i = 0; i < BPF; ++i) {
    const bool force = first && i == 0;
    const uint32_t mt = __ldg(meta + i) | (force ? (0xFu << 20) : 0u);
    const int o = ((mt >> 24) & 31) * (3 * TPITCH);
    if (!(mt & VMETA_VALID)) {                           // padded vertex: its dvp columns must be 0
      g_lane_u[o] = 0u;
      g_lane_u[o + TPITCH] = 0u;
      g_lane_u[o + 2 * TPITCH] = 0u;
      continue;
    }
    const float4 w = __ldg(wts + i);
    if (mt & (0xFu << 20)) {
      if (mt & (1u << 20)) { if (!force) flush_slot(s.d0, dA_s, s.j0, lane); s.j0 = mt & 31; load_rot(s.a0, A_s, s.j0, lane); }
      if (mt & (1u << 21)) { if (!force) flush_slot(s.d1, dA_s, s.j1, lane); s.j1 = (mt >> 5) & 31; load_rot(s.a1, A_s, s.j1, lane); }
      if (mt & (1u << 22)) { if (!force) flush_slot(s.d2, dA_s, s.j2, lane); s.j2 = (mt >> 10) & 31; load_rot(s.a2, A_s, s.j2, lane); }
      if (mt & (1u << 23)) { if (!force) flush_slot(s.d3, dA_s, s.j3, lane); s.j3 = (mt >> 15) & 31; load_rot(s.a3, A_s, s.j3, lane); }
    }
    const float px = P[i * 3], py = P[i * 3 + 1], pz = P[i * 3 + 2];
    const float gx = g_lane[o], gy = g_lane[o + TPITCH], gz = g_lane[o + 2 * TPITCH];
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
    g_lane_u[o] = pack_hi_lo(qx);
    g_lane_u[o + TPITCH] = pack_hi_lo(qy);
    g_lane_u[o + 2 * TPITCH] = pack_hi_lo(qz);
  }
}

// stage 8 body rows x 96 columns of dV into the transposition tile; FULL: no bounds needed
template <bool FULL>
__device__ __forceinline__ void stage_rows(const float* __restrict__ src0, size_t row_stride, float* dstc, int r0,
                                           int nrows_valid, int ncols, int lane) {
  float t[24];
#pragma unroll
  for (int rr = 0; rr < 8; ++rr) {
    const float* src = src0 + (size_t)(r0 + rr) * row_stride;
    if (FULL) {
      t[rr * 3 + 0] = ld_stream(src);
      t[rr * 3 + 1] = ld_stream(src + 32);
      t[rr * 3 + 2] = ld_stream(src + 64);
    } else {
      const bool rowok = r0 + rr < nrows_valid;
      t[rr * 3 + 0] = (rowok && lane < ncols) ? ld_stream(src) : 0.f;
      t[rr * 3 + 1] = (rowok && lane + 32 < ncols) ? ld_stream(src + 32) : 0.f;
      t[rr * 3 + 2] = (rowok && lane + 64 < ncols) ? ld_stream(src + 64) : 0.f;
    }
  }
#pragma unroll
  for (int rr = 0; rr < 8; ++rr) {
    dstc[r0 + rr] = t[rr * 3 + 0];
    dstc[32 * TPITCH + r0 + rr] = t[rr * 3 + 1];
    dstc[64 * TPITCH + r0 + rr] = t[rr * 3 + 2];
  }
}

__global__ void __launch_bounds__(BWD_THREADS, 1)
lbs_bwd_kernel(const float* __restrict__ vpB, int G, const float* __restrict__ A_blk, int b0, int nb,
               const float* __restrict__ grad_verts, int V, int ntiles, int tiles_per_split,
               const uint32_t* __restrict__ vmeta, const float4* __restrict__ vwts,
               __nv_bfloat16* __restrict__ dvp_hi, __nv_bfloat16* __restrict__ dvp_lo, int n_pad,
               float* __restrict__ dA_acc, float* __restrict__ dtr_acc) {
  extern __shared__ __align__(128) float smem[];
  float* A_s = smem;                                         // [288][32]
  float* dA_s = A_s + AG_WORDS;                              // [288][32]
  float* dtr_s = dA_s + AG_WORDS;                            // [3][32]
  float* tiles = dtr_s + 96;                                 // [BWD_WARPS][96][33]: dV in / packed dvp out, in place
  uint64_t* bar = reinterpret_cast<uint64_t*>(tiles + BWD_WARPS * TTILE_WORDS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.x;
  const int col0 = g * 32;
  const int gb0 = b0 + col0;
  const int tile_begin = blockIdx.y * tiles_per_split;
  const int tile_end = min(ntiles, tile_begin + tiles_per_split);
  if (threadIdx.x == 0) fetch_group_transforms(A_s, A_blk, g, bar);

  float P[BPF * 3], Q[BPF * 3];
  int tile = tile_begin + warp;
  const float* chunk0 = vpB + (size_t)g * CHUNK_WORDS + lane;
  const size_t tstride = (size_t)G * CHUNK_WORDS;
  for (int r = threadIdx.x; r < AG_WORDS; r += BWD_THREADS) dA_s[r] = 0.f;
  if (threadIdx.x < 96) dtr_s[threadIdx.x] = 0.f;
  __syncthreads();
  mbar_wait(bar, 0);

  float* g_s = tiles + warp * TTILE_WORDS;
  float* g_lane = g_s + lane;
  const int nrows_valid = min(32, nb - col0);
  BwdState st;
#pragma unroll
  for (int e = 0; e < AELEMS; ++e) st.d0[e] = st.d1[e] = st.d2[e] = st.d3[e] = 0.f;
  st.j0 = st.j1 = st.j2 = st.j3 = 0;
  st.sx = st.sy = st.sz = 0.f;

  for (; tile < tile_end; tile += BWD_WARPS) {
    const int vbase = tile * TILE_V;
    const int ncols = min(TILE_V, V - vbase) * 3;
    const uint32_t* meta = vmeta + vbase;
    const float4* wts = vwts + vbase;
    const float* chunk = chunk0 + tile * tstride;
    fetch_vp4(P, chunk, meta);
    // ---- stage dV rows of this tile (coalesced 384 B per body row), 8 rows = 24 loads in flight ----
    {
      const float* src0 = grad_verts + ((size_t)gb0 * V + vbase) * 3 + lane;
      float* dstc = g_s + lane * TPITCH;
      if (nrows_valid == 32 && ncols == TILE_V * 3) {
#pragma unroll 1
        for (int r0 = 0; r0 < 32; r0 += 8) stage_rows<true>(src0, (size_t)V * 3, dstc, r0, 32, 96, lane);
      } else {
#pragma unroll 1
        for (int r0 = 0; r0 < 32; r0 += 8) stage_rows<false>(src0, (size_t)V * 3, dstc, r0, nrows_valid, ncols, lane);
      }
    }
    __syncwarp();
    // ---- arithmetic in place on the tile: 8 units of 4 vertices, v_posed prefetched one unit ahead ----
    fetch_vp4(Q, chunk, meta + BPF);
    skin_bwd(P, st, A_s, dA_s, lane, meta, wts, true, g_lane);
#pragma unroll 1
    for (int u = 1; u < TILE_V / BPF - 1; u += 2) {
      fetch_vp4(P, chunk, meta + (u + 1) * BPF);
      skin_bwd(Q, st, A_s, dA_s, lane, meta + u * BPF, wts + u * BPF, false, g_lane);
      fetch_vp4(Q, chunk, meta + (u + 2) * BPF);
      skin_bwd(P, st, A_s, dA_s, lane, meta + (u + 1) * BPF, wts + (u + 1) * BPF, false, g_lane);
    }
    skin_bwd(Q, st, A_s, dA_s, lane, meta + (TILE_V - BPF), wts + (TILE_V - BPF), false, g_lane);
    // every tile starts with all four slots reloaded: close this tile's accumulators now
    flush_slot(st.d0, dA_s, st.j0, lane);
    flush_slot(st.d1, dA_s, st.j1, lane);
    flush_slot(st.d2, dA_s, st.j2, lane);
    flush_slot(st.d3, dA_s, st.j3, lane);
    __syncwarp();
    flush_dvp_tile(reinterpret_cast<const uint32_t*>(g_s), dvp_hi, dvp_lo, (size_t)col0, n_pad, (size_t)vbase * 3, lane);
    __syncwarp();
  }
  atomicAdd(&dtr_s[lane], st.sx);
  atomicAdd(&dtr_s[32 + lane], st.sy);
  atomicAdd(&dtr_s[64 + lane], st.sz);
  __syncthreads();
  // accumulate into the slab-wide buffers [group][288][32] / [group][3][32]
  accumulate_rows(dA_acc + (size_t)g * AG_WORDS, dA_s, NJ * AELEMS, warp, BWD_WARPS, lane);
  if (warp < 3) atomicAdd(dtr_acc + (size_t)g * 96 + warp * 32 + lane, dtr_s[warp * 32 + lane]);
}

// ---------------------------------------------------------------------------------------------
// (nsplit, tiles per split): every split is a multiple of `warps` tiles; among 2..8 tiles per warp pick
// the split that wastes the least of the last wave (1 CTA / SM) and of the padded tile range
static void split_plan(int ntiles, int groups, int num_sms, int warps, int& nsplit, int& tps) {
  double best = -1.0;
  nsplit = 1;
  tps = (ntiles + warps - 1) / warps * warps;
  for (int tpw = 8; tpw >= 1; --tpw) {
    const int t = tpw * warps;
    const int ns = (ntiles + t - 1) / t;
    const long long ctas = (long long)groups * ns;
    const long long waves = (ctas + num_sms - 1) / num_sms;
    double eff = (double)ctas / (double)(waves * num_sms) * (double)ntiles / (double)(ns * t);
    if (tpw == 1) eff *= 0.9;                    // the per-CTA transform load is poorly amortised
    if (eff > best + 1e-9) {
      best = eff;
      nsplit = ns;
      tps = t;
    }
  }
}

int launch_lbs_fwd(const DevModel& m, const float* vpB, int S, const float* A_blk, int b0, int nb,
                   const float* transl, float* verts, int num_sms, cudaStream_t st) {
  if (nb <= 0) return 0;
  const int groups = (nb + 31) / 32;
  int nsplit, tps;
  split_plan(m.ntiles, groups, num_sms, FWD_WARPS, nsplit, tps);
  B200_CUDA_TRY(cudaFuncSetAttribute(lbs_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM));
  LaunchTimer _timer("lbs_fwd", st);
  lbs_fwd_kernel<<<dim3(groups, nsplit), FWD_THREADS, FWD_SMEM, st>>>(vpB, S / 32, A_blk, b0, nb, transl, verts, m.V,
                                                                       m.ntiles, tps, m.vmeta, m.vwts);
  B200_LAUNCH_CHECK("lbs_fwd");
  return 0;
}

// Sw = active slab width (multiple of 32, >= nb): rows of absent bodies get zero dvp / partials
// dA_acc [S/32][288][32] and dtr_acc [S/32][3][32] must be zeroed by the caller; CTAs add into them
int launch_lbs_bwd(const DevModel& m, const float* vpB, int S, int Sw, const float* A_blk, int b0, int nb,
                   const float* grad_verts, __nv_bfloat16* dvp_hi, __nv_bfloat16* dvp_lo, float* dA_part,
                   float* dtr_part, int num_sms, cudaStream_t st) {
  int nsplit, tps;
  split_plan(m.ntiles, Sw / 32, num_sms, BWD_WARPS, nsplit, tps);
  B200_CUDA_TRY(cudaFuncSetAttribute(lbs_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM));
  LaunchTimer _timer("lbs_bwd", st);
  lbs_bwd_kernel<<<dim3(Sw / 32, nsplit), BWD_THREADS, BWD_SMEM, st>>>(vpB, S / 32, A_blk, b0, nb, grad_verts, m.V,
                                                                        m.ntiles, tps, m.vmeta, m.vwts, dvp_hi,
                                                                        dvp_lo, m.n_pad, dA_part, dtr_part);
  B200_LAUNCH_CHECK("lbs_bwd");
  return 0;
}

}  // namespace b200smpl
