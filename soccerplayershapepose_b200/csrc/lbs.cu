// Linear blend skinning, forward and backward (smplx.lbs.lbs tail:
// T = W.A ; v = T.[v_posed;1]).  Memory-bound kernels.
//
// Mapping: lane = body, a warp owns 32 bodies and walks 32-vertex tiles.  Everything that is
// per-vertex (4 joint ids, 4 weights) is warp-uniform; everything per-body (the 4 cached joint
// transforms, 48 floats) lives in registers, re-loaded from the CTA's shared-memory copy of
// A[32 bodies][24][12] only when the packed plan says a slot's joint changes.  v_posed is read
// from the GEMM's body-fastest layout vpT[row][body] (one coalesced 128 B line per warp load),
// results are transposed through a per-warp shared-memory tile so that the (B, 6890, 3) output
// is written as contiguous 384 B row segments, once.
#include "common.cuh"

namespace b200smpl {

constexpr int LBS_WARPS = 8;
constexpr int LBS_THREADS = LBS_WARPS * 32;
constexpr int STAGE_PITCH = TILE_V * 3 + 1;  // 97 words: lane b -> bank (b + c) % 32, conflict-free

__device__ __forceinline__ void load_slot(float (&a)[AELEMS], const float* A_s, int joint, int lane) {
#pragma unroll
  for (int e = 0; e < AELEMS; ++e) a[e] = A_s[(joint * AELEMS + e) * 32 + lane];
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LBS_THREADS, 1)
lbs_fwd_kernel(const float* __restrict__ vpT, int S, const float* __restrict__ A_T, int b0, int nb,
               const float* __restrict__ transl, float* __restrict__ verts, int V, int ntiles,
               int tiles_per_split, const uint32_t* __restrict__ vmeta, const float4* __restrict__ vwts) {
  extern __shared__ float smem[];
  float* A_s = smem;                                   // [288][32]
  float* stage_all = smem + NJ * AELEMS * 32;          // [LBS_WARPS][32][STAGE_PITCH]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col0 = blockIdx.x * 32;                    // first slab column of this CTA
  const int col = col0 + lane;                         // slab column of this lane's body
  const int gb0 = b0 + col0;                           // global index of the CTA's first body
  const int b = gb0 + lane;

  for (int r = warp; r < NJ * AELEMS; r += LBS_WARPS) A_s[r * 32 + lane] = A_T[(size_t)r * S + col];
  float tx = 0.f, ty = 0.f, tz = 0.f;
  if (transl != nullptr && col < nb) {
    tx = transl[b * 3 + 0];
    ty = transl[b * 3 + 1];
    tz = transl[b * 3 + 2];
  }
  __syncthreads();

  float* stage = stage_all + warp * 32 * STAGE_PITCH;
  const int tile_begin = blockIdx.y * tiles_per_split;
  const int tile_end = min(ntiles, tile_begin + tiles_per_split);
  const int nrows_valid = min(32, nb - col0);          // bodies of this CTA that exist

  float a0[AELEMS], a1[AELEMS], a2[AELEMS], a3[AELEMS];
  for (int tile = tile_begin + warp; tile < tile_end; tile += LBS_WARPS) {
    const int vbase = tile * TILE_V;
    uint32_t force = 0xFu << 20;                       // a warp starts every tile with all slots loaded
#pragma unroll 4
    for (int i = 0; i < TILE_V; ++i) {
      uint32_t mt = __ldg(vmeta + vbase + i) | force;
      force = 0;
      if (!(mt & VMETA_VALID)) continue;
      const float4 w = __ldg(vwts + vbase + i);
      if (mt & (1u << 20)) load_slot(a0, A_s, mt & 31, lane);
      if (mt & (1u << 21)) load_slot(a1, A_s, (mt >> 5) & 31, lane);
      if (mt & (1u << 22)) load_slot(a2, A_s, (mt >> 10) & 31, lane);
      if (mt & (1u << 23)) load_slot(a3, A_s, (mt >> 15) & 31, lane);
      const int ol = (mt >> 24) & 31;
      const size_t row = (size_t)(vbase + ol) * 3;
      const float px = vpT[row * S + col];
      const float py = vpT[(row + 1) * S + col];
      const float pz = vpT[(row + 2) * S + col];
      float ox = tx, oy = ty, oz = tz;
#define B200_SKIN(a, wk)                                                        \
  ox = fmaf(wk, fmaf(a[0], px, fmaf(a[1], py, fmaf(a[2], pz, a[3]))), ox);      \
  oy = fmaf(wk, fmaf(a[4], px, fmaf(a[5], py, fmaf(a[6], pz, a[7]))), oy);      \
  oz = fmaf(wk, fmaf(a[8], px, fmaf(a[9], py, fmaf(a[10], pz, a[11]))), oz);
      B200_SKIN(a0, w.x)
      B200_SKIN(a1, w.y)
      B200_SKIN(a2, w.z)
      B200_SKIN(a3, w.w)
#undef B200_SKIN
      float* srow = stage + lane * STAGE_PITCH + ol * 3;
      srow[0] = ox;
      srow[1] = oy;
      srow[2] = oz;
    }
    __syncwarp();
    // flush: each body row of the tile is 96 contiguous floats of the (B, V, 3) output
    const int ncols = min(TILE_V, V - vbase) * 3;
    for (int r = 0; r < nrows_valid; ++r) {
      float* dst = verts + ((size_t)(gb0 + r) * V + vbase) * 3;
      const float* src = stage + r * STAGE_PITCH;
#pragma unroll
      for (int c = lane; c < TILE_V * 3; c += 32)
        if (c < ncols) dst[c] = src[c];
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
__device__ __forceinline__ void flush_slot(float (&d)[AELEMS], float* dA_s, int joint, int lane) {
#pragma unroll
  for (int e = 0; e < AELEMS; ++e) {
    atomicAdd(&dA_s[(joint * AELEMS + e) * 32 + lane], d[e]);
    d[e] = 0.f;
  }
}

__device__ __forceinline__ uint32_t pack_hi_lo(float x) {
  __nv_bfloat16 hi = __float2bfloat16_rn(x);
  __nv_bfloat16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
  return (uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(lo) << 16);
}

__global__ void __launch_bounds__(LBS_THREADS, 1)
lbs_bwd_kernel(const float* __restrict__ vpT, int S, const float* __restrict__ A_T, int b0, int nb,
               const float* __restrict__ grad_verts, int V, int ntiles, int tiles_per_split,
               const uint32_t* __restrict__ vmeta, const float4* __restrict__ vwts,
               __nv_bfloat16* __restrict__ dvp_hi, __nv_bfloat16* __restrict__ dvp_lo, int n_pad,
               float* __restrict__ dA_part, float* __restrict__ dtr_part) {
  extern __shared__ float smem[];
  float* A_s = smem;                                   // [288][32]
  float* dA_s = smem + NJ * AELEMS * 32;               // [288][32]
  float* dtr_s = dA_s + NJ * AELEMS * 32;              // [3][32]
  float* stage_all = dtr_s + 3 * 32;                   // [LBS_WARPS][32][STAGE_PITCH]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col0 = blockIdx.x * 32;
  const int col = col0 + lane;
  const int gb0 = b0 + col0;

  for (int r = warp; r < NJ * AELEMS; r += LBS_WARPS) {
    A_s[r * 32 + lane] = A_T[(size_t)r * S + col];
    dA_s[r * 32 + lane] = 0.f;
  }
  if (warp == 0) dtr_s[lane] = dtr_s[32 + lane] = dtr_s[64 + lane] = 0.f;
  __syncthreads();

  float* stage = stage_all + warp * 32 * STAGE_PITCH;
  uint32_t* stage_u = reinterpret_cast<uint32_t*>(stage);
  const int tile_begin = blockIdx.y * tiles_per_split;
  const int tile_end = min(ntiles, tile_begin + tiles_per_split);
  const int nrows_valid = min(32, nb - col0);

  float a0[AELEMS], a1[AELEMS], a2[AELEMS], a3[AELEMS];
  float d0[AELEMS], d1[AELEMS], d2[AELEMS], d3[AELEMS];
#pragma unroll
  for (int e = 0; e < AELEMS; ++e) d0[e] = d1[e] = d2[e] = d3[e] = 0.f;
  int j0 = 0, j1 = 0, j2 = 0, j3 = 0;
  bool have = false;
  float sx = 0.f, sy = 0.f, sz = 0.f;

  for (int tile = tile_begin + warp; tile < tile_end; tile += LBS_WARPS) {
    const int vbase = tile * TILE_V;
    const int ncols = min(TILE_V, V - vbase) * 3;
    // stage dV rows of this tile (coalesced 384 B per body row), zeros for absent bodies / vertices
    for (int r = 0; r < 32; ++r) {
      const float* src = grad_verts + ((size_t)(gb0 + r) * V + vbase) * 3;
      float* dst = stage + r * STAGE_PITCH;
#pragma unroll
      for (int c = lane; c < TILE_V * 3; c += 32) dst[c] = (r < nrows_valid && c < ncols) ? src[c] : 0.f;
    }
    __syncwarp();
    uint32_t force = 0xFu << 20;
#pragma unroll 2
    for (int i = 0; i < TILE_V; ++i) {
      uint32_t mt = __ldg(vmeta + vbase + i) | force;
      force = 0;
      const int ol = (mt >> 24) & 31;
      float* srow = stage + lane * STAGE_PITCH + ol * 3;
      if (!(mt & VMETA_VALID)) {                       // padded vertex: its dvp columns must be 0
        stage_u[lane * STAGE_PITCH + ol * 3 + 0] = 0u;
        stage_u[lane * STAGE_PITCH + ol * 3 + 1] = 0u;
        stage_u[lane * STAGE_PITCH + ol * 3 + 2] = 0u;
        continue;
      }
      const float4 w = __ldg(vwts + vbase + i);
      if (mt & (1u << 20)) { if (have) flush_slot(d0, dA_s, j0, lane); j0 = mt & 31; load_slot(a0, A_s, j0, lane); }
      if (mt & (1u << 21)) { if (have) flush_slot(d1, dA_s, j1, lane); j1 = (mt >> 5) & 31; load_slot(a1, A_s, j1, lane); }
      if (mt & (1u << 22)) { if (have) flush_slot(d2, dA_s, j2, lane); j2 = (mt >> 10) & 31; load_slot(a2, A_s, j2, lane); }
      if (mt & (1u << 23)) { if (have) flush_slot(d3, dA_s, j3, lane); j3 = (mt >> 15) & 31; load_slot(a3, A_s, j3, lane); }
      have = true;
      const size_t row = (size_t)(vbase + ol) * 3;
      const float px = vpT[row * S + col];
      const float py = vpT[(row + 1) * S + col];
      const float pz = vpT[(row + 2) * S + col];
      const float gx = srow[0], gy = srow[1], gz = srow[2];
      sx += gx; sy += gy; sz += gz;
      float qx = 0.f, qy = 0.f, qz = 0.f;
#define B200_SKIN_BWD(a, d, wk)                                                   \
  {                                                                               \
    const float hx = wk * gx, hy = wk * gy, hz = wk * gz;                         \
    qx = fmaf(a[0], hx, fmaf(a[4], hy, fmaf(a[8], hz, qx)));                      \
    qy = fmaf(a[1], hx, fmaf(a[5], hy, fmaf(a[9], hz, qy)));                      \
    qz = fmaf(a[2], hx, fmaf(a[6], hy, fmaf(a[10], hz, qz)));                     \
    d[0] = fmaf(hx, px, d[0]); d[1] = fmaf(hx, py, d[1]); d[2] = fmaf(hx, pz, d[2]); d[3] += hx;    \
    d[4] = fmaf(hy, px, d[4]); d[5] = fmaf(hy, py, d[5]); d[6] = fmaf(hy, pz, d[6]); d[7] += hy;    \
    d[8] = fmaf(hz, px, d[8]); d[9] = fmaf(hz, py, d[9]); d[10] = fmaf(hz, pz, d[10]); d[11] += hz; \
  }
      B200_SKIN_BWD(a0, d0, w.x)
      B200_SKIN_BWD(a1, d1, w.y)
      B200_SKIN_BWD(a2, d2, w.z)
      B200_SKIN_BWD(a3, d3, w.w)
#undef B200_SKIN_BWD
      stage_u[lane * STAGE_PITCH + ol * 3 + 0] = pack_hi_lo(qx);
      stage_u[lane * STAGE_PITCH + ol * 3 + 1] = pack_hi_lo(qy);
      stage_u[lane * STAGE_PITCH + ol * 3 + 2] = pack_hi_lo(qz);
    }
    __syncwarp();
    // flush dv_posed rows: 96 bf16 per body row per array; lane handles 2 consecutive columns (u32 store)
    for (int r = 0; r < 32; ++r) {
      const size_t o = (size_t)(col0 + r) * n_pad + (size_t)vbase * 3;
      const uint32_t* src = stage_u + r * STAGE_PITCH;
#pragma unroll
      for (int c = lane * 2; c < TILE_V * 3; c += 64) {
        const uint32_t u0 = src[c], u1 = src[c + 1];
        *reinterpret_cast<uint32_t*>(dvp_hi + o + c) = (u0 & 0xFFFFu) | (u1 << 16);
        if (dvp_lo != nullptr) *reinterpret_cast<uint32_t*>(dvp_lo + o + c) = (u0 >> 16) | (u1 & 0xFFFF0000u);
      }
    }
    __syncwarp();
  }
  if (have) {
    flush_slot(d0, dA_s, j0, lane);
    flush_slot(d1, dA_s, j1, lane);
    flush_slot(d2, dA_s, j2, lane);
    flush_slot(d3, dA_s, j3, lane);
  }
  atomicAdd(&dtr_s[lane], sx);
  atomicAdd(&dtr_s[32 + lane], sy);
  atomicAdd(&dtr_s[64 + lane], sz);
  __syncthreads();
  float* dA_out = dA_part + (size_t)blockIdx.y * NJ * AELEMS * S;
  for (int r = warp; r < NJ * AELEMS; r += LBS_WARPS) dA_out[(size_t)r * S + col] = dA_s[r * 32 + lane];
  if (warp < 3) dtr_part[((size_t)blockIdx.y * 3 + warp) * S + col] = dtr_s[warp * 32 + lane];
}

// ---------------------------------------------------------------------------------------------
// (nsplit, tiles per split): enough CTAs for ~2 per SM, every split a multiple of LBS_WARPS tiles
static void split_plan(int ntiles, int groups, int num_sms, int& nsplit, int& tps) {
  int want = (2 * num_sms + groups - 1) / groups;
  const int max_splits = (ntiles + LBS_WARPS - 1) / LBS_WARPS;
  want = max(1, min(want, max_splits));
  tps = ((ntiles + want - 1) / want + LBS_WARPS - 1) / LBS_WARPS * LBS_WARPS;
  nsplit = (ntiles + tps - 1) / tps;
}

int launch_lbs_fwd(const DevModel& m, const float* vpT, int S, const float* A_T, int b0, int nb,
                   const float* transl, float* verts, int num_sms, cudaStream_t st) {
  if (nb <= 0) return 0;
  const int groups = (nb + 31) / 32;
  int nsplit, tps;
  split_plan(m.ntiles, groups, num_sms, nsplit, tps);
  const size_t smem = (size_t)(NJ * AELEMS * 32 + LBS_WARPS * 32 * STAGE_PITCH) * sizeof(float);
  B200_CUDA_TRY(cudaFuncSetAttribute(lbs_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  LaunchTimer _timer_261("lbs_fwd", st);
  lbs_fwd_kernel<<<dim3(groups, nsplit), LBS_THREADS, smem, st>>>(vpT, S, A_T, b0, nb, transl, verts, m.V, m.ntiles,
                                                                  tps, m.vmeta, m.vwts);
  B200_LAUNCH_CHECK("lbs_fwd");
  return 0;
}

// number of dA / dtransl partials the backward skinning kernel writes per body for slab pitch S
int lbs_bwd_splits(const DevModel& m, int S, int num_sms) {
  int nsplit, tps;
  split_plan(m.ntiles, (S + 31) / 32, num_sms, nsplit, tps);
  return nsplit;
}

// Sw = active slab width (multiple of 32, >= nb): rows of absent bodies get zero dvp / partials
int launch_lbs_bwd(const DevModel& m, const float* vpT, int S, int Sw, const float* A_T, int b0, int nb,
                   const float* grad_verts, __nv_bfloat16* dvp_hi, __nv_bfloat16* dvp_lo, float* dA_part,
                   float* dtr_part, int nsplit, cudaStream_t st) {
  const int tps = ((m.ntiles + nsplit - 1) / nsplit + LBS_WARPS - 1) / LBS_WARPS * LBS_WARPS;
  const size_t smem = (size_t)(2 * NJ * AELEMS * 32 + 3 * 32 + LBS_WARPS * 32 * STAGE_PITCH) * sizeof(float);
  B200_CUDA_TRY(cudaFuncSetAttribute(lbs_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  LaunchTimer _timer_281("lbs_bwd", st);
  lbs_bwd_kernel<<<dim3(Sw / 32, nsplit), LBS_THREADS, smem, st>>>(vpT, S, A_T, b0, nb, grad_verts, m.V, m.ntiles,
                                                                   tps, m.vmeta, m.vwts, dvp_hi, dvp_lo, m.n_pad,
                                                                   dA_part, dtr_part);
  B200_LAUNCH_CHECK("lbs_bwd");
  return 0;
}

}  // namespace b200smpl
