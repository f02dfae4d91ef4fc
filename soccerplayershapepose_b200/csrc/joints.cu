// Joint outputs: the 66 non-chain joints of the 90-joint
// superset of models/smpl_official.py:27-41 -- 21 joints picked from vertices (smplx
// VertexJointSelector) and 45 joints regressed from vertices (J_regressor_extra / cocoplus / h36m).
//
// No vertex is re-read: a regressed joint  sum_v Jr[j][v] * skin(v)  is rewritten at pack time as
//   sum_i  A_i . [ q_ji ; c_ji ],   q_ji = sum_v Jr[j][v] w_vi p_v  (linear in the blend features),
// so the q_ji are extra "virtual" rows of the blend GEMM, grouped 32 q-groups (96 rows) per virtual
// tile, and a joint costs a handful of 3x4 transforms.  These kernels are the skinning kernels of
// lbs.cu specialised to virtual tiles: lane = body, group transforms in shared memory via one TMA
// bulk copy, q rows prefetched from the group-blocked blend output, results transposed through a
// [96][33] shared tile and written as contiguous row segments of joints (B, 90, 3) / dvp.
// The 24 chain joints are written by the pose kernel; reprojection is the orthographic kernel.
#include "skin_common.cuh"

namespace b200smpl {

constexpr int JW = 6;                       // warps per CTA; one CTA per body group, warps stride over the virtual tiles
constexpr int JT = JW * 32;
constexpr size_t JFWD_SMEM = (size_t)(AG_WORDS + JW * TTILE_WORDS) * 4 + 16;
constexpr size_t JBWD_SMEM = (size_t)(2 * AG_WORDS + 96 + JW * 2 * TTILE_WORDS) * 4 + 16;

__global__ void __launch_bounds__(JT, 1)
joints_fwd_kernel(DevModel m, const float* __restrict__ vpB, int G, const float* __restrict__ A_blk, int b0, int nb,
                  const float* __restrict__ transl, float* __restrict__ joints) {
  extern __shared__ __align__(128) float smem[];
  float* A_s = smem;
  float* tiles = smem + AG_WORDS;
  uint64_t* bar = reinterpret_cast<uint64_t*>(tiles + JW * TTILE_WORDS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.x, col0 = g * 32, gb0 = b0 + col0;
  if (threadIdx.x == 0) fetch_group_transforms(A_s, A_blk, g, bar);
  float tx = 0.f, ty = 0.f, tz = 0.f;
  if (transl != nullptr && col0 + lane < nb) {
    const float* t = transl + (size_t)(gb0 + lane) * 3;
    tx = t[0]; ty = t[1]; tz = t[2];
  }
  __syncthreads();
  float* out_s = tiles + warp * TTILE_WORDS;
  float* out_lane = out_s + lane;
  const int nrows = min(32, nb - col0);
  const size_t ncol_all = (size_t)m.njout * 3;
  bool waited = false;
  for (int tv = warp; tv < m.ntv; tv += JW) {
    float q[96];
    const float* chunk = vpB + ((size_t)(m.ntiles + tv) * G + g) * CHUNK_WORDS + lane;
#pragma unroll
    for (int i = 0; i < 96; ++i) q[i] = ld_stream(chunk + i * 32);
    if (!waited) { mbar_wait(bar, 0); waited = true; }
    const uint32_t* meta = m.qmeta + tv * 32;
    const float* coef = m.qcoef + tv * 32;
    float a[AELEMS];
    float x = tx, y = ty, z = tz;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const uint32_t mt = __ldg(meta + i);
      if (!(mt & (1u << 14))) continue;
      if ((mt & (1u << 5)) || i == 0) load_slot(a, A_s, mt & 31, lane);
      const float c = __ldg(coef + i);
      const float qx = q[i * 3], qy = q[i * 3 + 1], qz = q[i * 3 + 2];
      x += fmaf(a[0], qx, fmaf(a[1], qy, fmaf(a[2], qz, a[3] * c)));
      y += fmaf(a[4], qx, fmaf(a[5], qy, fmaf(a[6], qz, a[7] * c)));
      z += fmaf(a[8], qx, fmaf(a[9], qy, fmaf(a[10], qz, a[11] * c)));
      if (mt & (1u << 13)) {                               // last term of this joint
        float* o = out_lane + ((mt >> 8) & 31) * (3 * TPITCH);
        o[0] = x; o[TPITCH] = y; o[2 * TPITCH] = z;
        x = tx; y = ty; z = tz;
      }
    }
    __syncwarp();
    // flush the tile's joints: 3 nj contiguous floats per body row, flattened so that loads/stores pipeline
    const int ncols = m.vt_nj[tv] * 3;
    float* dst0 = joints + (size_t)gb0 * ncol_all + (size_t)(NJ + m.vt_j0[tv]) * 3;
    for (int idx = lane; idx < nrows * ncols; idx += 32) {
      const int r = idx / ncols, c = idx - r * ncols;
      dst0[(size_t)r * ncol_all + c] = out_s[c * TPITCH + r];
    }
    __syncwarp();
  }
}

// backward over virtual tiles.  dJ: total joint gradient (B, NJout, 3).
//   dq -> virtual columns of dvp ; dA -> dA_part[blockIdx.y] ; dtransl partial = sum over the tile's joints
__global__ void __launch_bounds__(JT, 1)
joints_bwd_kernel(DevModel m, const float* __restrict__ vpB, int G, const float* __restrict__ A_blk, int b0, int nb,
                  const float* __restrict__ dJ, __nv_bfloat16* __restrict__ dvp_hi,
                  __nv_bfloat16* __restrict__ dvp_lo, float* __restrict__ dA_acc, float* __restrict__ dtr_acc) {
  extern __shared__ __align__(128) float smem[];
  float* A_s = smem;
  float* dA_s = A_s + AG_WORDS;
  float* dtr_s = dA_s + AG_WORDS;
  float* tiles = dtr_s + 96;                               // [JW][2][96][33]: joint gradients | packed dq
  uint64_t* bar = reinterpret_cast<uint64_t*>(tiles + JW * 2 * TTILE_WORDS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.x, col0 = g * 32, gb0 = b0 + col0;
  if (threadIdx.x == 0) fetch_group_transforms(A_s, A_blk, g, bar);
  for (int r = threadIdx.x; r < AG_WORDS; r += JT) dA_s[r] = 0.f;
  if (threadIdx.x < 96) dtr_s[threadIdx.x] = 0.f;
  __syncthreads();
  float* g_s = tiles + warp * 2 * TTILE_WORDS;
  uint32_t* q_s = reinterpret_cast<uint32_t*>(g_s + TTILE_WORDS);
  const int nrows = min(32, nb - col0);
  const size_t ncol_all = (size_t)m.njout * 3;
  float sx = 0.f, sy = 0.f, sz = 0.f;
  bool waited = false;
  for (int tv = warp; tv < m.ntv; tv += JW) {
    float q[96];
    const float* chunk = vpB + ((size_t)(m.ntiles + tv) * G + g) * CHUNK_WORDS + lane;
#pragma unroll
    for (int i = 0; i < 96; ++i) q[i] = ld_stream(chunk + i * 32);
    const int ncols = m.vt_nj[tv] * 3;
    // stage the gradients of this tile's joints (3 nj floats of each body row), flattened
    {
      const float* src0 = dJ + (size_t)gb0 * ncol_all + (size_t)(NJ + m.vt_j0[tv]) * 3;
      for (int idx = lane; idx < 32 * ncols; idx += 32) {
        const int r = idx / ncols, c = idx - r * ncols;
        g_s[c * TPITCH + r] = (r < nrows) ? ld_stream(src0 + (size_t)r * ncol_all + c) : 0.f;
      }
    }
    if (!waited) { mbar_wait(bar, 0); waited = true; }
    __syncwarp();
    const uint32_t* meta = m.qmeta + tv * 32;
    const float* coef = m.qcoef + tv * 32;
    float a[9], d[AELEMS];
#pragma unroll
    for (int e = 0; e < AELEMS; ++e) d[e] = 0.f;
    int jcur = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const uint32_t mt = __ldg(meta + i);
      uint32_t* qo = q_s + (i * 3) * TPITCH + lane;
      if (!(mt & (1u << 14))) {                            // dummy slot: its dvp columns must be 0
        qo[0] = 0u; qo[TPITCH] = 0u; qo[2 * TPITCH] = 0u;
        continue;
      }
      if ((mt & (1u << 5)) || i == 0) {
        if (i) flush_slot(d, dA_s, jcur, lane);
        jcur = mt & 31;
        load_rot(a, A_s, jcur, lane);
      }
      const float c = __ldg(coef + i);
      const float* gj = g_s + ((mt >> 8) & 31) * (3 * TPITCH) + lane;
      const float gx = gj[0], gy = gj[TPITCH], gz = gj[2 * TPITCH];
      const float qx = q[i * 3], qy = q[i * 3 + 1], qz = q[i * 3 + 2];
      qo[0] = pack_hi_lo(fmaf(a[0], gx, fmaf(a[3], gy, a[6] * gz)));
      qo[TPITCH] = pack_hi_lo(fmaf(a[1], gx, fmaf(a[4], gy, a[7] * gz)));
      qo[2 * TPITCH] = pack_hi_lo(fmaf(a[2], gx, fmaf(a[5], gy, a[8] * gz)));
      d[0] = fmaf(gx, qx, d[0]); d[1] = fmaf(gx, qy, d[1]); d[2] = fmaf(gx, qz, d[2]); d[3] = fmaf(gx, c, d[3]);
      d[4] = fmaf(gy, qx, d[4]); d[5] = fmaf(gy, qy, d[5]); d[6] = fmaf(gy, qz, d[6]); d[7] = fmaf(gy, c, d[7]);
      d[8] = fmaf(gz, qx, d[8]); d[9] = fmaf(gz, qy, d[9]); d[10] = fmaf(gz, qz, d[10]); d[11] = fmaf(gz, c, d[11]);
      if (mt & (1u << 13)) { sx += gx; sy += gy; sz += gz; }
    }
    flush_slot(d, dA_s, jcur, lane);
    __syncwarp();
    flush_dvp_tile(q_s, dvp_hi, dvp_lo, (size_t)col0, m.n_pad, (size_t)m.n_virt0 + (size_t)tv * 96, lane);
    __syncwarp();
  }
  atomicAdd(&dtr_s[lane], sx);
  atomicAdd(&dtr_s[32 + lane], sy);
  atomicAdd(&dtr_s[64 + lane], sz);
  __syncthreads();
  accumulate_rows(dA_acc + (size_t)g * AG_WORDS, dA_s, NJ * AELEMS, warp, JW, lane);
  if (warp < 3) atomicAdd(dtr_acc + (size_t)g * 96 + warp * 32 + lane, dtr_s[warp * 32 + lane]);
}

// total joint gradient when a 2D reprojection gradient is present:
//   dJ = grad_joints + [s g2d_x, s g2d_y, 0] ;  dcam = [sum g2d.(xy + t), s sum g2d_x, s sum g2d_y]
__global__ void __launch_bounds__(128)
joint_grad_total_kernel(const float* __restrict__ joints, const float* __restrict__ cam,
                        const float* __restrict__ gj, const float* __restrict__ g2d, float* __restrict__ dJ,
                        float* __restrict__ gcam, int nj) {
  __shared__ float sh[3][4];
  const int b = blockIdx.x;
  const float s = cam[b * 3], tx = cam[b * 3 + 1], ty = cam[b * 3 + 2];
  float gs = 0.f, gtx = 0.f, gty = 0.f;
  for (int j = threadIdx.x; j < nj; j += blockDim.x) {
    const size_t o3 = ((size_t)b * nj + j) * 3, o2 = ((size_t)b * nj + j) * 2;
    const float gu = g2d[o2], gv = g2d[o2 + 1];
    dJ[o3] = (gj ? gj[o3] : 0.f) + s * gu;
    dJ[o3 + 1] = (gj ? gj[o3 + 1] : 0.f) + s * gv;
    dJ[o3 + 2] = gj ? gj[o3 + 2] : 0.f;
    gs += gu * (joints[o3] + tx) + gv * (joints[o3 + 1] + ty);
    gtx += s * gu;
    gty += s * gv;
  }
  float v[3] = {gs, gtx, gty};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    if ((threadIdx.x & 31) == 0) sh[k][threadIdx.x >> 5] = v[k];
  }
  __syncthreads();
  if (gcam != nullptr && threadIdx.x < 3) gcam[b * 3 + threadIdx.x] = sh[threadIdx.x][0] + sh[threadIdx.x][1] + sh[threadIdx.x][2] + sh[threadIdx.x][3];
}

int launch_joints_fwd(const DevModel& m, const float* vpB, int S, const float* A_blk, int b0, int nb,
                      const float* transl, float* joints, cudaStream_t st) {
  if (nb <= 0 || m.ntv == 0) return 0;
  B200_CUDA_TRY(cudaFuncSetAttribute(joints_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)JFWD_SMEM));
  LaunchTimer _timer("joints_fwd", st);
  joints_fwd_kernel<<<(nb + 31) / 32, JT, JFWD_SMEM, st>>>(m, vpB, S / 32, A_blk, b0, nb,
                                                                                        transl, joints);
  B200_LAUNCH_CHECK("joints_fwd");
  return 0;
}

// Sw: active slab width (multiple of 32); absent bodies get zero dvp columns / partials.
int launch_joints_bwd(const DevModel& m, const float* vpB, int S, int Sw, const float* A_blk, int b0, int nb,
                      const float* dJ, __nv_bfloat16* dvp_hi, __nv_bfloat16* dvp_lo, float* dA_part, float* dtr_part,
                      cudaStream_t st) {
  if (m.ntv == 0) return 0;
  B200_CUDA_TRY(cudaFuncSetAttribute(joints_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)JBWD_SMEM));
  LaunchTimer _timer("joints_bwd", st);
  joints_bwd_kernel<<<Sw / 32, JT, JBWD_SMEM, st>>>(m, vpB, S / 32, A_blk, b0, nb, dJ,
                                                                                 dvp_hi, dvp_lo, dA_part, dtr_part);
  B200_LAUNCH_CHECK("joints_bwd");
  return 0;
}

int launch_joint_grad_total(const float* joints, const float* cam, const float* gj, const float* g2d, float* dJ,
                            float* gcam, int B, int nj, cudaStream_t st) {
  LaunchTimer _timer("joint_grad_total", st);
  joint_grad_total_kernel<<<B, 128, 0, st>>>(joints, cam, gj, g2d, dJ, gcam, nj);
  B200_LAUNCH_CHECK("joint_grad_total");
  return 0;
}

}  // namespace b200smpl
