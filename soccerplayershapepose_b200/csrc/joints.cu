// Joint outputs: the 90-joint superset of
// models/smpl_official.py:27-41 -- 24 posed chain joints, 21 joints picked from vertices
// (smplx VertexJointSelector), 45 joints regressed from vertices (J_regressor_extra / cocoplus /
// h36m) -- plus transl and the optional weak-perspective reprojection (utils/cam_utils.py:5-26).
//
// No vertex is re-read: a regressed joint  sum_v Jr[j][v] * skin(v)  is rewritten at pack time as
//   sum_i  A_i . [ q_ji ; c_ji ],   q_ji = sum_v Jr[j][v] w_vi p_v  (linear in the blend features),
// so the q_ji are extra "virtual" rows of the blend GEMM and a joint costs a handful of 3x4
// transforms here.  Picked joints are virtual copies of their vertex's rows.  lane = body.
#include "common.cuh"

namespace b200smpl {

constexpr int JT_WARPS = 8;
constexpr int JT_THREADS = JT_WARPS * 32;

__global__ void __launch_bounds__(JT_THREADS)
joints_fwd_kernel(DevModel m, const float* __restrict__ vpT, int S, const float* __restrict__ A_T,
                  const float* __restrict__ jposed_T, int b0, int nb, const float* __restrict__ transl,
                  const float* __restrict__ cam, float* __restrict__ joints, float* __restrict__ joints2d) {
  extern __shared__ float smem[];
  const int ncol = m.njout * 3;
  const int pitch = ncol | 1;                          // odd pitch -> conflict-free column access by lane
  float* sJ = smem;                                    // [32][pitch]
  float* sCam = smem + 32 * pitch;                     // [32][3]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col0 = blockIdx.x * 32;
  const int col = col0 + lane;
  const int gb0 = b0 + col0;
  const int b = gb0 + lane;
  const bool live = col < nb;
  float tx = 0.f, ty = 0.f, tz = 0.f;
  if (transl != nullptr && live) {
    tx = transl[b * 3 + 0];
    ty = transl[b * 3 + 1];
    tz = transl[b * 3 + 2];
  }
  if (warp == 0 && cam != nullptr) {
    sCam[lane * 3 + 0] = live ? cam[b * 3 + 0] : 0.f;
    sCam[lane * 3 + 1] = live ? cam[b * 3 + 1] : 0.f;
    sCam[lane * 3 + 2] = live ? cam[b * 3 + 2] : 0.f;
  }
  // chain joints
  for (int r = warp; r < NJ * 3; r += JT_WARPS) {
    const float t = (r % 3 == 0) ? tx : ((r % 3 == 1) ? ty : tz);
    sJ[lane * pitch + r] = jposed_T[(size_t)r * S + col] + t;
  }
  // picked + regressed joints
  for (int J = NJ + warp; J < m.njout; J += JT_WARPS) {
    float x = tx, y = ty, z = tz;
    const int t0 = m.term_ptr[J - NJ], t1 = m.term_ptr[J - NJ + 1];
    for (int t = t0; t < t1; ++t) {
      const int i = m.term_joint[t];
      const size_t q = (size_t)m.term_qrow[t];
      const float c = m.term_c[t];
      const float qx = vpT[q * S + col], qy = vpT[(q + 1) * S + col], qz = vpT[(q + 2) * S + col];
      const float* a = A_T + (size_t)(i * AELEMS) * S + col;
      x += a[0] * qx + a[(size_t)1 * S] * qy + a[(size_t)2 * S] * qz + a[(size_t)3 * S] * c;
      y += a[(size_t)4 * S] * qx + a[(size_t)5 * S] * qy + a[(size_t)6 * S] * qz + a[(size_t)7 * S] * c;
      z += a[(size_t)8 * S] * qx + a[(size_t)9 * S] * qy + a[(size_t)10 * S] * qz + a[(size_t)11 * S] * c;
    }
    sJ[lane * pitch + J * 3 + 0] = x;
    sJ[lane * pitch + J * 3 + 1] = y;
    sJ[lane * pitch + J * 3 + 2] = z;
  }
  __syncthreads();
  const int nrows = min(32, nb - col0);
  for (int idx = threadIdx.x; idx < nrows * ncol; idx += JT_THREADS) {
    const int r = idx / ncol, c = idx - r * ncol;
    joints[(size_t)gb0 * ncol + idx] = sJ[r * pitch + c];
  }
  if (joints2d != nullptr && cam != nullptr) {
    const int n2 = m.njout * 2;
    for (int idx = threadIdx.x; idx < nrows * n2; idx += JT_THREADS) {
      const int r = idx / n2, c = idx - r * n2;
      const int J = c >> 1, k = c & 1;
      const float s = sCam[r * 3 + 0], t = sCam[r * 3 + 1 + k];
      joints2d[(size_t)gb0 * n2 + idx] = s * (sJ[r * pitch + J * 3 + k] + t);
    }
  }
}

__device__ __forceinline__ void store_hi_lo(__nv_bfloat16* hi, __nv_bfloat16* lo, size_t idx, float x) {
  const __nv_bfloat16 h = __float2bfloat16_rn(x);
  hi[idx] = h;
  if (lo != nullptr) lo[idx] = __float2bfloat16_rn(x - __bfloat162float(h));
}

// backward of the above.  dJ_total = grad_joints + [s*g2d_x, s*g2d_y, 0]
//   chain joints (J < 24): dJposed_T[r][b] (consumed by pose_bwd)
//   other joints: dq -> virtual columns of dvp ; dA -> dA_part (this kernel's own partial) ;
//   dtransl partial = sum_J dJ_total ; dcam from the 2D gradient.
__global__ void __launch_bounds__(JT_THREADS)
joints_bwd_kernel(DevModel m, const float* __restrict__ vpT, int S, const float* __restrict__ A_T, int b0,
                  int nb, const float* __restrict__ cam, const float* __restrict__ joints,
                  const float* __restrict__ grad_joints, const float* __restrict__ grad_joints2d,
                  __nv_bfloat16* __restrict__ dvp_hi, __nv_bfloat16* __restrict__ dvp_lo,
                  float* __restrict__ dA_part, float* __restrict__ dtr_part, float* __restrict__ dJposed_T,
                  float* __restrict__ grad_cam) {
  extern __shared__ float smem[];
  const int ncol = m.njout * 3;
  const int pitch = ncol | 1;
  float* sG = smem;                                    // [32][pitch]   total joint gradient
  float* sdA = sG + 32 * pitch;                        // [288][32]
  float* sCam = sdA + NJ * AELEMS * 32;                // [32][3]
  float* sdCam = sCam + 96;                            // [32][3]
  float* sdT = sdCam + 96;                             // [3][32]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col0 = blockIdx.x * 32;
  const int col = col0 + lane;
  const int gb0 = b0 + col0;
  const int b = gb0 + lane;
  const int nrows = min(32, nb - col0);
  const bool has2d = grad_joints2d != nullptr && cam != nullptr;

  for (int r = warp; r < NJ * AELEMS; r += JT_WARPS) sdA[r * 32 + lane] = 0.f;
  if (warp == 0) {
    const bool live = col < nb;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      sCam[lane * 3 + k] = (cam != nullptr && live) ? cam[b * 3 + k] : 0.f;
      sdCam[lane * 3 + k] = 0.f;
      sdT[k * 32 + lane] = 0.f;
    }
  }
  __syncthreads();
  // stage total joint gradient, coalesced over the [32 bodies][njout*3] block
  for (int idx = threadIdx.x; idx < 32 * ncol; idx += JT_THREADS) {
    const int r = idx / ncol, c = idx - r * ncol;
    float g = 0.f;
    if (r < nrows) {
      if (grad_joints != nullptr) g = grad_joints[(size_t)gb0 * ncol + idx];
      if (has2d) {
        const int J = c / 3, k = c - J * 3;
        if (k < 2) {
          const float g2 = grad_joints2d[((size_t)(gb0 + r) * m.njout + J) * 2 + k];
          const float s = sCam[r * 3 + 0];
          g += s * g2;
          // d/ds : g2 * (x + t_k) ; d/dt_k : s * g2
          const float xk = joints[(size_t)gb0 * ncol + idx];
          atomicAdd(&sdCam[r * 3 + 0], g2 * (xk + sCam[r * 3 + 1 + k]));
          atomicAdd(&sdCam[r * 3 + 1 + k], s * g2);
        }
      }
    }
    sG[r * pitch + c] = g;
  }
  __syncthreads();
  // chain joints: transpose out
  for (int r = warp; r < NJ * 3; r += JT_WARPS) dJposed_T[(size_t)r * S + col] = sG[lane * pitch + r];
  // translation partial: sum over all joints
  {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    for (int J = warp; J < m.njout; J += JT_WARPS) {
      s0 += sG[lane * pitch + J * 3 + 0];
      s1 += sG[lane * pitch + J * 3 + 1];
      s2 += sG[lane * pitch + J * 3 + 2];
    }
    atomicAdd(&sdT[lane], s0);
    atomicAdd(&sdT[32 + lane], s1);
    atomicAdd(&sdT[64 + lane], s2);
  }
  // picked / regressed joints
  for (int J = NJ + warp; J < m.njout; J += JT_WARPS) {
    const float gx = sG[lane * pitch + J * 3 + 0], gy = sG[lane * pitch + J * 3 + 1],
                gz = sG[lane * pitch + J * 3 + 2];
    const int t0 = m.term_ptr[J - NJ], t1 = m.term_ptr[J - NJ + 1];
    int t = t0;
    while (t < t1) {
      const int qrow = m.term_qrow[t];
      const size_t q = (size_t)qrow;
      const float qx = vpT[q * S + col], qy = vpT[(q + 1) * S + col], qz = vpT[(q + 2) * S + col];
      float dx = 0.f, dy = 0.f, dz = 0.f;
      // all terms that share this q-group are consecutive
      for (; t < t1 && m.term_qrow[t] == qrow; ++t) {
        const int i = m.term_joint[t];
        const float c = m.term_c[t];
        const float* a = A_T + (size_t)(i * AELEMS) * S + col;
        dx += a[0] * gx + a[(size_t)4 * S] * gy + a[(size_t)8 * S] * gz;
        dy += a[(size_t)1 * S] * gx + a[(size_t)5 * S] * gy + a[(size_t)9 * S] * gz;
        dz += a[(size_t)2 * S] * gx + a[(size_t)6 * S] * gy + a[(size_t)10 * S] * gz;
        float* d = sdA + (i * AELEMS) * 32 + lane;
        atomicAdd(d + 0 * 32, gx * qx); atomicAdd(d + 1 * 32, gx * qy); atomicAdd(d + 2 * 32, gx * qz);
        atomicAdd(d + 3 * 32, gx * c);
        atomicAdd(d + 4 * 32, gy * qx); atomicAdd(d + 5 * 32, gy * qy); atomicAdd(d + 6 * 32, gy * qz);
        atomicAdd(d + 7 * 32, gy * c);
        atomicAdd(d + 8 * 32, gz * qx); atomicAdd(d + 9 * 32, gz * qy); atomicAdd(d + 10 * 32, gz * qz);
        atomicAdd(d + 11 * 32, gz * c);
      }
      const size_t o = (size_t)col * m.n_pad + q;
      store_hi_lo(dvp_hi, dvp_lo, o + 0, dx);
      store_hi_lo(dvp_hi, dvp_lo, o + 1, dy);
      store_hi_lo(dvp_hi, dvp_lo, o + 2, dz);
    }
  }
  __syncthreads();
  for (int r = warp; r < NJ * AELEMS; r += JT_WARPS) dA_part[(size_t)r * S + col] = sdA[r * 32 + lane];
  if (warp < 3) dtr_part[(size_t)warp * S + col] = sdT[warp * 32 + lane];
  if (grad_cam != nullptr && warp == 0 && col < nb) {
#pragma unroll
    for (int k = 0; k < 3; ++k) grad_cam[b * 3 + k] = sdCam[lane * 3 + k];
  }
}

int launch_joints_fwd(const DevModel& m, const float* vpT, int S, const float* A_T, const float* jposed_T, int b0,
                      int nb, const float* transl, const float* cam, float* joints, float* joints2d,
                      cudaStream_t st) {
  if (nb <= 0) return 0;
  const int pitch = (m.njout * 3) | 1;
  const size_t smem = (size_t)(32 * pitch + 96) * sizeof(float);
  B200_CUDA_TRY(cudaFuncSetAttribute(joints_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  LaunchTimer _timer_211("joints_fwd", st);
  joints_fwd_kernel<<<(nb + 31) / 32, JT_THREADS, smem, st>>>(m, vpT, S, A_T, jposed_T, b0, nb, transl, cam, joints,
                                                              joints2d);
  B200_LAUNCH_CHECK("joints_fwd");
  return 0;
}

// Sw: active slab width (multiple of 32); absent bodies get zero dvp columns / partials.
int launch_joints_bwd(const DevModel& m, const float* vpT, int S, int Sw, const float* A_T, int b0, int nb,
                      const float* cam, const float* joints, const float* grad_joints,
                      const float* grad_joints2d, __nv_bfloat16* dvp_hi, __nv_bfloat16* dvp_lo, float* dA_part,
                      float* dtr_part, float* dJposed_T, float* grad_cam, cudaStream_t st) {
  const int pitch = (m.njout * 3) | 1;
  const size_t smem = (size_t)(32 * pitch + NJ * AELEMS * 32 + 96 * 3) * sizeof(float);
  B200_CUDA_TRY(cudaFuncSetAttribute(joints_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  LaunchTimer _timer_225("joints_bwd", st);
  joints_bwd_kernel<<<Sw / 32, JT_THREADS, smem, st>>>(m, vpT, S, A_T, b0, nb, cam, joints, grad_joints,
                                                       grad_joints2d, dvp_hi, dvp_lo, dA_part, dtr_part, dJposed_T,
                                                       grad_cam);
  B200_LAUNCH_CHECK("joints_bwd");
  return 0;
}

}  // namespace b200smpl
