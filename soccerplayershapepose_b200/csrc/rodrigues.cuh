// Rodrigues formula exactly as smplx.lbs.batch_rodrigues evaluates it (including its +1e-8 on the angle), forward
// and hand-written backward; shared by the pose kernels (pose.cu) and the standalone batch_rodrigues op (aux_ops.cu).
#pragma once

#include "common.cuh"

namespace b200smpl {

// smplx.lbs.batch_rodrigues for one joint: theta = ||r + 1e-8||, axis = r / theta.
__device__ __forceinline__ void rodrigues_fwd(const float r[3], float R[9]) {
  const float ex = r[0] + 1e-8f, ey = r[1] + 1e-8f, ez = r[2] + 1e-8f;
  const float theta = sqrtf(ex * ex + ey * ey + ez * ez);
  const float inv = 1.0f / theta;
  const float dx = r[0] * inv, dy = r[1] * inv, dz = r[2] * inv;
  float s, c;
  sincosf(theta, &s, &c);
  const float oc = 1.0f - c;
  // K = [[0,-dz,dy],[dz,0,-dx],[-dy,dx,0]] ; K^2 = d d^T - |d|^2 I
  const float n2 = dx * dx + dy * dy + dz * dz;
  R[0] = 1.0f + oc * (dx * dx - n2);
  R[1] = -s * dz + oc * (dx * dy);
  R[2] = s * dy + oc * (dx * dz);
  R[3] = s * dz + oc * (dx * dy);
  R[4] = 1.0f + oc * (dy * dy - n2);
  R[5] = -s * dx + oc * (dy * dz);
  R[6] = -s * dy + oc * (dx * dz);
  R[7] = s * dx + oc * (dy * dz);
  R[8] = 1.0f + oc * (dz * dz - n2);
}

// gradient of the above: G = dL/dR (row-major) -> dL/dr
__device__ __forceinline__ void rodrigues_bwd(const float r[3], const float G[9], float dr[3]) {
  const float e[3] = {r[0] + 1e-8f, r[1] + 1e-8f, r[2] + 1e-8f};
  const float theta = sqrtf(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
  const float inv = 1.0f / theta;
  const float d[3] = {r[0] * inv, r[1] * inv, r[2] * inv};
  float s, c;
  sincosf(theta, &s, &c);
  const float oc = 1.0f - c;
  const float K[9] = {0.f, -d[2], d[1], d[2], 0.f, -d[0], -d[1], d[0], 0.f};
  float K2[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) K2[i * 3 + j] = K[i * 3] * K[j] + K[i * 3 + 1] * K[3 + j] + K[i * 3 + 2] * K[6 + j];
  // dL/dtheta (direct) = cos <G,K> + sin <G,K^2>
  float gk = 0.f, gk2 = 0.f;
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    gk += G[i] * K[i];
    gk2 += G[i] * K2[i];
  }
  const float dtheta_direct = c * gk + s * gk2;
  // dL/dK = sin G + (1-cos) (G K^T + K^T G)
  float dK[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float gkt = 0.f, ktg = 0.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        gkt += G[i * 3 + k] * K[j * 3 + k];   // (G K^T)[i][j]
        ktg += K[k * 3 + i] * G[k * 3 + j];   // (K^T G)[i][j]
      }
      dK[i * 3 + j] = s * G[i * 3 + j] + oc * (gkt + ktg);
    }
  const float dd[3] = {dK[7] - dK[5], dK[2] - dK[6], dK[3] - dK[1]};
  // d = r / theta ; theta = ||r + eps||
  const float ddr = dd[0] * r[0] + dd[1] * r[1] + dd[2] * r[2];
  const float dtheta = dtheta_direct - ddr * inv * inv;
#pragma unroll
  for (int k = 0; k < 3; ++k) dr[k] = dd[k] * inv + dtheta * e[k] * inv;
}

}  // namespace b200smpl
