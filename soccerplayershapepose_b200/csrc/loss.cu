// Fused homoscedastic multi-task loss of the regressor training step (the reference's
// HomoscedasticUncertaintyWeightedMultiTaskLoss, PlayerReconstruction/losses/multi_task_loss.py:92-130, as used by
// PyTorch3DTest.py:1072-1106): five MSE(mean) terms -- vertices, 2D joints (orthographic projection of the
// mapped joints in pixels, both sides normalised 2x/wh - 1, optional visibility mask), 3D joints (mapped), shape
// parameters, pose rotation matrices -- each weighted  mse * exp(-log_var) + log_var  with trainable
// log-variances.  The reference evaluates it as ~40 eager PyTorch kernels that read the (B, 6890, 3) vertices
// four times forward and backward; here:
//
//   mtl_reduce    one pass: squared-error sums of all five terms (vertices as 8-byte vectors, grid-stride;
//                 one block-level reduction per term, fp32 atomics into 8 device scalars)
//   mtl_finalise  the loss, its five weighted parts, d loss / d log_var, and the per-term gradient scales
//   mtl_grad      one pass: dL/dvertices = scale * (v - label) (never more than one read of v and label and one
//                 write), dL/djoints (2D and 3D terms accumulated into the 90-joint gradient), dL/dcam,
//                 dL/dshape, dL/dpose -- all multiplied by the upstream gradient read from DEVICE memory
//
// Everything (upstream gradient, log-variances, counts) lives in device memory: the three launches are CUDA-graph
// capturable and never synchronise.
#include "common.cuh"

namespace b200smpl {

namespace {

constexpr int T_VERTS = 0, T_J2D = 1, T_J3D = 2, T_SHAPE = 3, T_POSE = 4;

// scratch layout (floats): [0..4] squared-error sums, [5] visible (body, joint) pairs of the 2D term,
// [8..12] gradient scales 2 * exp(-lv) / N per term (0 for an absent term)
constexpr int SC_SUM = 0, SC_VIS = 5, SC_SCALE = 8;

__device__ __forceinline__ float block_reduce(float v, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  v = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.f;
  if (w == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  }
  return v;                                   // valid in thread 0
}

struct MtlArgs {
  const float *verts, *verts_label;          // [B][V][3] or null (term absent)
  const float *joints, *cam;                 // [B][NJ][3], [B][3]
  const int32_t* map2d;                      // [n2] or null
  const float* label2d;                      // [B][n2][2] pixels
  const uint8_t* vis;                        // [B][n2] or null
  const int32_t* map3d;                      // [n3] or null
  const float* label3d;                      // [B][n3][3]
  const float *shape, *shape_label;          // [B][nb] or null
  const float *pose, *pose_label;            // [B][npose] or null
  const float* log_var;                      // [5] device
  long long nv_elems;                        // B * V * 3
  int B, NJ, n2, n3, nb, npose;
  float proj_wh, norm_wh;
};

__device__ __forceinline__ float sq(float x) { return x * x; }

__global__ void __launch_bounds__(256)
mtl_reduce_kernel(const MtlArgs a, float* __restrict__ scratch) {
  __shared__ float sh[32];
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nth = (long long)gridDim.x * blockDim.x;
  float s_v = 0.f, s_2 = 0.f, s_3 = 0.f, s_s = 0.f, s_p = 0.f, n_vis = 0.f;
  if (a.verts != nullptr) {
    const long long n2v = a.nv_elems >> 1;                          // 8-byte vectors (B * V * 3 is even or handled below)
    const float2* v = reinterpret_cast<const float2*>(a.verts);
    const float2* l = reinterpret_cast<const float2*>(a.verts_label);
    for (long long i = tid; i < n2v; i += nth) {
      const float2 x = __ldg(v + i), y = __ldg(l + i);
      s_v += sq(x.x - y.x) + sq(x.y - y.y);
    }
    if ((a.nv_elems & 1) && tid == 0) s_v += sq(a.verts[a.nv_elems - 1] - a.verts_label[a.nv_elems - 1]);
  }
  if (a.map2d != nullptr) {
    const long long n = (long long)a.B * a.n2;
    for (long long i = tid; i < n; i += nth) {
      const int b = (int)(i / a.n2), k = (int)(i - (long long)b * a.n2);
      if (a.vis != nullptr && !a.vis[i]) continue;
      n_vis += 1.f;
      const float s = a.cam[b * 3], tx = a.cam[b * 3 + 1], ty = a.cam[b * 3 + 2];
      const float* p = a.joints + ((size_t)b * a.NJ + a.map2d[k]) * 3;
      const float t[2] = {tx, ty};
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const float u = s * (p[c] + t[c]);                          // utils/cam_utils.py:5-26
        const float pix = (u + 1.f) * (a.proj_wh / 2.0f);           // utils/joints2d_utils.py:5-10
        const float pn = (2.0f * pix) / a.norm_wh - 1.0f;           // losses/multi_task_loss.py:101-102
        const float ln = (2.0f * a.label2d[i * 2 + c]) / a.norm_wh - 1.0f;
        s_2 += sq(pn - ln);
      }
    }
  }
  if (a.map3d != nullptr) {
    const long long n = (long long)a.B * a.n3 * 3;
    for (long long i = tid; i < n; i += nth) {
      const long long bj = i / 3;
      const int c = (int)(i - bj * 3), b = (int)(bj / a.n3), k = (int)(bj - (long long)b * a.n3);
      s_3 += sq(a.joints[((size_t)b * a.NJ + a.map3d[k]) * 3 + c] - a.label3d[i]);
    }
  }
  if (a.shape != nullptr)
    for (long long i = tid; i < (long long)a.B * a.nb; i += nth) s_s += sq(a.shape[i] - a.shape_label[i]);
  if (a.pose != nullptr)
    for (long long i = tid; i < (long long)a.B * a.npose; i += nth) s_p += sq(a.pose[i] - a.pose_label[i]);
  float* sums = scratch + SC_SUM;
  float r;
  r = block_reduce(s_v, sh); if (threadIdx.x == 0 && a.verts) atomicAdd(sums + T_VERTS, r);
  r = block_reduce(s_2, sh); if (threadIdx.x == 0 && a.map2d) atomicAdd(sums + T_J2D, r);
  r = block_reduce(s_3, sh); if (threadIdx.x == 0 && a.map3d) atomicAdd(sums + T_J3D, r);
  r = block_reduce(s_s, sh); if (threadIdx.x == 0 && a.shape) atomicAdd(sums + T_SHAPE, r);
  r = block_reduce(s_p, sh); if (threadIdx.x == 0 && a.pose) atomicAdd(sums + T_POSE, r);
  r = block_reduce(n_vis, sh); if (threadIdx.x == 0 && a.map2d) atomicAdd(scratch + SC_VIS, r);
}

// out[0] = loss, out[1..5] = weighted parts mse * exp(-lv), out[6..10] = d loss / d log_var (upstream gradient 1)
__global__ void mtl_finalise_kernel(const MtlArgs a, float* __restrict__ scratch, float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const bool on[5] = {a.verts != nullptr, a.map2d != nullptr, a.map3d != nullptr, a.shape != nullptr, a.pose != nullptr};
  const float cnt[5] = {(float)a.nv_elems, 2.f * scratch[SC_VIS], (float)a.B * a.n3 * 3.f, (float)a.B * a.nb,
                        (float)a.B * a.npose};
  float total = 0.f;
  for (int t = 0; t < 5; ++t) {
    float part = 0.f, dlv = 0.f, scale = 0.f;
    if (on[t]) {
      const float lv = a.log_var[t], ew = expf(-lv);
      const float mse = cnt[t] > 0.f ? scratch[SC_SUM + t] / cnt[t] : 0.f / 0.f;   // empty mean is NaN, as torch.mean
      part = mse * ew;
      total += part + lv;
      dlv = 1.f - part;
      scale = cnt[t] > 0.f ? 2.f * ew / cnt[t] : 0.f;
    }
    out[1 + t] = part;
    out[6 + t] = dlv;
    scratch[SC_SCALE + t] = scale;
  }
  out[0] = total;
}

// one launch for every gradient.  blocks [0, vblocks): the vertex gradient (grid-stride over 8-byte vectors);
// blocks [vblocks, vblocks + B): one body each for joints / cam / shape / pose.
__global__ void __launch_bounds__(256)
mtl_grad_kernel(const MtlArgs a, const float* __restrict__ scratch, const float* __restrict__ gup, int vblocks,
                float* __restrict__ gverts, float* __restrict__ gjoints, float* __restrict__ gcam,
                float* __restrict__ gshape, float* __restrict__ gpose) {
  __shared__ float sh[32];
  const float g = gup != nullptr ? *gup : 1.f;
  if ((int)blockIdx.x < vblocks) {
    if (a.verts == nullptr || gverts == nullptr) return;
    const float c = g * scratch[SC_SCALE + T_VERTS];
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)vblocks * blockDim.x;
    const long long n2v = a.nv_elems >> 1;
    const float2* v = reinterpret_cast<const float2*>(a.verts);
    const float2* l = reinterpret_cast<const float2*>(a.verts_label);
    float2* o = reinterpret_cast<float2*>(gverts);
    for (long long i = tid; i < n2v; i += nth) {
      const float2 x = __ldg(v + i), y = __ldg(l + i);
      o[i] = make_float2(c * (x.x - y.x), c * (x.y - y.y));
    }
    if ((a.nv_elems & 1) && tid == 0)
      gverts[a.nv_elems - 1] = c * (a.verts[a.nv_elems - 1] - a.verts_label[a.nv_elems - 1]);
    return;
  }
  const int b = (int)blockIdx.x - vblocks;
  if (b >= a.B) return;
  // joints: zero, then the 2D and 3D terms accumulate (the two maps share joints: COCO in both)
  if (gjoints != nullptr)
    for (int i = threadIdx.x; i < a.NJ * 3; i += blockDim.x) gjoints[(size_t)b * a.NJ * 3 + i] = 0.f;
  __syncthreads();
  float gs = 0.f, gtx = 0.f, gty = 0.f;
  if (a.map2d != nullptr) {
    const float c2 = g * scratch[SC_SCALE + T_J2D];
    const float s = a.cam[b * 3], tx = a.cam[b * 3 + 1], ty = a.cam[b * 3 + 2];
    for (int k = threadIdx.x; k < a.n2; k += blockDim.x) {
      const size_t i = (size_t)b * a.n2 + k;
      if (a.vis != nullptr && !a.vis[i]) continue;
      const int J = a.map2d[k];
      const float* p = a.joints + ((size_t)b * a.NJ + J) * 3;
      const float t[2] = {tx, ty};
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const float u = s * (p[cc] + t[cc]);
        const float pix = (u + 1.f) * (a.proj_wh / 2.0f);
        const float pn = (2.0f * pix) / a.norm_wh - 1.0f;
        const float ln = (2.0f * a.label2d[i * 2 + cc]) / a.norm_wh - 1.0f;
        const float gu = c2 * (pn - ln) * (2.0f / a.norm_wh) * (a.proj_wh / 2.0f);      // d loss / d u
        if (gjoints != nullptr) atomicAdd(&gjoints[((size_t)b * a.NJ + J) * 3 + cc], s * gu);
        gs += gu * (p[cc] + t[cc]);
        if (cc == 0) gtx += s * gu; else gty += s * gu;
      }
    }
  }
  if (a.map3d != nullptr && gjoints != nullptr) {
    const float c3 = g * scratch[SC_SCALE + T_J3D];
    for (int i = threadIdx.x; i < a.n3 * 3; i += blockDim.x) {
      const int k = i / 3, cc = i - k * 3;
      const size_t jo = ((size_t)b * a.NJ + a.map3d[k]) * 3 + cc;
      atomicAdd(&gjoints[jo], c3 * (a.joints[jo] - a.label3d[((size_t)b * a.n3 + k) * 3 + cc]));
    }
  }
  if (gcam != nullptr) {
    const float r0 = block_reduce(gs, sh), r1 = block_reduce(gtx, sh), r2 = block_reduce(gty, sh);
    if (threadIdx.x == 0) { gcam[b * 3] = r0; gcam[b * 3 + 1] = r1; gcam[b * 3 + 2] = r2; }
  }
  if (a.shape != nullptr && gshape != nullptr) {
    const float cs = g * scratch[SC_SCALE + T_SHAPE];
    for (int i = threadIdx.x; i < a.nb; i += blockDim.x)
      gshape[(size_t)b * a.nb + i] = cs * (a.shape[(size_t)b * a.nb + i] - a.shape_label[(size_t)b * a.nb + i]);
  }
  if (a.pose != nullptr && gpose != nullptr) {
    const float cp = g * scratch[SC_SCALE + T_POSE];
    for (int i = threadIdx.x; i < a.npose; i += blockDim.x)
      gpose[(size_t)b * a.npose + i] = cp * (a.pose[(size_t)b * a.npose + i] - a.pose_label[(size_t)b * a.npose + i]);
  }
}

int fill_args(const b200smpl_multitask_loss_args* in, MtlArgs& a) {
  if (in == nullptr) return fail(B200SMPL_ERR_INVALID, "null args");
  if (in->batch < 1) return fail(B200SMPL_ERR_INVALID, "batch must be >= 1");
  if (in->log_var == nullptr || in->scratch == nullptr) return fail(B200SMPL_ERR_INVALID, "log_var and scratch are required");
  if ((in->verts == nullptr) != (in->verts_label == nullptr) || (in->shape == nullptr) != (in->shape_label == nullptr) ||
      (in->pose == nullptr) != (in->pose_label == nullptr))
    return fail(B200SMPL_ERR_INVALID, "a prediction and its label must be given together");
  if ((in->map2d != nullptr || in->map3d != nullptr) && in->joints == nullptr)
    return fail(B200SMPL_ERR_INVALID, "the joint terms need joints");
  if (in->map2d != nullptr && (in->cam == nullptr || in->label2d == nullptr))
    return fail(B200SMPL_ERR_INVALID, "the 2D joint term needs cam and label2d");
  if (in->map3d != nullptr && in->label3d == nullptr) return fail(B200SMPL_ERR_INVALID, "the 3D joint term needs label3d");
  if (in->verts != nullptr && ((reinterpret_cast<uintptr_t>(in->verts) | reinterpret_cast<uintptr_t>(in->verts_label)) & 7))
    return fail(B200SMPL_ERR_INVALID, "verts and verts_label must be 8-byte aligned");
  a.verts = in->verts; a.verts_label = in->verts_label; a.joints = in->joints; a.cam = in->cam;
  a.map2d = in->map2d; a.label2d = in->label2d; a.vis = in->vis; a.map3d = in->map3d; a.label3d = in->label3d;
  a.shape = in->shape; a.shape_label = in->shape_label; a.pose = in->pose; a.pose_label = in->pose_label;
  a.log_var = in->log_var;
  a.B = in->batch; a.NJ = in->num_joints; a.n2 = in->nmap2d; a.n3 = in->nmap3d; a.nb = in->num_betas; a.npose = in->pose_cols;
  a.nv_elems = (long long)in->batch * in->num_verts * 3;
  a.proj_wh = in->proj_wh; a.norm_wh = in->norm_wh;
  return 0;
}

int vertex_blocks(const MtlArgs& a) {
  if (a.verts == nullptr) return 0;
  return (int)std::max<long long>(1, std::min<long long>(148 * 8, (a.nv_elems / 2 + 255) / 256));
}

}  // namespace
}  // namespace b200smpl

using namespace b200smpl;

extern "C" {

int b200smpl_multitask_loss(const b200smpl_multitask_loss_args* args, float* out, void* stream) {
  B200_NVTX("b200smpl_multitask_loss");
  MtlArgs a{};
  int rc = fill_args(args, a);
  if (rc) return rc;
  if (out == nullptr) return fail(B200SMPL_ERR_INVALID, "null out");
  cudaStream_t st = (cudaStream_t)stream;
  float* scratch = args->scratch;
  B200_CUDA_TRY(cudaMemsetAsync(scratch, 0, 16 * sizeof(float), st));
  const int blocks = std::max(vertex_blocks(a), std::min(148 * 4, (a.B * std::max(a.n2, 1) + 255) / 256));
  {
    LaunchTimer _timer("mtl_reduce", st);
    mtl_reduce_kernel<<<blocks, 256, 0, st>>>(a, scratch);
    B200_LAUNCH_CHECK("mtl_reduce");
  }
  {
    LaunchTimer _timer("mtl_finalise", st);
    mtl_finalise_kernel<<<1, 32, 0, st>>>(a, scratch, out);
    B200_LAUNCH_CHECK("mtl_finalise");
  }
  return 0;
}

int b200smpl_multitask_loss_backward(const b200smpl_multitask_loss_args* args, const float* grad_loss, float* grad_verts,
                                     float* grad_joints, float* grad_cam, float* grad_shape, float* grad_pose,
                                     void* stream) {
  B200_NVTX("b200smpl_multitask_loss_backward");
  MtlArgs a{};
  int rc = fill_args(args, a);
  if (rc) return rc;
  if (grad_verts != nullptr && (reinterpret_cast<uintptr_t>(grad_verts) & 7))
    return fail(B200SMPL_ERR_INVALID, "grad_verts must be 8-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int vb = grad_verts != nullptr ? vertex_blocks(a) : 0;
  LaunchTimer _timer("mtl_grad", st);
  mtl_grad_kernel<<<vb + a.B, 256, 0, st>>>(a, args->scratch, grad_loss, vb, grad_verts, grad_joints, grad_cam, grad_shape,
                                            grad_pose);
  B200_LAUNCH_CHECK("mtl_grad");
  return 0;
}

}  // extern "C"
