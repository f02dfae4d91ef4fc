"""Seeded synthetic (betas, pose, trans, cam) batches for benches / smoke:
betas ~ N(0,1), axis-angle pose ~ N(0, 0.3^2), trans ~ U(-1,1)^3, cam = [U(.6,1.2), U(-.2,.2), U(-.2,.2)].
Input *generation* only (host-side torch); the rotation matrices are made from the axis-angle
draw with the textbook Rodrigues formula so that the rotmat surface gets orthonormal inputs."""
import torch


def axis_angle_to_rotmat(r: torch.Tensor) -> torch.Tensor:
    theta = r.norm(dim=1, keepdim=True).clamp_min(1e-12)
    d = r / theta
    K = torch.zeros(r.shape[0], 3, 3, dtype=r.dtype)
    K[:, 0, 1], K[:, 0, 2] = -d[:, 2], d[:, 1]
    K[:, 1, 0], K[:, 1, 2] = d[:, 2], -d[:, 0]
    K[:, 2, 0], K[:, 2, 1] = -d[:, 1], d[:, 0]
    s, c = torch.sin(theta)[:, :, None], torch.cos(theta)[:, :, None]
    return torch.eye(3, dtype=r.dtype) + s * K + (1 - c) * (K @ K)


def make_smpl_inputs(B: int, seed: int = 0):
    g = torch.Generator().manual_seed(seed)
    betas = torch.randn(B, 10, generator=g)
    pose_aa = torch.randn(B, 72, generator=g) * 0.3
    trans = torch.rand(B, 3, generator=g) * 2 - 1
    cam = torch.stack([torch.rand(B, generator=g) * 0.6 + 0.6, torch.rand(B, generator=g) * 0.4 - 0.2,
                       torch.rand(B, generator=g) * 0.4 - 0.2], 1)
    rotmats = axis_angle_to_rotmat(pose_aa.double().reshape(-1, 3)).float().reshape(B, 24, 3, 3)
    return dict(betas=betas, pose_aa=pose_aa, rotmats=rotmats, trans=trans, cam=cam)


def make_upstream_grads(B: int, seed: int = 0, num_verts: int = 6890, num_joints: int = 90):
    g = torch.Generator().manual_seed(seed + 7919)
    return torch.randn(B, num_verts, 3, generator=g), torch.randn(B, num_joints, 3, generator=g)
