#!/usr/bin/env python
"""Freeze known-answer vectors for the SMPL path: the fp64 oracle evaluated on the seeded
synthetic model (model_io.make_synthetic_smpl(1234)) for seeded inputs, plus the literal inputs
the reference embeds at PyTorch3DTest.py:106-202 (read from intree_golden.npz).

These are the BUILD'S OWN known answers (the reference pins nothing for
this path -- "parity unpinned").  They guard the oracle and the model generator against drift
and give the GPU tests a committed target that does not depend on running the oracle.

Output: tests/golden/smpl_kat.npz  (vertices at a fixed 256-vertex subset + all 90 joints +
per-body vertex checksums, float64 values stored as float64).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from oracle.smpl_oracle import SMPLOracle, batch_rodrigues                # noqa: E402
from soccerplayershapepose_b200.model_io import make_synthetic_smpl       # noqa: E402


def make_inputs(B, seed):
    g = torch.Generator().manual_seed(seed)
    betas = torch.randn(B, 10, generator=g, dtype=torch.float64)
    pose_aa = torch.randn(B, 72, generator=g, dtype=torch.float64) * 0.3
    trans = torch.rand(B, 3, generator=g, dtype=torch.float64) * 2 - 1
    return betas, pose_aa, trans


def main():
    model = make_synthetic_smpl(1234)
    orc = SMPLOracle(model, dtype=torch.float64)
    lit = np.load(os.path.join(HERE, "intree_golden.npz"))
    betas, pose_aa, trans = make_inputs(5, seed=7)
    rot = batch_rodrigues(pose_aa.reshape(-1, 3)).reshape(5, 24, 3, 3)
    # body 5 = the reference's literal (approximately orthonormal, 4-decimal) rotation matrices
    lit_rot = torch.from_numpy(np.concatenate([lit["lit_global_orient"], lit["lit_bodypose"]], 1)).double()
    rot = torch.cat([rot, lit_rot], 0)
    betas = torch.cat([betas, torch.from_numpy(lit["lit_betas"]).double()], 0)
    trans = torch.cat([trans, torch.from_numpy(lit["lit_translation"]).double()[None]], 0)
    # rotmat surface (pose2rot=False), with transl
    out = orc.forward_flat(betas, rot, trans, pose2rot=False)
    # axis-angle surface (pose2rot=True), no transl
    out_aa = orc.forward_flat(betas[:5], pose_aa, None, pose2rot=True)
    sub = np.random.default_rng(99).choice(6890, 256, replace=False)
    sub.sort()
    np.savez_compressed(
        os.path.join(HERE, "smpl_kat.npz"),
        betas=betas.numpy(), rotmats=rot.numpy(), pose_aa=pose_aa.numpy(), trans=trans.numpy(),
        vert_subset=sub,
        verts_sub=out.vertices[:, sub].numpy(), joints=out.joints.numpy(),
        verts_sum=out.vertices.sum(1).numpy(), verts_abs_sum=out.vertices.abs().sum(1).numpy(),
        aa_verts_sub=out_aa.vertices[:, sub].numpy(), aa_joints=out_aa.joints.numpy(),
        aa_verts_sum=out_aa.vertices.sum(1).numpy(),
    )
    print("wrote smpl_kat.npz")


if __name__ == "__main__":
    main()
