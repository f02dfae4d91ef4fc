"""Second, independent CPU restatement (numpy fp64, one body at a time) -- TEST INFRASTRUCTURE.

Textbook SMPL (Loper et al. 2015, eq. 2-4) written from the maths rather than from the smplx
op sequence: v' = sum_j w_vj G_j(theta, J) [I | -J_j] [p_v; 1].  It exists only to cross-check
`oracle/smpl_oracle.py` (two differently-structured restatements must agree to ~1e-12), since
the smplx part of the oracle is "parity unpinned" (see smpl_oracle.py header).
"""
from __future__ import annotations

import numpy as np


def rodrigues_one(r: np.ndarray) -> np.ndarray:
    """Same quirk as smplx: theta = ||r + 1e-8||, axis = r / theta."""
    theta = np.sqrt(np.sum((r + 1e-8) ** 2))
    d = r / theta
    K = np.array([[0.0, -d[2], d[1]], [d[2], 0.0, -d[0]], [-d[1], d[0], 0.0]])
    return np.eye(3) + np.sin(theta) * K + (1.0 - np.cos(theta)) * (K @ K)


def smpl_one_body(model, betas, rotmats, transl=None):
    """betas (10,), rotmats (24,3,3) -> vertices (6890,3), joints (90,3); all float64."""
    f = lambda k: np.asarray(model[k], dtype=np.float64)  # noqa: E731
    v_template, shapedirs, posedirs = f("v_template"), f("shapedirs"), f("posedirs")
    J_regressor, weights = f("J_regressor"), f("lbs_weights")
    parents = np.asarray(model["parents"])
    nj = len(parents)

    v_shaped = v_template + shapedirs @ betas                       # (V,3)
    J = J_regressor @ v_shaped                                       # (24,3)
    pose_feature = (rotmats[1:] - np.eye(3)).reshape(-1)             # (207,) joint-major, row-major 3x3
    v_posed = v_shaped + (pose_feature @ posedirs).reshape(-1, 3)

    G = np.zeros((nj, 4, 4))
    for j in range(nj):
        local = np.eye(4)
        local[:3, :3] = rotmats[j]
        local[:3, 3] = J[j] if j == 0 else J[j] - J[parents[j]]
        G[j] = local if j == 0 else G[parents[j]] @ local
    posed_joints = G[:, :3, 3].copy()

    verts = np.zeros_like(v_posed)
    for j in range(nj):
        unpose = np.eye(4)
        unpose[:3, 3] = -J[j]
        Gp = G[j] @ unpose
        w = weights[:, j:j + 1]
        verts += w * (v_posed @ Gp[:3, :3].T + Gp[:3, 3])

    if transl is not None:                      # smplx adds transl to vertices and the 45 joints,
        verts = verts + transl                  # the wrapper then regresses from translated vertices
        posed_joints = posed_joints + transl
    picked = verts[np.asarray(model["extra_joints_idxs"])]
    regs = [f(k) @ verts for k in ("J_regressor_extra", "J_regressor_cocoplus", "J_regressor_h36m")]
    joints = np.concatenate([posed_joints, picked] + regs, 0)
    return verts, joints
