#!/usr/bin/env python
"""Freeze golden input/output vectors by EXECUTING the reference's own in-tree Python files.

Run in the build container only (needs /root/reference); the GPU box never reads the reference.
Imports, unmodified, from /root/reference/Python/Soccer/PlayerReconstruction:
    utils/cam_utils.py            orthographic_project_torch, perspective_project_torch,
                                  convert_weak_perspective_to_camera_translation_torch (+ inverse),
                                  get_intrinsics_matrix
    utils/joints2d_utils.py       undo_keypoint_normalisation, check_joints2d_visibility_torch
    utils/rigid_transform_utils.py rot6d_to_rotmat
    losses/multi_task_loss.py     HomoscedasticUncertaintyWeightedMultiTaskLoss (joints2D / shape terms)
    config.py                     integer joint maps + FOCAL_LENGTH / REGRESSOR_IMG_WH
The smplx-backed `models/smpl_official.py` cannot be imported (smplx absent) -- that part of
the oracle stays "parity unpinned".

Also records the literal SMPL inputs embedded at PyTorch3DTest.py:106-199 (inputs only; the
reference records no expected outputs for them) by parsing that file's literals.

Output: tests/golden/intree_golden.npz
"""
import ast
import os
import re
import sys

import numpy as np
import torch

PR = "/root/reference/Python/Soccer/PlayerReconstruction"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "intree_golden.npz")


def literal_inputs():
    """Pull the bodypose / global_orient / betas / translation literals out of PyTorch3DTest.py."""
    src = open(os.path.join(PR, "PyTorch3DTest.py")).read().splitlines()
    seg = "\n".join(src[104:204])          # lines 105..204 (1-based)
    out = {}
    m = re.search(r"translation = torch\.Tensor\((\[.*?\])\)", seg)
    out["lit_translation"] = np.array(ast.literal_eval(m.group(1)), np.float32)
    for name in ("bodypose", "global_orient", "betas"):
        at = seg.index("\n    %s = [" % name) + len("\n    %s = " % name)
        depth, end = 0, None
        for i in range(at, len(seg)):           # the literal ends where its brackets balance
            depth += seg[i] == "["
            depth -= seg[i] == "]"
            if depth == 0:
                end = i + 1
                break
        out["lit_" + name] = np.array(ast.literal_eval(seg[at:end]), np.float32)
    assert out["lit_bodypose"].shape == (1, 23, 3, 3)
    assert out["lit_global_orient"].shape == (1, 1, 3, 3)
    assert out["lit_betas"].shape == (1, 10)
    return out


def main():
    if not os.path.isdir(PR):
        sys.exit("reference tree not found; run in the build container")
    sys.path.insert(0, PR)
    import config as ref_config                                    # noqa: E402
    from utils import cam_utils as ref_cam                         # noqa: E402
    from utils import joints2d_utils as ref_j2d                    # noqa: E402
    from utils import rigid_transform_utils as ref_rigid           # noqa: E402
    from losses.multi_task_loss import HomoscedasticUncertaintyWeightedMultiTaskLoss as RefLoss  # noqa: E402

    g = torch.Generator().manual_seed(20261018)
    B = 6
    G = {}

    # --- integer tables (bit-exact) -----------------------------------------------------------
    for k in ("ALL_JOINTS_TO_COCO_MAP", "ALL_JOINTS_TO_H36M_MAP", "H36M_TO_J17", "H36M_TO_J14",
              "SMPL_TO_KPRCNN_MAP"):
        G["cfg_" + k] = np.asarray(getattr(ref_config, k), np.int64)
    G["cfg_FOCAL_LENGTH"] = np.float64(ref_config.FOCAL_LENGTH)
    G["cfg_REGRESSOR_IMG_WH"] = np.int64(ref_config.REGRESSOR_IMG_WH)

    # --- orthographic projection --------------------------------------------------------------
    pts = torch.randn(B, 90, 3, generator=g)
    cam = torch.stack([torch.rand(B, generator=g) * 0.6 + 0.6,
                       torch.rand(B, generator=g) * 0.4 - 0.2,
                       torch.rand(B, generator=g) * 0.4 - 0.2], 1)
    G["ortho_points"], G["ortho_cam"] = pts.numpy(), cam.numpy()
    G["ortho_out"] = ref_cam.orthographic_project_torch(pts, cam).numpy()

    # --- wp <-> translation -------------------------------------------------------------------
    G["wp2t_out"] = ref_cam.convert_weak_perspective_to_camera_translation_torch(cam, 5000.0, 512).numpy()
    G["t2wp_out"] = ref_cam.convert_camera_translation_to_weak_perspective_torch(
        torch.from_numpy(G["wp2t_out"]), 5000.0, 512).numpy()
    G["intrinsics_512_5000"] = ref_cam.get_intrinsics_matrix(512, 512, 5000.0)

    # --- perspective projection ---------------------------------------------------------------
    rot = ref_rigid.rot6d_to_rotmat(torch.randn(B, 6, generator=g))
    trans = torch.from_numpy(G["wp2t_out"]).float()
    G["persp_rot"], G["persp_trans"] = rot.numpy(), trans.numpy()
    G["persp_out"] = ref_cam.perspective_project_torch(pts, rot, trans, focal_length=5000.0, img_wh=512).numpy()

    # --- keypoint (de)normalisation + visibility ----------------------------------------------
    j2d = torch.from_numpy(G["ortho_out"])
    G["undo_norm_out"] = ref_j2d.undo_keypoint_normalisation(j2d, 512).numpy()
    G["vis_out"] = ref_j2d.check_joints2d_visibility_torch(torch.from_numpy(G["undo_norm_out"]), 512).numpy()

    # --- rot6d -> rotmat ----------------------------------------------------------------------
    x6 = torch.randn(B, 144, generator=g)
    G["rot6d_in"] = x6.numpy()
    G["rot6d_out"] = ref_rigid.rot6d_to_rotmat(x6).numpy()                     # (B*24, 3, 3)

    # --- joints2D / shape loss terms ----------------------------------------------------------
    crit = RefLoss(losses_on=["joints2D"], init_loss_weights={"joints2D": 1.0})
    pred = torch.rand(B, 17, 2, generator=g) * 256
    label = torch.rand(B, 17, 2, generator=g) * 256
    total, _ = crit({"joints2D": label}, {"joints2D": pred})
    G["loss_j2d_pred"], G["loss_j2d_label"] = pred.numpy(), label.numpy()
    G["loss_j2d_out"] = total.detach().numpy()
    G["loss_j2d_log_var"] = crit.joints2D_log_var.detach().numpy()
    vis = torch.rand(B, 17, generator=g) > 0.3
    total_v, _ = crit({"joints2D": label, "vis": vis}, {"joints2D": pred})
    G["loss_j2d_vis"], G["loss_j2d_vis_out"] = vis.numpy(), total_v.detach().numpy()
    crit2 = RefLoss(losses_on=["joints2D", "shape_params"],
                    init_loss_weights={"joints2D": 100.0, "shape_params": 0.01})
    sp, sl = torch.randn(B, 10, generator=g), torch.randn(B, 10, generator=g)
    total2, _ = crit2({"joints2D": label, "shape_params": sl}, {"joints2D": pred, "shape_params": sp})
    G["loss2_shape_pred"], G["loss2_shape_label"] = sp.numpy(), sl.numpy()
    G["loss2_out"] = total2.detach().numpy()
    G["loss2_log_vars"] = np.array([crit2.joints2D_log_var.item(), crit2.shape_params_log_var.item()], np.float32)

    # --- round 2 additions (own generator: the arrays above stay bit-identical) --------------------------------
    g2 = torch.Generator().manual_seed(20262)
    # explicit intrinsics: batched, unequal focal lengths, skew, off-centre principal point, non-trivial third row
    camK = torch.zeros(B, 3, 3)
    camK[:, 0, 0] = 800 + 4200 * torch.rand(B, generator=g2)
    camK[:, 1, 1] = 800 + 4200 * torch.rand(B, generator=g2)
    camK[:, 0, 1] = 5 * torch.randn(B, generator=g2)
    camK[:, 1, 0] = 0.5 * torch.randn(B, generator=g2)
    camK[:, 0, 2] = 256 + 20 * torch.randn(B, generator=g2)
    camK[:, 1, 2] = 256 + 20 * torch.randn(B, generator=g2)
    camK[:, 2] = torch.tensor([0.01, -0.02, 1.1])
    G["persp_camK"] = camK.numpy()
    G["persp_camK_out"] = ref_cam.perspective_project_torch(pts, rot, trans, cam_K=camK).numpy()
    # all five MSE terms of the multi-task loss (verts, joints2D with vis, joints3D, shape, pose), unequal weights
    names = ["verts", "joints2D", "joints3D", "shape_params", "pose_params"]
    w5 = {"verts": 2.0, "joints2D": 0.3, "joints3D": 1.5, "shape_params": 0.05, "pose_params": 0.7}
    crit5 = RefLoss(losses_on=names, init_loss_weights=w5)
    o5 = {"verts": torch.randn(B, 50, 3, generator=g2), "joints2D": torch.rand(B, 17, 2, generator=g2) * 256,
          "joints3D": torch.randn(B, 17, 3, generator=g2), "shape_params": torch.randn(B, 10, generator=g2),
          "pose_params_rot_matrices": torch.randn(B, 24, 3, 3, generator=g2)}
    l5 = {k: v + 0.3 * torch.randn(v.shape, generator=g2) for k, v in o5.items()}
    l5["joints2D"] = torch.rand(B, 17, 2, generator=g2) * 256
    l5["vis"] = torch.rand(B, 17, generator=g2) > 0.25
    total5, parts5 = crit5(l5, o5)
    for k, v in o5.items():
        G["loss5_pred_" + k] = v.numpy()
    for k, v in l5.items():
        G["loss5_label_" + k] = v.numpy()
    G["loss5_out"] = total5.detach().numpy()
    G["loss5_parts"] = np.array([parts5[k].item() for k in names], np.float32)
    G["loss5_log_vars"] = np.array([getattr(crit5, k + "_log_var").item() for k in names], np.float32)

    G.update(literal_inputs())
    np.savez_compressed(OUT, **G)
    print("wrote", OUT, "with", len(G), "arrays")


if __name__ == "__main__":
    main()
