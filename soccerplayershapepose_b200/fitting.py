"""Batched per-player SMPL fitting on 2D keypoints (BASELINE.json configs[2]).

Restates the optimisation loop of the reference's `single_view_optimization`
(PlayerReconstruction/player_recon.py:1172-1294) for a batch of independent players:

  * parameters, as there: global orientation and body pose as raw rotation matrices (`pose2rot=False`),
    weak-perspective camera `[s, tx, ty]` and betas; the hands / feet joints of the body pose (body_pose[:, 6:8]
    and body_pose[:, 21:], player_recon.py:1175-1177, 1202-1206) stay at their initial value;
  * loss per player: joints2D term of the multi-task loss on the 17 COCO joints, after
    `orthographic_project_torch` and `undo_keypoint_normalisation` (player_recon.py:1217-1221,
    losses/multi_task_loss.py:97-113), optionally plus a shape prior `shape_weight * mean(betas^2)`
    (BASELINE.json asks for one; the reference loop has none -- default 0 reproduces the reference);
    the silhouette term of the reference needs its neural renderer and is out of scope;
  * optimiser: `torch.optim.Adam(params, lr)` arithmetic (player_recon.py:1199), elementwise, so a batch of
    players with a summed loss is exactly a set of independent per-player optimisations;
  * the best iterate per player is kept (player_recon.py:1254-1266) -- here on the device, by loss value,
    instead of a host round trip per iteration;
  * results under the reference's `.npz` keys (player_recon.py:1293-1294): body_pose, global_orient, betas,
    translation (`convert_weak_perspective_to_camera_translation`, cam_utils.py:44-52).

Only the joints are computed (virtual rows of the blend GEMM, no vertex is skinned).  One iteration = 9 kernel
launches through the C-ABI; after one eager iteration the iteration is captured in a CUDA graph and replayed.
CUDA float32 only: there is no CPU path.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional

import torch

from . import _lib, config
from .cam_utils import convert_weak_perspective_to_camera_translation_torch
from .smpl import SMPL

FROZEN_FULL_JOINTS = (7, 8, 22, 23)       # body_pose[:, 6:8] and body_pose[:, 21:] in full-pose numbering


class BatchedFitter:
    def __init__(self, smpl: SMPL, lr: float = 1e-3, shape_weight: float = 0.0, joints2d_log_var: float = 0.0,
                 proj_wh: float = 512.0, norm_wh: float = float(config.REGRESSOR_IMG_WH),
                 betas: tuple = (0.9, 0.999), eps: float = 1e-8, use_cuda_graph: bool = True,
                 frozen_joints=FROZEN_FULL_JOINTS, mode: Optional[str] = None):
        dev = next(smpl.buffers()).device
        if dev.type != "cuda":
            raise RuntimeError("BatchedFitter needs the SMPL module on a CUDA device (no CPU path)")
        self.smpl, self.dev = smpl, dev
        self.eng = smpl._engine(dev)
        self.mode = _lib.MODES[mode] if mode is not None else smpl.mode
        self.lib = _lib.load()
        self.lr, self.shape_weight, self.log_var = float(lr), float(shape_weight), float(joints2d_log_var)
        self.proj_wh, self.norm_wh = float(proj_wh), float(norm_wh)
        self.b1, self.b2, self.eps = float(betas[0]), float(betas[1]), float(eps)
        self.use_graph = use_cuda_graph
        self.jmap = torch.tensor(config.SMPL_TO_KPRCNN_MAP, dtype=torch.int32, device=dev)
        frozen = torch.zeros(24, 9, dtype=torch.uint8)
        for j in frozen_joints:
            frozen[j] = 1
        self.frozen_rot = frozen.reshape(-1).to(dev)
        self._graph = None
        self._state = None

    # ---- one iteration: every call below is one C-ABI entry point on the current stream ----------------
    def _iteration(self, st: Dict[str, torch.Tensor]) -> None:
        lib, B = self.lib, st["rot"].shape[0]
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
        _, joints, _ = self.eng.forward(st["betas"], st["rot"], None, None, axis_angle=False, mode=self.mode,
                                        want_vertices=False)
        vis = st.get("vis")
        _lib.check(lib.b200smpl_fit_loss(
            joints.data_ptr(), st["cam"].data_ptr(), self.jmap.data_ptr(), st["label"].data_ptr(),
            None if vis is None else vis.data_ptr(), st["betas"].data_ptr(), B, joints.shape[1], self.jmap.numel(),
            st["betas"].shape[1], self.proj_wh, self.norm_wh, self.log_var, self.shape_weight, st["loss"].data_ptr(),
            st["gj"].data_ptr(), st["gcam"].data_ptr(), st["gbetas_prior"].data_ptr(), stream), "fit_loss")
        gb, gp, _, _ = self.eng.backward(st["betas"], st["rot"], None, None, None, None, st["gj"], None,
                                         axis_angle=False, mode=self.mode, need_transl=False, need_cam=False)
        _lib.check(lib.b200smpl_fit_mark_best(st["loss"].data_ptr(), st["best_loss"].data_ptr(),
                                              st["best_iter"].data_ptr(), st["improved"].data_ptr(),
                                              st["step"].data_ptr(), B, stream), "fit_mark_best")
        for name, grad, extra, frozen, commit in (("rot", gp, None, self.frozen_rot, 0),
                                                  ("betas", gb, st["gbetas_prior"], None, 0),
                                                  ("cam", st["gcam"], None, None, 1)):
            p = st[name]
            _lib.check(lib.b200smpl_fit_adam_step(
                p.data_ptr(), grad.data_ptr(), None if extra is None else extra.data_ptr(), st["m_" + name].data_ptr(),
                st["v_" + name].data_ptr(), st["best_" + name].data_ptr(), st["improved"].data_ptr(),
                None if frozen is None else frozen.data_ptr(), st["step"].data_ptr(), commit, B,
                p.numel() // B, self.lr, self.b1, self.b2, self.eps, stream), "fit_adam_step")

    def fit(self, rotmats: torch.Tensor, betas: torch.Tensor, cam: torch.Tensor, keypoints2d: torch.Tensor,
            vis: Optional[torch.Tensor] = None, iterations: int = 200) -> Dict[str, torch.Tensor]:
        """rotmats (B,24,3,3) initial pose, betas (B,10), cam (B,3) weak-perspective [s,tx,ty], keypoints2d
        (B,17,2) target COCO keypoints in pixels of the `proj_wh` image, vis (B,17) bool or None."""
        dev = self.dev
        B = rotmats.shape[0]
        f32 = dict(dtype=torch.float32, device=dev)
        st = {"rot": rotmats.to(**f32).reshape(B, 216).clone().contiguous(), "betas": betas.to(**f32).clone().contiguous(),
              "cam": cam.to(**f32).clone().contiguous(), "label": keypoints2d.to(**f32).contiguous()}
        if vis is not None:
            st["vis"] = vis.to(device=dev, dtype=torch.uint8).contiguous()
        for name in ("rot", "betas", "cam"):
            st["m_" + name] = torch.zeros_like(st[name])
            st["v_" + name] = torch.zeros_like(st[name])
            st["best_" + name] = st[name].clone()
        st["loss"] = torch.zeros(B, **f32)
        st["best_loss"] = torch.full((B,), float("inf"), **f32)
        st["best_iter"] = torch.zeros(B, dtype=torch.int32, device=dev)
        st["improved"] = torch.zeros(B, dtype=torch.uint8, device=dev)
        st["step"] = torch.zeros(2, dtype=torch.int32, device=dev)
        st["gj"] = torch.zeros((B, self.eng.num_joints_out, 3), **f32)
        st["gcam"] = torch.zeros((B, 3), **f32)
        st["gbetas_prior"] = torch.zeros((B, st["betas"].shape[1]), **f32)
        first_loss = None
        if iterations > 0:
            if self.use_graph and iterations > 2:
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):
                    self._iteration(st)                       # eager warm-up: allocations, attribute set-up
                    first_loss = st["loss"].clone()
                torch.cuda.current_stream(dev).wait_stream(side)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    self._iteration(st)
                for _ in range(iterations - 2):
                    graph.replay()
                # the capture itself does not execute: iterations = 1 eager + (iterations - 2) replays + 1 below
                graph.replay()
                self._graph = graph
            else:
                for i in range(iterations):
                    self._iteration(st)
                    if i == 0:
                        first_loss = st["loss"].clone()
        best_rot = st["best_rot"].reshape(B, 24, 3, 3)
        out = {"body_pose": best_rot[:, 1:].contiguous(), "global_orient": best_rot[:, :1].contiguous(),
               "betas": st["best_betas"], "cam": st["best_cam"],
               "translation": convert_weak_perspective_to_camera_translation_torch(
                   st["best_cam"], config.FOCAL_LENGTH, self.proj_wh),
               "best_loss": st["best_loss"], "best_iter": st["best_iter"], "initial_loss": first_loss,
               "final_rotmats": st["rot"].reshape(B, 24, 3, 3), "final_betas": st["betas"], "final_cam": st["cam"],
               "last_loss": st["loss"]}
        self._state = st
        return out
