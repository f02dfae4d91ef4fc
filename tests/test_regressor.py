"""Regressor head + multi-task loss: CPU checks with the oracle standing in for the SMPL
kernels, a world_size-2 gloo run of the data-parallel step, and (GPU) the same step through the C-ABI kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import smpl_oracle as O
from soccerplayershapepose_b200 import config, regressor

LOSSES = ("verts", "joints2D", "joints3D", "shape_params", "pose_params")
WEIGHTS = {"verts": 1.0, "joints2D": 0.1, "joints3D": 1.0, "shape_params": 0.1, "pose_params": 0.1}


class _OracleSMPL:
    def __init__(self, model, dtype):
        self.orc = O.SMPLOracle(model, dtype=dtype)

    def __call__(self, body_pose, global_orient, betas, pose2rot=False, return_verts=True):
        return self.orc.forward(betas, body_pose, global_orient, None, pose2rot)


def _cpu_ops():
    return O.rot6d_to_rotmat, (lambda joints, cam: O.undo_keypoint_normalisation(O.orthographic_project(joints, cam), 512))


def _batch(model, B, seed, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    feats = torch.randn(B, 64, generator=g)
    betas = torch.randn(B, 10, generator=g)
    pose = torch.randn(B, 72, generator=g) * 0.3
    rot = O.batch_rodrigues(pose.reshape(-1, 3)).reshape(B, 24, 3, 3)
    cam = torch.tensor([0.9, 0.0, 0.0]).repeat(B, 1)
    # labels in float64 (their float32 rounding would depend on the thread count of the process that makes them)
    betas, rot, cam = betas.double(), rot.double(), cam.double()
    out = O.SMPLOracle(model, dtype=torch.float64).forward(betas, rot[:, 1:], rot[:, :1], None, False)
    j2d = O.undo_keypoint_normalisation(O.orthographic_project(out.joints, cam), 512)[:, config.SMPL_TO_KPRCNN_MAP, :]
    labels = {"joints2D": j2d, "verts": out.vertices, "shape_params": betas, "pose_params_rot_matrices": rot,
              "joints3D": out.joints[:, config.ALL_JOINTS_TO_COCO_MAP, :]}
    return feats.to(dtype), {k: v.to(dtype) for k, v in labels.items()}


def _make(seed=0, dtype=torch.float32):
    torch.manual_seed(seed)
    head = regressor.IEFModule((48, 48), in_features=64).to(dtype)
    crit = regressor.MultiTaskLoss(LOSSES, WEIGHTS).to(dtype)
    return head, crit


def test_head_shapes_and_identity_init(synthetic_model):
    head, crit = _make()
    cam, pose6d, shape = head(torch.zeros(3, 64))
    assert cam.shape == (3, 3) and pose6d.shape == (3, 144) and shape.shape == (3, 10)
    with torch.no_grad():
        for fc in (head.fc1, head.fc2, head.fc3):
            fc.weight.zero_()
    cam, pose6d, shape = head(torch.randn(2, 64))
    R = O.rot6d_to_rotmat(pose6d).view(2, 24, 3, 3)
    assert torch.allclose(R, torch.eye(3).expand(2, 24, 3, 3), atol=1e-6)      # mean pose = identity rotations
    assert torch.allclose(cam, torch.tensor([0.9, 0.0, 0.0]).expand(2, 3))
    assert abs(float(crit.joints2D_log_var) + np.log(0.1 + 1e-6)) < 1e-6       # multi_task_loss.py:38


def test_train_step_reduces_loss(synthetic_model):
    head, crit = _make()
    feats, labels = _batch(synthetic_model, 4, 1)
    smpl = _OracleSMPL(synthetic_model, torch.float32)
    r6, proj = _cpu_ops()
    opt = torch.optim.Adam(list(head.parameters()) + list(crit.parameters()), lr=1e-3)
    losses = [float(regressor.train_step(head, crit, opt, smpl, feats, labels, r6, proj)) for _ in range(5)]
    assert losses[-1] < losses[0]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _ddp_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from soccerplayershapepose_b200.model_io import make_synthetic_smpl
        from soccerplayershapepose_b200 import sharding
        torch.set_num_threads(2)
        model = make_synthetic_smpl(1234)
        head, crit = _make(dtype=torch.float64)

        class Both(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.head, self.crit = head, crit

            def forward(self, feats, labels, smpl, r6, proj):
                outputs = regressor.predict(self.head, smpl, feats, r6, proj)
                return self.crit(labels, outputs)[0]

        ddp = torch.nn.parallel.DistributedDataParallel(Both())
        feats, labels = _batch(model, 8, 5, torch.float64)
        f = sharding.shard(feats, rank, world)
        lab = {k: sharding.shard(v, rank, world) for k, v in labels.items()}
        r6, proj = _cpu_ops()
        loss = ddp(f, lab, _OracleSMPL(model, torch.float64), r6, proj)
        loss.backward()                                   # gradient all-reduce (mean) happens here
        if rank == 0:
            np.savez(os.path.join(out_dir, "g.npz"), fc3=head.fc3.weight.grad.numpy(), fc1=head.fc1.weight.grad.numpy(),
                     lv=crit.verts_log_var.grad.numpy())
    finally:
        dist.destroy_process_group()


def test_data_parallel_gradients_match_full_batch(tmp_path, synthetic_model):
    mp.spawn(_ddp_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    got = np.load(os.path.join(str(tmp_path), "g.npz"))
    head, crit = _make(dtype=torch.float64)
    feats, labels = _batch(synthetic_model, 8, 5, torch.float64)
    r6, proj = _cpu_ops()
    outputs = regressor.predict(head, _OracleSMPL(synthetic_model, torch.float64), feats, r6, proj)
    crit(labels, outputs)[0].backward()
    np.testing.assert_allclose(got["fc3"], head.fc3.weight.grad.numpy(), rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(got["fc1"], head.fc1.weight.grad.numpy(), rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(got["lv"], crit.verts_log_var.grad.numpy(), rtol=1e-9)


@pytest.mark.gpu
def test_gpu_step_matches_oracle(synthetic_model):
    from soccerplayershapepose_b200.smpl import SMPL
    dev = torch.device("cuda", 0)
    feats, labels = _batch(synthetic_model, 6, 9)
    head64, crit64 = _make(dtype=torch.float64)
    r6, proj = _cpu_ops()
    out64 = regressor.predict(head64, _OracleSMPL(synthetic_model, torch.float64), feats.double(), r6, proj)
    loss64 = crit64({k: v.double() for k, v in labels.items()}, out64)[0]
    loss64.backward()
    head, crit = _make()
    head, crit = head.to(dev), crit.to(dev)
    smpl = SMPL(model_data=synthetic_model, mode="fp32").to(dev)
    g6, gproj = regressor.gpu_ops()
    out = regressor.predict(head, smpl, feats.to(dev), g6, gproj)
    loss = crit({k: v.to(dev) for k, v in labels.items()}, out)[0]
    loss.backward()
    assert abs(loss.item() - loss64.item()) < 1e-4 * abs(loss64.item())
    for name in ("fc1", "fc3"):
        g = getattr(head, name).weight.grad.cpu().double()
        g64 = getattr(head64, name).weight.grad
        assert (g - g64).abs().max().item() < 2e-4 * g64.abs().max().item()


@pytest.mark.gpu
def test_graphed_step_matches_eager_steps(synthetic_model):
    """The CUDA-graph-captured training step (flat gradient bucket, capturable Adam) follows the eager
    `train_step` with torch.optim.Adam: same losses and parameters after several steps."""
    from soccerplayershapepose_b200.smpl import SMPL
    dev = torch.device("cuda", 0)
    feats, labels = _batch(synthetic_model, 6, 9)
    feats, labels = feats.to(dev), {k: v.to(dev) for k, v in labels.items()}
    smpl = SMPL(model_data=synthetic_model, mode="fp32").to(dev)
    g6, gproj = regressor.gpu_ops()
    head_a, crit_a = _make()
    head_a, crit_a = head_a.to(dev), crit_a.to(dev)
    head_b, crit_b = _make()
    head_b, crit_b = head_b.to(dev), crit_b.to(dev)
    head_b.load_state_dict(head_a.state_dict())
    crit_b.load_state_dict(crit_a.state_dict())
    lr, warm, n = 1e-3, 2, 3
    opt = torch.optim.Adam(list(head_a.parameters()) + list(crit_a.parameters()), lr=lr)
    eager = [regressor.train_step(head_a, crit_a, opt, smpl, feats, labels, g6, gproj).item() for _ in range(warm + n)]
    gs = regressor.GraphedTrainStep(head_b, crit_b, smpl, feats, labels, g6, gproj, lr=lr, warmup=warm)
    graphed = [gs(feats, labels).item() for _ in range(n)]
    np.testing.assert_allclose(graphed, eager[warm:], rtol=2e-4)
    for pa, pb in zip(head_a.parameters(), head_b.parameters()):
        assert (pa - pb).abs().max().item() < 5e-3 * lr + 1e-6 + 2e-3 * pa.abs().max().item() * 0 + 2e-5
