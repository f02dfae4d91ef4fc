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


def test_single_input_regressor_matches_the_reference_layout():
    """models/regressor.py:7-46 with resnet_layers=18: ResNet-18 encoder (11.2 M parameters) + IEFModule([512, 512],
    512, 157); 11.9 M parameters = the 47.6 MB gradient all-reduce of SURVEY.md section 8e."""
    net = regressor.SingleInputRegressor(resnet_in_channels=18)
    n = sum(p.numel() for p in net.parameters())
    assert n == 11_909_789
    keys = set(net.state_dict())
    for k in ("image_encoder.conv1.weight", "image_encoder.layer1.0.conv1.weight", "image_encoder.layer2.0.downsample.0.weight",
              "image_encoder.layer4.1.bn2.running_var", "ief_module.fc1.weight", "ief_module.ief_layers.4.bias"):
        assert k in keys, k
    assert not any(k.endswith("initial_params_estimate") or ".fc." in k for k in keys)
    assert net.image_encoder.conv1.weight.shape == (64, 18, 7, 7)
    cam, pose6d, shape = net(torch.zeros(2, 18, 64, 64))
    assert cam.shape == (2, 3) and pose6d.shape == (2, 144) and shape.shape == (2, 10)


@pytest.mark.gpu
@pytest.mark.parametrize("with_vis", [False, True])
@pytest.mark.parametrize("losses", [LOSSES, ("joints2D", "shape_params"), ("verts",)])
def test_fused_multitask_loss_matches_eager(synthetic_model, with_vis, losses):
    """csrc/loss.cu against the module's eager arithmetic (= losses/multi_task_loss.py:92-130, pinned to the reference
    file in tests/test_oracle.py): value, parts and every gradient (outputs and log-variances), float64 eager reference."""
    from soccerplayershapepose_b200.smpl import SMPL
    dev = torch.device("cuda", 0)
    B = 7
    feats, labels = _batch(synthetic_model, B, 11)
    if with_vis:
        labels["vis"] = torch.rand(B, 17, generator=torch.Generator().manual_seed(3)) > 0.3
    smpl = SMPL(model_data=synthetic_model, mode="fp32").to(dev)
    g6, gproj = regressor.gpu_ops()
    head, _ = _make()
    head = head.to(dev)
    crit = regressor.MultiTaskLoss(losses, WEIGHTS).to(dev)
    lab = {k: v.to(dev) for k, v in labels.items()}
    out = regressor.predict(head, smpl, feats.to(dev), g6, gproj)
    leaves = {k: out[k].detach().clone().requires_grad_(True) for k in ("verts", "joints", "cam", "shape_params",
                                                                        "pose_params_rot_matrices")}
    loss, parts = crit.forward_fused(lab, leaves)
    (2.5 * loss).backward()
    # eager reference in float64 on the same leaves
    crit64 = regressor.MultiTaskLoss(losses, WEIGHTS).double().to(dev)
    l64 = {k: v.detach().double().requires_grad_(True) for k, v in leaves.items()}
    o64 = dict(l64)
    o64["joints2D"] = O.undo_keypoint_normalisation(O.orthographic_project(l64["joints"], l64["cam"]), 512)[
        :, config.SMPL_TO_KPRCNN_MAP, :]
    o64["joints3D"] = l64["joints"][:, config.ALL_JOINTS_TO_COCO_MAP, :]
    lab64 = {k: (v.double() if v.dtype.is_floating_point else v) for k, v in lab.items()}
    ref, parts64 = crit64(lab64, o64)
    (2.5 * ref).backward()
    assert abs(loss.item() - ref.item()) < 1e-5 * max(1.0, abs(ref.item()))
    for name in losses:
        assert abs(parts[name].item() - parts64[name].item()) < 1e-5 * max(1e-3, abs(parts64[name].item()))
        g, g64 = getattr(crit, name + "_log_var").grad, getattr(crit64, name + "_log_var").grad
        assert abs(g.item() - g64.item()) < 1e-5 * max(1.0, abs(g64.item()))
    for k in leaves:
        g64 = l64[k].grad
        if g64 is None:
            assert leaves[k].grad is None or leaves[k].grad.abs().max().item() == 0.0, k
            continue
        err = (leaves[k].grad.double() - g64).abs().max().item()
        assert err < 2e-5 * g64.abs().max().item() + 1e-12, (k, err)


@pytest.mark.gpu
def test_graphed_step_with_fused_loss_and_encoder(synthetic_model):
    """The captured step with the fused loss kernels and the ResNet-18 encoder in front follows the eager step with
    the eager loss (same initial weights, same batches)."""
    from soccerplayershapepose_b200.smpl import SMPL
    dev = torch.device("cuda", 0)
    B = 4
    _, labels = _batch(synthetic_model, B, 13)
    labels = {k: v.to(dev) for k, v in labels.items()}
    crops = torch.randn(B, 18, 64, 64, generator=torch.Generator().manual_seed(1)).to(dev)
    smpl = SMPL(model_data=synthetic_model, mode="fp32").to(dev)
    g6, gproj = regressor.gpu_ops()
    torch.manual_seed(0)
    net_a = regressor.SingleInputRegressor(18).to(dev)
    net_b = regressor.SingleInputRegressor(18).to(dev)
    net_b.load_state_dict(net_a.state_dict())
    crit_a = regressor.MultiTaskLoss(LOSSES, WEIGHTS).to(dev)
    crit_b = regressor.MultiTaskLoss(LOSSES, WEIGHTS).to(dev)
    lr, warm, n = 1e-4, 2, 3
    opt = torch.optim.Adam(list(net_a.parameters()) + list(crit_a.parameters()), lr=lr)
    eager = [regressor.train_step(net_a, crit_a, opt, smpl, crops, labels, g6, gproj).item() for _ in range(warm + n)]
    gs = regressor.GraphedTrainStep(net_b, crit_b, smpl, crops, labels, g6, None, lr=lr, warmup=warm, fused_loss=True)
    graphed = [gs(crops, labels).item() for _ in range(n)]
    np.testing.assert_allclose(graphed, eager[warm:], rtol=2e-3)
