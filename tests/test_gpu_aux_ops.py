"""GPU tests of the small in-tree helpers (rot6d, projections, joints2D loss): forward against the
golden vectors produced by the reference's own files, backward against autograd of the oracle."""
import numpy as np
import pytest
import torch

from oracle import smpl_oracle as O
from soccerplayershapepose_b200 import cam_utils, config, joints2d_utils, ops, rigid_transform_utils

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda", 0)


def _rel(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


def test_rot6d_golden_and_grad(intree_golden, dev):
    g = intree_golden
    x = torch.from_numpy(g["rot6d_in"]).to(dev).requires_grad_(True)
    R = rigid_transform_utils.rot6d_to_rotmat(x)
    assert R.shape == (6 * 24, 3, 3)
    # fp32 Gram-Schmidt: same formula, different op order than torch's -> ulp-level differences that
    # the normalisations amplify for nearly-parallel input columns; compare against the fp64 oracle
    R64 = O.rot6d_to_rotmat(torch.from_numpy(g["rot6d_in"]).double())
    assert (R.detach().cpu().double() - R64).abs().max().item() < 1e-5
    assert np.abs(R.detach().cpu().numpy() - g["rot6d_out"]).max() < 1e-5
    w = torch.randn(R.shape, generator=torch.Generator().manual_seed(0))
    (R * w.to(dev)).sum().backward()
    x64 = torch.from_numpy(g["rot6d_in"]).double().requires_grad_(True)
    (O.rot6d_to_rotmat(x64) * w.double()).sum().backward()
    assert _rel(x.grad.cpu().double(), x64.grad) < 1e-5


def test_orthographic_golden_and_grad(intree_golden, dev):
    g = intree_golden
    pts = torch.from_numpy(g["ortho_points"]).to(dev).requires_grad_(True)
    cam = torch.from_numpy(g["ortho_cam"]).to(dev).requires_grad_(True)
    out = cam_utils.orthographic_project_torch(pts, cam)
    assert np.abs(out.detach().cpu().numpy() - g["ortho_out"]).max() < 1e-6
    px = joints2d_utils.undo_keypoint_normalisation(out, 512)
    assert np.abs(px.detach().cpu().numpy() - g["undo_norm_out"]).max() < 2e-4
    fused = ops.orthographic_project(pts, cam, 512.0)
    assert np.abs(fused.detach().cpu().numpy() - g["undo_norm_out"]).max() < 2e-4
    vis = joints2d_utils.check_joints2d_visibility_torch(px.detach(), 512)
    assert np.array_equal(vis.cpu().numpy(), g["vis_out"])
    w = torch.randn(out.shape, generator=torch.Generator().manual_seed(1))
    (fused * w.to(dev)).sum().backward()
    p64 = torch.from_numpy(g["ortho_points"]).double().requires_grad_(True)
    c64 = torch.from_numpy(g["ortho_cam"]).double().requires_grad_(True)
    (O.undo_keypoint_normalisation(O.orthographic_project(p64, c64), 512) * w.double()).sum().backward()
    assert _rel(pts.grad.cpu().double(), p64.grad) < 1e-5
    assert _rel(cam.grad.cpu().double(), c64.grad) < 1e-5


def test_weak_perspective_conversions(intree_golden, dev):
    g = intree_golden
    cam = torch.from_numpy(g["ortho_cam"]).to(dev)
    t = cam_utils.convert_weak_perspective_to_camera_translation_torch(cam, 5000.0, 512)
    np.testing.assert_allclose(t.cpu().numpy(), g["wp2t_out"], rtol=1e-6)
    wp = cam_utils.convert_camera_translation_to_weak_perspective_torch(t, 5000.0, 512)
    np.testing.assert_allclose(wp.cpu().numpy(), g["t2wp_out"], rtol=1e-6)
    assert np.array_equal(cam_utils.get_intrinsics_matrix(512, 512, 5000.0), g["intrinsics_512_5000"])


def test_perspective_golden_and_grad(intree_golden, dev):
    g = intree_golden
    pts = torch.from_numpy(g["ortho_points"]).to(dev).requires_grad_(True)
    rot = torch.from_numpy(g["persp_rot"]).to(dev).requires_grad_(True)
    tr = torch.from_numpy(g["persp_trans"]).to(dev).requires_grad_(True)
    out = cam_utils.perspective_project_torch(pts, rot, tr, focal_length=5000.0, img_wh=512)
    np.testing.assert_allclose(out.detach().cpu().numpy(), g["persp_out"], rtol=2e-6, atol=1e-3)
    w = torch.randn(out.shape, generator=torch.Generator().manual_seed(2))
    (out * w.to(dev)).sum().backward()
    p64, r64, t64 = (torch.from_numpy(g[k]).double().requires_grad_(True)
                     for k in ("ortho_points", "persp_rot", "persp_trans"))
    (O.perspective_project(p64, r64, t64, focal_length=5000.0, img_wh=512) * w.double()).sum().backward()
    assert _rel(pts.grad.cpu().double(), p64.grad) < 1e-5
    assert _rel(rot.grad.cpu().double(), r64.grad) < 1e-5
    assert _rel(tr.grad.cpu().double(), t64.grad) < 1e-5


def test_fused_joints2d_loss(dev):
    """player_recon.py:1217-1221 + losses/multi_task_loss.py:97-113 as one kernel."""
    B = 7
    gen = torch.Generator().manual_seed(5)
    joints = torch.randn(B, 90, 3, generator=gen) * 0.5
    cam = torch.stack([torch.rand(B, generator=gen) * 0.6 + 0.6, torch.rand(B, generator=gen) * 0.4 - 0.2,
                       torch.rand(B, generator=gen) * 0.4 - 0.2], 1)
    label = torch.rand(B, 17, 2, generator=gen) * 512
    vis = torch.rand(B, 17, generator=gen) > 0.3
    jmap = torch.tensor(config.SMPL_TO_KPRCNN_MAP, dtype=torch.int32)
    lv = O.init_log_var(1.0)
    for v in (None, vis):
        j = joints.clone().to(dev).requires_grad_(True)
        c = cam.clone().to(dev).requires_grad_(True)
        loss = ops.joints2d_loss(j, c, jmap.to(dev), label.to(dev), None if v is None else v.to(dev),
                                 proj_wh=512.0, norm_wh=256.0, log_var=lv)
        (3.0 * loss).backward()
        j64 = joints.double().requires_grad_(True)
        c64 = cam.double().requires_grad_(True)
        pred = O.undo_keypoint_normalisation(O.orthographic_project(j64, c64)[:, jmap.long()], 512.0)
        ref = O.joints2d_loss(pred, label.double(), torch.tensor(lv, dtype=torch.float64), 256.0, vis=v)
        (3.0 * ref).backward()
        assert abs(loss.item() - ref.item()) < 1e-5 * max(1.0, abs(ref.item()))
        assert _rel(j.grad.cpu().double(), j64.grad) < 1e-5
        assert _rel(c.grad.cpu().double(), c64.grad) < 1e-5


def test_ops_refuse_cpu_tensors():
    with pytest.raises(RuntimeError):
        ops.orthographic_project(torch.zeros(1, 4, 3), torch.zeros(1, 3))
    with pytest.raises(RuntimeError):
        rigid_transform_utils.rot6d_to_rotmat(torch.zeros(2, 6))


def test_batch_rodrigues_matches_smplx_arithmetic(dev):
    """smplx.lbs.batch_rodrigues as the reference calls it directly (player_recon.py:201,655): forward and
    gradient against the oracle restatement in float64, including the zero vector (angle = ||0 + 1e-8||) and
    rotations close to pi."""
    from soccerplayershapepose_b200.lbs import batch_rodrigues
    g = torch.Generator().manual_seed(3)
    r = torch.randn(500, 3, generator=g) * 0.8
    r[0] = 0.0
    r[1] = torch.tensor([3.1, 0.0, 0.0])
    r[2] = torch.tensor([0.0, -1e-4, 2e-4])
    x = r.clone().to(dev).requires_grad_(True)
    R = batch_rodrigues(x)
    assert R.shape == (500, 3, 3)
    x64 = r.double().requires_grad_(True)
    R64 = O.batch_rodrigues(x64)
    assert (R.detach().cpu().double() - R64.detach()).abs().max().item() < 2e-6
    w = torch.randn(500, 3, 3, generator=g)
    (R * w.to(dev)).sum().backward()
    (R64 * w.double()).sum().backward()
    # the zero vector sits on smplx's 1e-8 offset, where fp32 and fp64 legitimately differ: compare the others
    err = (x.grad.cpu().double()[1:] - x64.grad[1:]).abs().max().item() / x64.grad[1:].abs().max().item()
    assert err < 1e-4
    assert torch.isfinite(x.grad).all()
    # (B, 72) poses: same layout the reference reshapes from
    pose = torch.randn(4, 72, generator=g).to(dev)
    assert batch_rodrigues(pose.view(-1, 3)).view(4, 24, 3, 3).shape == (4, 24, 3, 3)


def test_perspective_with_explicit_intrinsics_golden_and_grad(intree_golden, dev):
    """General cam_K (utils/cam_utils.py:54-85 called with an intrinsics matrix): golden from the reference file,
    gradients against the fp64 oracle; a shared (3,3) matrix equals the batched call."""
    g = intree_golden
    pts = torch.from_numpy(g["ortho_points"]).to(dev).requires_grad_(True)
    rot = torch.from_numpy(g["persp_rot"]).to(dev).requires_grad_(True)
    tr = torch.from_numpy(g["persp_trans"]).to(dev).requires_grad_(True)
    K = torch.from_numpy(g["persp_camK"])
    out = cam_utils.perspective_project_torch(pts, rot, tr, cam_K=K.to(dev))
    np.testing.assert_allclose(out.detach().cpu().numpy(), g["persp_camK_out"], rtol=2e-6, atol=1e-3)
    w = torch.randn(out.shape, generator=torch.Generator().manual_seed(3))
    (out * w.to(dev)).sum().backward()
    p64, r64, t64 = (torch.from_numpy(g[k]).double().requires_grad_(True)
                     for k in ("ortho_points", "persp_rot", "persp_trans"))
    (O.perspective_project(p64, r64, t64, cam_K=K.double()) * w.double()).sum().backward()
    assert _rel(pts.grad.cpu().double(), p64.grad) < 1e-5
    assert _rel(rot.grad.cpu().double(), r64.grad) < 1e-5
    assert _rel(tr.grad.cpu().double(), t64.grad) < 1e-5
    one = cam_utils.perspective_project_torch(pts.detach(), rot.detach(), tr.detach(), cam_K=K[0].to(dev))
    rep = cam_utils.perspective_project_torch(pts.detach(), rot.detach(), tr.detach(), cam_K=K[:1].expand(K.shape[0], 3, 3).contiguous().to(dev))
    assert torch.equal(one, rep)
    with pytest.raises(ValueError):
        ops.perspective_project_camk(pts, rot, tr, K[:2].to(dev))


def test_fused_multitask_loss_against_the_reference_golden(intree_golden, dev):
    """csrc/loss.cu directly against HomoscedasticUncertaintyWeightedMultiTaskLoss executed from the reference file
    (tests/golden/make_golden_from_reference.py): the predicted pixels of the golden are fed as joints whose
    orthographic projection with cam = [1, 0, 0] lands exactly there (pix = (x + 1) * 256)."""
    g = intree_golden
    B = g["loss5_pred_verts"].shape[0]
    joints = torch.zeros(B, 34, 3)
    joints[:, :17, :2] = torch.from_numpy(g["loss5_pred_joints2D"]) / 256.0 - 1.0
    joints[:, 17:] = torch.from_numpy(g["loss5_pred_joints3D"])
    cam = torch.tensor([1.0, 0.0, 0.0]).repeat(B, 1)
    d = lambda k: torch.from_numpy(g[k]).to(dev)   # noqa: E731
    loss, parts = ops.multitask_loss(
        torch.from_numpy(g["loss5_log_vars"]).to(dev), verts=d("loss5_pred_verts"), verts_label=d("loss5_label_verts"),
        joints=joints.to(dev), cam=cam.to(dev), map2d=torch.arange(17, dtype=torch.int32, device=dev),
        label2d=d("loss5_label_joints2D"), vis=d("loss5_label_vis"), map3d=torch.arange(17, 34, dtype=torch.int32, device=dev),
        label3d=d("loss5_label_joints3D"), shape=d("loss5_pred_shape_params"), shape_label=d("loss5_label_shape_params"),
        pose=d("loss5_pred_pose_params_rot_matrices"), pose_label=d("loss5_label_pose_params_rot_matrices"),
        proj_wh=512.0, norm_wh=256.0)
    np.testing.assert_allclose(loss.item(), g["loss5_out"], rtol=2e-5)
    np.testing.assert_allclose(parts.cpu().numpy(), g["loss5_parts"], rtol=2e-5)
