"""GPU tests of the batched fitting loop against a restatement of the reference's
single_view_optimization (player_recon.py:1172-1294) on the CPU oracle with torch.optim.Adam, float64."""
import numpy as np
import pytest
import torch

from oracle import smpl_oracle as O
from soccerplayershapepose_b200 import config
from soccerplayershapepose_b200.fitting import BatchedFitter, FROZEN_FULL_JOINTS
from soccerplayershapepose_b200.smpl import SMPL

pytestmark = pytest.mark.gpu


def _problem(B, seed=0):
    g = torch.Generator().manual_seed(seed)
    betas_t = torch.randn(B, 10, generator=g) * 0.8
    pose_t = torch.randn(B, 72, generator=g) * 0.25
    cam_t = torch.stack([0.6 + 0.6 * torch.rand(B, generator=g), 0.4 * torch.rand(B, generator=g) - 0.2,
                         0.4 * torch.rand(B, generator=g) - 0.2], 1)
    rot_t = O.batch_rodrigues(pose_t.reshape(-1, 3)).reshape(B, 24, 3, 3)
    # initial guess: perturbed truth
    betas0 = betas_t + 0.5 * torch.randn(B, 10, generator=g)
    pose0 = pose_t + 0.15 * torch.randn(B, 72, generator=g)
    cam0 = cam_t + 0.05 * torch.randn(B, 3, generator=g)
    rot0 = O.batch_rodrigues(pose0.reshape(-1, 3)).reshape(B, 24, 3, 3)
    return betas_t, rot_t, cam_t, betas0, rot0, cam0


def _ref_fit(model, rot0, betas0, cam0, label, iters, lr, shape_w):
    """The reference loop restated (float64, CPU): Adam over raw rotation matrices / betas / cam, hands and feet
    frozen, per-player joints2D loss (+ shape term), best iterate by loss."""
    orc = O.SMPLOracle(model, dtype=torch.float64)
    rot = rot0.double().clone().requires_grad_(True)
    betas = betas0.double().clone().requires_grad_(True)
    cam = cam0.double().clone().requires_grad_(True)
    opt = torch.optim.Adam([rot, betas, cam], lr=lr)
    B = rot.shape[0]
    best_loss = torch.full((B,), float("inf"), dtype=torch.float64)
    best = [rot.detach().clone(), betas.detach().clone(), cam.detach().clone()]
    best_iter = torch.zeros(B, dtype=torch.int64)
    losses = []
    for it in range(1, iters + 1):
        out = orc.forward_flat(betas, rot, None, pose2rot=False)
        j2d = O.orthographic_project(out.joints, cam)[:, config.SMPL_TO_KPRCNN_MAP, :]
        px = O.undo_keypoint_normalisation(j2d, 512)
        d = (2.0 * px / config.REGRESSOR_IMG_WH - 1.0) - (2.0 * label.double() / config.REGRESSOR_IMG_WH - 1.0)
        loss_b = (d * d).mean(dim=(1, 2)) + shape_w * (betas * betas).mean(1)
        losses.append(loss_b.detach().clone())
        imp = loss_b.detach() < best_loss
        best_loss = torch.where(imp, loss_b.detach(), best_loss)
        best_iter[imp] = it
        for cur, bst in zip((rot, betas, cam), best):
            bst[imp] = cur.detach()[imp]
        opt.zero_grad()
        loss_b.sum().backward()
        rot.grad[:, list(FROZEN_FULL_JOINTS)] = 0.0
        opt.step()
    return dict(rot=rot.detach(), betas=betas.detach(), cam=cam.detach(), best_loss=best_loss, best_iter=best_iter,
                best=best, losses=torch.stack(losses))


@pytest.fixture(scope="module")
def setup(synthetic_model):
    dev = torch.device("cuda", 0)
    smpl = SMPL(model_data=synthetic_model, mode="fp32").to(dev)
    return synthetic_model, smpl, dev


def _labels(model, rot_t, betas_t, cam_t):
    orc = O.SMPLOracle(model, dtype=torch.float64)
    out = orc.forward_flat(betas_t.double(), rot_t.double(), None, pose2rot=False)
    j2d = O.orthographic_project(out.joints, cam_t.double())[:, config.SMPL_TO_KPRCNN_MAP, :]
    return O.undo_keypoint_normalisation(j2d, 512).float()


@pytest.mark.parametrize("graph", [False, True])
def test_fit_matches_reference_loop(setup, graph):
    model, smpl, dev = setup
    B, iters, lr, sw = 6, 10, 2e-3, 0.05
    betas_t, rot_t, cam_t, betas0, rot0, cam0 = _problem(B)
    label = _labels(model, rot_t, betas_t, cam_t)
    ref = _ref_fit(model, rot0, betas0, cam0, label, iters, lr, sw)
    fitter = BatchedFitter(smpl, lr=lr, shape_weight=sw, use_cuda_graph=graph, graph_iterations=4)
    res = fitter.fit(rot0.to(dev), betas0.to(dev), cam0.to(dev), label.to(dev), iterations=iters)
    torch.cuda.synchronize()
    # loss of the first and of the last iteration, per player
    np.testing.assert_allclose(res["initial_loss"].cpu().double().numpy(), ref["losses"][0].numpy(), rtol=2e-4)
    np.testing.assert_allclose(res["last_loss"].cpu().double().numpy(), ref["losses"][-1].numpy(), rtol=2e-3)
    np.testing.assert_allclose(res["best_loss"].cpu().double().numpy(), ref["best_loss"].numpy(), rtol=2e-3)
    assert torch.equal(res["best_iter"].cpu().long(), ref["best_iter"])
    # parameters after `iters` Adam steps (Adam normalises the gradient: entries with a vanishing gradient are
    # sensitive to rounding, hence the robust statistic next to the max)
    for got, want in ((res["final_rotmats"], ref["rot"]), (res["final_betas"], ref["betas"]), (res["final_cam"], ref["cam"])):
        diff = (got.cpu().double() - want).abs()
        assert diff.median().item() < 1e-5 and diff.max().item() < 2 * lr * iters
        assert torch.quantile(diff.flatten(), 0.99).item() < 2e-3
    # frozen joints never move
    assert torch.equal(res["final_rotmats"][:, list(FROZEN_FULL_JOINTS)].cpu(), rot0[:, list(FROZEN_FULL_JOINTS)])
    # the loss goes down
    assert (res["best_loss"] < res["initial_loss"]).all()
    # a second call of the same shape reuses the state buffers (and the cached graph: replays only)
    res2 = fitter.fit(rot0.to(dev), betas0.to(dev), cam0.to(dev), label.to(dev), iterations=iters)
    torch.cuda.synchronize()
    np.testing.assert_allclose(res2["initial_loss"].cpu().numpy(), res["initial_loss"].cpu().numpy(), rtol=1e-5)
    np.testing.assert_allclose(res2["best_loss"].cpu().numpy(), res["best_loss"].cpu().numpy(), rtol=1e-3)
    assert torch.equal(res2["best_iter"], res["best_iter"])
    assert (res2["final_betas"] - res["final_betas"]).abs().max().item() < 1e-3


def test_fit_result_keys_and_translation(setup):
    model, smpl, dev = setup
    B = 3
    betas_t, rot_t, cam_t, betas0, rot0, cam0 = _problem(B, seed=3)
    label = _labels(model, rot_t, betas_t, cam_t)
    vis = torch.ones(B, 17, dtype=torch.bool)
    vis[0, 3] = False
    res = BatchedFitter(smpl, lr=1e-3).fit(rot0.to(dev), betas0.to(dev), cam0.to(dev), label.to(dev), vis=vis.to(dev),
                                           iterations=5)
    # the reference's .npz keys (player_recon.py:1293-1294)
    assert res["body_pose"].shape == (B, 23, 3, 3) and res["global_orient"].shape == (B, 1, 3, 3)
    assert res["betas"].shape == (B, 10) and res["translation"].shape == (B, 3)
    t = O.weak_perspective_to_translation(res["cam"].cpu().double(), config.FOCAL_LENGTH, 512)
    np.testing.assert_allclose(res["translation"].cpu().double().numpy(), t.numpy(), rtol=1e-5)


# ---------------------------------------------------------------------------------------------
# multi-view optimisation (player_recon.py:1568-1999)
# ---------------------------------------------------------------------------------------------
def _ref_multiview(model, bp0, betas0, go0, cam0, label, iters, lr, rounds, orders):
    """The reference's alternating loop restated on the oracle (float64, CPU), joints2D loss only: stacked per-view
    tensors under ONE Adam (so a step on one view moves the others by momentum), fresh optimisers per phase, a
    validation pass per epoch, per-phase restore to the best epoch, global best kept for the result."""
    from soccerplayershapepose_b200.fitting import MultiViewFitter  # noqa: F401  (same module is under test)
    orc = O.SMPLOracle(model, dtype=torch.float64)
    P, V = go0.shape[0], go0.shape[1]
    go = go0.double().transpose(0, 1).contiguous().clone()          # (V,P,3,3)
    cam = cam0.double().transpose(0, 1).contiguous().clone()        # (V,P,3)
    bp = bp0.double().clone()
    betas = betas0.double().clone()
    lab = label.double().transpose(0, 1)
    frozen = [6, 7, 21, 22]

    def loss_view(v):
        rot = torch.cat([go[v][:, None], bp], 1)
        out = orc.forward_flat(betas, rot, None, pose2rot=False)
        j2d = O.orthographic_project(out.joints, cam[v])[:, config.SMPL_TO_KPRCNN_MAP, :]
        px = O.undo_keypoint_normalisation(j2d, 512)
        d = (2.0 * px / config.REGRESSOR_IMG_WH - 1.0) - (2.0 * lab[v] / config.REGRESSOR_IMG_WH - 1.0)
        return (d * d).mean(dim=(1, 2))

    best_metric = torch.full((P,), float("inf"), dtype=torch.float64)
    best = {"go": go.clone(), "cam": cam.clone(), "bp": bp.clone(), "betas": betas.clone()}
    final = {k: t.clone() for k, t in best.items()}
    e = 0
    first_val = None
    for _ in range(rounds):
        for phase in ("A", "B"):
            params = {"A": {"cam": cam, "go": go}, "B": {"bp": bp, "betas": betas}}[phase]
            for t in params.values():
                t.requires_grad_(True)
            opt = torch.optim.Adam(list(params.values()), lr=lr)
            for _e in range(iters):
                for v in orders[e]:
                    opt.zero_grad()
                    loss_view(v).sum().backward()
                    if phase == "B":
                        bp.grad[:, frozen] = 0.0
                    opt.step()
                e += 1
                with torch.no_grad():
                    val = sum(loss_view(v) for v in range(V))
                    if first_val is None:
                        first_val = val.clone()
                    imp = val < best_metric
                    best_metric = torch.where(imp, val, best_metric)
                    cur = {"go": go, "cam": cam, "bp": bp, "betas": betas}
                    for k in cur:
                        dst = [final[k]] + ([best[k]] if k in params else [])
                        for d_ in dst:
                            if k in ("go", "cam"):
                                d_[:, imp] = cur[k].detach()[:, imp]
                            else:
                                d_[imp] = cur[k].detach()[imp]
            for t in params.values():
                t.requires_grad_(False)
            with torch.no_grad():
                for k in params:
                    params[k].copy_(best[k])
    return {"final": final, "best_metric": best_metric, "first_val": first_val,
            "last": {"go": go, "cam": cam, "bp": bp, "betas": betas}}


@pytest.mark.parametrize("use_graph", [True, False])
def test_multi_view_fit_matches_reference_loop(setup, use_graph):
    """use_graph: every (phase, view) step and both validation passes run eagerly once, are captured on their second
    use and replayed from then on -- all three paths inside one fit; a second fit on the cached graphs must agree."""
    from soccerplayershapepose_b200.fitting import MultiViewFitter
    model, smpl, dev = setup
    P, V, iters, rounds, lr = 4, 3, 3, 2, 2e-3
    g = torch.Generator().manual_seed(5)
    betas_t = torch.randn(P, 10, generator=g) * 0.8
    bp_aa = torch.randn(P, 69, generator=g) * 0.25
    go_aa = torch.randn(P, V, 3, generator=g) * 0.6
    cam_t = torch.stack([0.6 + 0.6 * torch.rand(P, V, generator=g), 0.4 * torch.rand(P, V, generator=g) - 0.2,
                         0.4 * torch.rand(P, V, generator=g) - 0.2], -1)
    bp_t = O.batch_rodrigues(bp_aa.reshape(-1, 3)).reshape(P, 23, 3, 3)
    go_t = O.batch_rodrigues(go_aa.reshape(-1, 3)).reshape(P, V, 3, 3)
    orc = O.SMPLOracle(model, dtype=torch.float64)
    label = torch.empty(P, V, 17, 2)
    for v in range(V):
        out = orc.forward_flat(betas_t.double(), torch.cat([go_t[:, v:v + 1], bp_t], 1).double(), None, pose2rot=False)
        j2d = O.orthographic_project(out.joints, cam_t[:, v].double())[:, config.SMPL_TO_KPRCNN_MAP, :]
        label[:, v] = O.undo_keypoint_normalisation(j2d, 512).float()
    betas0 = betas_t + 0.5 * torch.randn(P, 10, generator=g)
    bp0 = O.batch_rodrigues((bp_aa + 0.15 * torch.randn(P, 69, generator=g)).reshape(-1, 3)).reshape(P, 23, 3, 3)
    go0 = O.batch_rodrigues((go_aa + 0.1 * torch.randn(P, V, 3, generator=g)).reshape(-1, 3)).reshape(P, V, 3, 3)
    cam0 = cam_t + 0.05 * torch.randn(P, V, 3, generator=g)
    rng = np.random.default_rng(0)
    orders = [rng.permutation(V).tolist() for _ in range(rounds * 2 * iters)]
    ref = _ref_multiview(model, bp0, betas0, go0, cam0, label, iters, lr, rounds, orders)
    fitter = MultiViewFitter(smpl, lr=lr, rounds=rounds, use_cuda_graph=use_graph)
    res = fitter.fit(bp0.to(dev), betas0.to(dev), go0.to(dev), cam0.to(dev), label.to(dev), iterations=iters,
                     view_orders=orders)
    torch.cuda.synchronize()
    if use_graph:                                           # replays only, from re-initialised state buffers
        assert len(fitter._cache[(P, V, 10, False)]["graphs"]) == 2 * V + 2
        again = fitter.fit(bp0.to(dev), betas0.to(dev), go0.to(dev), cam0.to(dev), label.to(dev), iterations=iters,
                           view_orders=orders)
        for k in ("body_pose", "betas", "global_orient", "cam", "best_loss"):
            assert (again[k] - res[k]).abs().max().item() < 1e-5, k
    np.testing.assert_allclose(res["initial_loss"].cpu().double().numpy(), ref["first_val"].numpy(), rtol=2e-3)
    np.testing.assert_allclose(res["best_loss"].cpu().double().numpy(), ref["best_metric"].numpy(), rtol=5e-3)
    nsteps = rounds * iters * V
    pairs = ((res["body_pose"].reshape(P, 207), ref["final"]["bp"].reshape(P, 207)),
             (res["betas"], ref["final"]["betas"]),
             (res["global_orient"].reshape(P, V, 9), ref["final"]["go"].transpose(0, 1).reshape(P, V, 9)),
             (res["cam"], ref["final"]["cam"].transpose(0, 1)),
             (res["last"]["body_pose"].reshape(P, 207), ref["last"]["bp"].reshape(P, 207)),
             (res["last"]["global_orient"].reshape(P, V, 9), ref["last"]["go"].transpose(0, 1).reshape(P, V, 9)))
    for got, want in pairs:
        diff = (got.cpu().double() - want.detach()).abs()
        assert diff.median().item() < 2e-5 and diff.max().item() < 2 * lr * nsteps
        assert torch.quantile(diff.flatten(), 0.99).item() < 3e-3
    # hands and feet of the shared body pose never move; the result carries the reference's keys
    fz = [6, 7, 21, 22]
    assert torch.equal(res["last"]["body_pose"][:, fz].cpu(), bp0[:, fz])
    assert res["translation"].shape == (P, V, 3) and res["global_orient"].shape == (P, V, 3, 3)
    assert (res["best_loss"] < res["initial_loss"]).all()


@pytest.mark.parametrize("cols,batch", [((216, 10, 3), 37), ((72, 10, 3), 50), ((3,), 100), ((300, 7), 9), ((1, 1, 1, 1), 65)])
def test_fit_update_matches_torch_adam(cols, batch):
    """b200smpl_fit_update alone (best-iterate copy + Adam of up to four tensors in one launch) against
    torch.optim.Adam over three steps: several bodies per CTA (short rows), rows longer than a CTA, a ragged last
    CTA, frozen columns and an extra gradient term."""
    import ctypes
    from soccerplayershapepose_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(len(cols) * 1000 + batch)
    n = len(cols)
    p = [torch.randn(batch, c, generator=g).to(dev) for c in cols]
    frozen = [None if k % 2 else (torch.rand(c, generator=g) < 0.3).to(torch.uint8).to(dev) for k, c in enumerate(cols)]
    ref = [x.double().cpu().clone().requires_grad_(True) for x in p]
    opt = torch.optim.Adam(ref, lr=1e-2)
    m = [torch.zeros_like(x) for x in p]
    v = [torch.zeros_like(x) for x in p]
    best = [torch.zeros_like(x) for x in p]
    best_loss = torch.full((batch,), float("inf"), device=dev)
    best_iter = torch.zeros(batch, dtype=torch.int32, device=dev)
    first_loss = torch.zeros(batch, device=dev)
    step = torch.zeros(2, dtype=torch.int32, device=dev)
    want_best = [x.double().cpu().clone() for x in p]
    want_best_loss = torch.full((batch,), float("inf"), dtype=torch.float64)
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    for it in range(3):
        grads = [torch.randn(batch, c, generator=g) for c in cols]
        extra = [torch.randn(batch, c, generator=g) if k == 1 else None for k, c in enumerate(cols)]
        loss = torch.rand(batch, generator=g) + (0.0 if it != 1 else 0.5)     # some bodies improve, some do not
        gd = [x.to(dev) for x in grads]
        ed = [None if x is None else x.to(dev) for x in extra]
        groups = (_lib.FitGroup * n)()
        for k in range(n):
            groups[k] = _lib.FitGroup(p[k].data_ptr(), gd[k].data_ptr(), None if ed[k] is None else ed[k].data_ptr(),
                                      m[k].data_ptr(), v[k].data_ptr(), best[k].data_ptr(),
                                      None if frozen[k] is None else frozen[k].data_ptr(), cols[k])
        _lib.check(lib.b200smpl_fit_update(ctypes.cast(groups, ctypes.c_void_p), n, loss.to(dev).data_ptr(),
                                           best_loss.data_ptr(), best_iter.data_ptr(), first_loss.data_ptr(),
                                           step.data_ptr(), it & 1, batch, 1e-2, 0.9, 0.999, 1e-8, stream), "fit_update")
        torch.cuda.synchronize()
        # reference: keep the parameters that produced this loss where it improved, then one Adam step
        imp = loss.double() < want_best_loss
        want_best_loss = torch.where(imp, loss.double(), want_best_loss)
        for k in range(n):
            want_best[k][imp] = ref[k].detach()[imp]
            gr = grads[k].double() + (0.0 if extra[k] is None else extra[k].double())
            if frozen[k] is not None:
                gr[:, frozen[k].cpu().bool()] = 0.0
            ref[k].grad = gr
        opt.step()
        if it == 0:
            assert torch.equal(first_loss.cpu(), loss)
    assert int(step[1].item()) == 3                       # written to step[1 - parity] of the last call (parity 0)
    assert torch.allclose(best_loss.cpu().double(), want_best_loss, atol=1e-7)
    for k in range(n):
        assert torch.allclose(p[k].cpu().double(), ref[k].detach(), atol=2e-6), (k, (p[k].cpu().double() - ref[k].detach()).abs().max())
        assert torch.allclose(best[k].cpu().double(), want_best[k], atol=2e-6), k
