#!/usr/bin/env python
"""Gradient error against the fp64 oracle and step time for every (forward mode, backward mode) pair, on the compact
and the wide synthetic model: mode_matrix.py [out.json].  Decides what the "bf16-GEMM mode" backward has to be to
meet the north_star's 1e-4 relative gradient bound."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import smpl_oracle as O                                            # noqa: E402
from soccerplayershapepose_b200 import _lib                                   # noqa: E402
from soccerplayershapepose_b200.engine import SMPLEngine                      # noqa: E402
from soccerplayershapepose_b200.model_io import make_synthetic_smpl           # noqa: E402
from soccerplayershapepose_b200.synthetic_inputs import make_smpl_inputs, make_upstream_grads  # noqa: E402

dev = torch.device("cuda", 0)
rows = []
for stats in ("compact", "wide"):
    model = make_synthetic_smpl(1234, statistics=stats)
    eng = SMPLEngine(model, dev)
    orc = O.SMPLOracle(model, dtype=torch.float64)
    B = 64
    x = make_smpl_inputs(B, 3)
    dV, dJ = make_upstream_grads(B, 3)
    b64, r64, t64 = (x[k].double().requires_grad_(True) for k in ("betas", "rotmats", "trans"))
    ref = orc.forward_flat(b64, r64, t64, pose2rot=False)
    ((ref.vertices * dV.double()).sum() + (ref.joints * dJ.double()).sum()).backward()
    want = (b64.grad, r64.grad.reshape(B, -1), t64.grad)
    betas, rot, trans, dVd, dJd = (t.to(dev) for t in (x["betas"], x["rotmats"], x["trans"], dV, dJ))
    B2 = 4096
    y = make_smpl_inputs(B2, 0)
    dV2, dJ2 = make_upstream_grads(B2, 0)
    betas2, rot2, trans2, dV2, dJ2 = (t.to(dev) for t in (y["betas"], y["rotmats"], y["trans"], dV2, dJ2))
    for fm in ("fp32", "bf16"):
        for bm in ("fp32", "bf16"):
            v, j, _, sv = eng.forward(betas, rot, trans, None, mode=_lib.MODES[fm], save=True)
            ev = (v.cpu().double() - ref.vertices.detach()).abs().max().item()
            got = eng.backward(betas, rot, trans, None, None, dVd, dJd, None, mode=_lib.MODES[bm], saved=sv)
            errs = [((g.cpu().double().reshape(w.shape) - w).abs().max() / w.abs().max()).item() for g, w in zip(got[:3], want)]

            def step():
                s = eng.forward(betas2, rot2, trans2, None, mode=_lib.MODES[fm], save=True)[3]
                eng.backward(betas2, rot2, trans2, None, None, dV2, dJ2, None, mode=_lib.MODES[bm], saved=s)
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            rows.append({"model": stats, "virtual_groups": int(eng.info.num_virtual_groups), "fwd_mode": fm, "bwd_mode": bm,
                         "vertex_err_m": ev, "grad_rel_err_betas_pose_transl": errs, "ms_per_step_b4096": ms})
            print(rows[-1], flush=True)
if len(sys.argv) > 1:
    json.dump(rows, open(sys.argv[1], "w"), indent=1)
