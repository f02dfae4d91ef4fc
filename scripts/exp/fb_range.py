"""Debug helper: dense random dV on the vertex range [v0, v1): fused (fp32) vs unfused (fp32_simt) gradients."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from soccerplayershapepose_b200 import _lib
from soccerplayershapepose_b200.engine import SMPLEngine
from soccerplayershapepose_b200.model_io import make_synthetic_smpl
from soccerplayershapepose_b200.synthetic_inputs import make_smpl_inputs
B = int(sys.argv[1])
dev = torch.device("cuda", 0)
eng = SMPLEngine(make_synthetic_smpl(1234), dev)
x = make_smpl_inputs(B, 0)
b, r, t = (x[k].to(dev) for k in ("betas", "rotmats", "trans"))
g0 = torch.Generator().manual_seed(5)
full = torch.randn(B, 6890, 3, generator=g0).to(dev)
sets = [[2], [22], [0, 1, 2, 3], list(range(72))]
for pcs in sets:
    dV = torch.zeros_like(full)
    for pc in pcs:
        dV[:, 96 * pc:96 * pc + 96] = full[:, 96 * pc:96 * pc + 96]
    v0, v1 = pcs[0], pcs[-1]
    res = {}
    for mode in ("fp32", "fp32_simt"):
        m = _lib.MODES[mode]
        sv = eng.forward(b, r, t, None, mode=m, save=True)[3]
        g = eng.backward(b, r, t, None, None, dV, None, None, mode=m, saved=sv)
        res[mode] = [z.cpu() for z in g[:3]]
    errs = [(a - c).abs().max().item() / c.abs().max().item() for a, c in zip(res["fp32"], res["fp32_simt"])]
    print(pcs, "pieces [%d, %d]: betas %.2e pose %.2e transl %.2e" % (v0, v1, *errs))
    continue
    d = (res["fp32"][0] - res["fp32_simt"][0])
    print(" bodies with error:", (d.abs().max(1).values > 1e-3 * res["fp32_simt"][0].abs().max()).nonzero().flatten().tolist())
    print(" body 0 diff:", [round(float(z), 4) for z in d[0]], "ref:", [round(float(z), 4) for z in res["fp32_simt"][0][0]])
    dp = (res["fp32"][1] - res["fp32_simt"][1]).reshape(B, 24, 9)
    print(" pose diff per joint (body 0):", [round(float(z), 4) for z in dp[0].abs().max(1).values])
