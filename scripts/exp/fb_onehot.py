"""Debug helper: which vertices' gradient contributions the fused backward loses (one-hot dV, fp32 vs the SIMT mode's unfused path)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from soccerplayershapepose_b200 import _lib
from soccerplayershapepose_b200.engine import SMPLEngine
from soccerplayershapepose_b200.model_io import make_synthetic_smpl
from soccerplayershapepose_b200.synthetic_inputs import make_smpl_inputs
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda", 0)
eng = SMPLEngine(make_synthetic_smpl(1234), dev)
x = make_smpl_inputs(B, 0)
b, r, t = (x[k].to(dev) for k in ("betas", "rotmats", "trans"))
bad = []
for v in [2112, 2120, 2121, 2144, 2160, 2192, 2200, 200, 230, 260]:
    dV = torch.zeros(B, 6890, 3, device=dev)
    dV[:, v, :] = 1.0
    res = {}
    for mode in ("fp32", "fp32_simt"):
        m = _lib.MODES[mode]
        sv = eng.forward(b, r, t, None, mode=m, save=True)[3]
        g = eng.backward(b, r, t, None, None, dV, None, None, mode=m, saved=sv)
        res[mode] = g[0].cpu()
    err = (res["fp32"] - res["fp32_simt"]).abs().max().item() / res["fp32_simt"].abs().max().item()
    print(v, "err %.3f" % err, "fused", [round(float(z), 4) for z in res["fp32"][0][:6]], "ref", [round(float(z), 4) for z in res["fp32_simt"][0][:6]])
print("vertices with a wrong grad_betas:", [(v, round(e, 3)) for v, e in bad])
