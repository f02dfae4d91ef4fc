"""Debug helper: gradients of one backward call saved to a file (run with B200_FUSED_BWD=0 / 1 and compare)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from soccerplayershapepose_b200 import _lib
from soccerplayershapepose_b200.engine import SMPLEngine
from soccerplayershapepose_b200.model_io import make_synthetic_smpl
from soccerplayershapepose_b200.synthetic_inputs import make_smpl_inputs, make_upstream_grads
out, B, use_j = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
mode = sys.argv[4] if len(sys.argv) > 4 else "fp32"
dev = torch.device("cuda", 0)
eng = SMPLEngine(make_synthetic_smpl(1234), dev)
x = make_smpl_inputs(B, 0)
dV, dJ = make_upstream_grads(B, 0)
b, r, t, dV, dJ = (z.to(dev) for z in (x["betas"], x["rotmats"], x["trans"], dV, dJ))
m = _lib.MODES[mode]
sv = eng.forward(b, r, t, None, mode=m, save=True)[3]
g = eng.backward(b, r, t, None, None, dV, dJ if use_j else None, None, mode=m, saved=sv)
torch.cuda.synchronize()
torch.save([z.cpu() for z in g[:3]], out)
if os.path.exists(out + ".ref"):
    ref = torch.load(out + ".ref")
    for name, a, c in zip(("betas", "pose", "transl"), g[:3], ref):
        a = a.cpu()
        print("%s: max rel diff %.3e" % (name, (a - c).abs().max().item() / c.abs().max().item()))
