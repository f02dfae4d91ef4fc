import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from soccerplayershapepose_b200 import _lib
from soccerplayershapepose_b200.engine import SMPLEngine
from soccerplayershapepose_b200.model_io import make_synthetic_smpl
from soccerplayershapepose_b200.synthetic_inputs import make_smpl_inputs, make_upstream_grads
dev = torch.device("cuda", 0)
eng = SMPLEngine(make_synthetic_smpl(1234), dev)
B = 4096
x = make_smpl_inputs(B, 0)
dV, dJ = make_upstream_grads(B, 0)
betas, rot, trans, dV, dJ = (t.to(dev) for t in (x["betas"], x["rotmats"], x["trans"], dV, dJ))
for _ in range(2):
    sv = eng.forward(betas, rot, trans, None, mode=_lib.MODES["fp32"], save=True)[3]
    eng.backward(betas, rot, trans, None, None, dV, dJ, None, mode=_lib.MODES["fp32"], saved=sv)
    torch.cuda.synchronize()
    print("----", flush=True)
