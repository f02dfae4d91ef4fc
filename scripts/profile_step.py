#!/usr/bin/env python
"""One warm-up + N forward+backward steps of the B=4096 workload, for ncu captures."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from soccerplayershapepose_b200 import _lib                                   # noqa: E402
from soccerplayershapepose_b200.engine import SMPLEngine                      # noqa: E402
from soccerplayershapepose_b200.model_io import make_synthetic_smpl           # noqa: E402
from soccerplayershapepose_b200.synthetic_inputs import make_smpl_inputs, make_upstream_grads  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
B = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
dev = torch.device("cuda", 0)
eng = SMPLEngine(make_synthetic_smpl(1234), dev)
x = make_smpl_inputs(B, 0)
dV, dJ = make_upstream_grads(B, 0)
betas, rot, trans, dV, dJ = (t.to(dev) for t in (x["betas"], x["rotmats"], x["trans"], dV, dJ))
m = _lib.MODES[mode]
for _ in range(steps):
    sv = eng.forward(betas, rot, trans, None, mode=m, save=True)[3]
    eng.backward(betas, rot, trans, None, None, dV, dJ, None, mode=m, saved=sv)
torch.cuda.synchronize()
print("ok")
