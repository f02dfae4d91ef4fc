#!/usr/bin/env python
"""Forward-only device time over the batch size (run with B200_FUSED_FWD=0 and 1 to compare the two paths)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from soccerplayershapepose_b200 import _lib                                   # noqa: E402
from soccerplayershapepose_b200.engine import SMPLEngine                      # noqa: E402
from soccerplayershapepose_b200.model_io import make_synthetic_smpl           # noqa: E402
from soccerplayershapepose_b200.synthetic_inputs import make_smpl_inputs      # noqa: E402
dev = torch.device("cuda", 0)
eng = SMPLEngine(make_synthetic_smpl(1234), dev)
m = _lib.MODES["fp32"]
out = []
for B in (1, 64, 256, 512, 1024, 2048, 4096, 8192):
    x = make_smpl_inputs(B, 0)
    b, r, t = (x[k].to(dev) for k in ("betas", "rotmats", "trans"))
    for _ in range(5):
        eng.forward(b, r, t, None, mode=m)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 30
    e0.record()
    for _ in range(n):
        eng.forward(b, r, t, None, mode=m)
    e1.record()
    torch.cuda.synchronize()
    out.append("B=%d %.4f ms" % (B, e0.elapsed_time(e1) / n))
print("B200_FUSED_FWD=%s: " % os.environ.get("B200_FUSED_FWD", "default") + " | ".join(out))
