#!/usr/bin/env python
"""Does a 4096-body step slow down when other (untouched / touched) allocations occupy HBM first?
footprint_probe.py [GB ...]  -- looks for the cause of the 10 % slower kernels at 32 768 bodies per GPU."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from soccerplayershapepose_b200 import _lib                                   # noqa: E402
from soccerplayershapepose_b200.engine import SMPLEngine                      # noqa: E402
from soccerplayershapepose_b200.model_io import make_synthetic_smpl           # noqa: E402
from soccerplayershapepose_b200.synthetic_inputs import make_smpl_inputs, make_upstream_grads  # noqa: E402

dev = torch.device("cuda", 0)
eng = SMPLEngine(make_synthetic_smpl(1234), dev)
m = _lib.MODES["fp32"]
B = 4096
x = make_smpl_inputs(B, 0)
dV0, dJ0 = make_upstream_grads(B, 0)


def run(tag):
    betas, rot, trans, dV, dJ = (t.to(dev) for t in (x["betas"], x["rotmats"], x["trans"], dV0, dJ0))
    def step():
        sv = eng.forward(betas, rot, trans, None, mode=m, save=True)[3]
        eng.backward(betas, rot, trans, None, None, dV, dJ, None, mode=m, saved=sv)
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        step()
    e1.record()
    torch.cuda.synchronize()
    print("%s: step %.4f ms, allocated %.1f GB" % (tag, e0.elapsed_time(e1) / 20, torch.cuda.memory_allocated() / 2**30), flush=True)


run("baseline")
hold = []
for gb in [float(a) for a in sys.argv[1:]] or [8.0, 16.0, 64.0]:
    hold.append(torch.empty(int(gb * 2**30), dtype=torch.uint8, device=dev).zero_())
    torch.cuda.synchronize()
    run("after +%g GB (touched once)" % gb)
