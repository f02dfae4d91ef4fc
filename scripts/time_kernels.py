#!/usr/bin/env python
"""Per-kernel device times (library launch timers) of forward+backward at batch B: time_kernels.py [mode] [B ...]"""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from soccerplayershapepose_b200 import _lib                                   # noqa: E402
from soccerplayershapepose_b200.engine import SMPLEngine                      # noqa: E402
from soccerplayershapepose_b200.model_io import make_synthetic_smpl           # noqa: E402
from soccerplayershapepose_b200.synthetic_inputs import make_smpl_inputs, make_upstream_grads  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
Bs = [int(x) for x in sys.argv[2:]] or [4096]
dev = torch.device("cuda", 0)
eng = SMPLEngine(make_synthetic_smpl(1234), dev)
lib = _lib.load()
m = _lib.MODES[mode]
for B in Bs:
    x = make_smpl_inputs(B, 0)
    dV, dJ = make_upstream_grads(B, 0)
    betas, rot, trans, dV, dJ = (t.to(dev) for t in (x["betas"], x["rotmats"], x["trans"], dV, dJ))
    def step():
        sv = eng.forward(betas, rot, trans, None, mode=m, save=True)[3]
        eng.backward(betas, rot, trans, None, None, dV, dJ, None, mode=m, saved=sv)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    lib.b200smpl_timing_enable(1)
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    lib.b200smpl_timing_enable(0)
    buf = ctypes.create_string_buffer(1 << 16)
    lib.b200smpl_timing_report(buf, 1 << 16)
    rows = [l.split() for l in buf.value.decode().splitlines()]
    per = {r[0]: float(r[2]) / int(r[1]) * 1e3 for r in rows}
    print("B=%d step %.3f ms (%.2f M meshes/s) | " % (B, ms, B / ms / 1e3) + " ".join("%s=%.0fus" % (k, v) for k, v in sorted(per.items())), flush=True)
