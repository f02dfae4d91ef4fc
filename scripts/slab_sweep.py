#!/usr/bin/env python
"""Step time and per-kernel device times of forward+backward at batch B for several slab sizes, with the
forward products saved for the backward or recomputed by it: slab_sweep.py [mode] [B] [out.json]

The question this answers: does processing the batch in L2-sized slabs (blend GEMM ->
skinning back to back on 256..1024 bodies) beat one 4096-body pass whose intermediates cross HBM?"""
import ctypes, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from soccerplayershapepose_b200 import _lib                                   # noqa: E402
from soccerplayershapepose_b200.engine import SMPLEngine                      # noqa: E402
from soccerplayershapepose_b200.model_io import make_synthetic_smpl           # noqa: E402
from soccerplayershapepose_b200.synthetic_inputs import make_smpl_inputs, make_upstream_grads  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
out = sys.argv[3] if len(sys.argv) > 3 else None
dev = torch.device("cuda", 0)
eng = SMPLEngine(make_synthetic_smpl(1234), dev)
lib = _lib.load()
m = _lib.MODES[mode]
x = make_smpl_inputs(B, 0)
dV, dJ = make_upstream_grads(B, 0)
betas, rot, trans, dV, dJ = (t.to(dev) for t in (x["betas"], x["rotmats"], x["trans"], dV, dJ))
rows = []
for save in (True, False):
    for slab in (256, 512, 1024, 2048, 4096):
        def step():
            if save:
                sv = eng.forward(betas, rot, trans, None, mode=m, slab=slab, save=True)[3]
                eng.backward(betas, rot, trans, None, None, dV, dJ, None, mode=m, slab=slab, saved=sv)
            else:
                eng.forward(betas, rot, trans, None, mode=m, slab=slab)
                eng.backward(betas, rot, trans, None, None, dV, dJ, None, mode=m, slab=slab)
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
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        lib.b200smpl_timing_enable(0)
        buf = ctypes.create_string_buffer(1 << 16)
        lib.b200smpl_timing_report(buf, 1 << 16)
        per = {r[0]: float(r[2]) / 3 * 1e3 for r in (l.split() for l in buf.value.decode().splitlines())}
        rows.append({"mode": mode, "batch": B, "slab": slab, "saved": save, "ms_per_step": ms,
                     "kernel_us_per_step": per})
        print("slab=%d saved=%d step %.3f ms | " % (slab, save, ms) +
              " ".join("%s=%.0f" % (k, v) for k, v in sorted(per.items())), flush=True)
if out:
    json.dump(rows, open(out, "w"), indent=1)
