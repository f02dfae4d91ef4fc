#!/usr/bin/env python
"""Bring-up diagnostic (run on the GPU box): per-mode error statistics of every stage against the
fp64 oracle, then per-kernel timings at B=4096.  Prints instead of asserting."""
import ctypes
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import smpl_oracle as O                                     # noqa: E402
from soccerplayershapepose_b200 import _lib                             # noqa: E402
from soccerplayershapepose_b200.engine import SMPLEngine                # noqa: E402
from soccerplayershapepose_b200.model_io import make_synthetic_smpl     # noqa: E402


def rel(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


def timing_report(lib):
    buf = ctypes.create_string_buffer(1 << 16)      # the call consumes the records: one shot
    lib.b200smpl_timing_report(buf, 1 << 16)
    return buf.value.decode()


def main():
    modes = sys.argv[1].split(",") if len(sys.argv) > 1 else ["fp32_simt", "fp32", "bf16"]
    dev = torch.device("cuda", 0)
    print(torch.cuda.get_device_name(0), flush=True)
    model = make_synthetic_smpl(1234)
    eng = SMPLEngine(model, dev)
    lib = _lib.load()
    orc = O.SMPLOracle(model, dtype=torch.float64)
    B = 70
    g = torch.Generator().manual_seed(0)
    betas = torch.randn(B, 10, generator=g)
    pose = torch.randn(B, 72, generator=g) * 0.3
    trans = torch.rand(B, 3, generator=g)
    cam = torch.rand(B, 3, generator=g) + 0.5
    dV = torch.randn(B, 6890, 3, generator=g)
    dJ = torch.randn(B, 90, 3, generator=g)
    dJ2 = torch.randn(B, 90, 2, generator=g)
    b64, p64, t64, c64 = (x.double().requires_grad_(True) for x in (betas, pose, trans, cam))
    ref = orc.forward_flat(b64, p64, t64, pose2rot=True)
    ((ref.vertices * dV.double()).sum() + (ref.joints * dJ.double()).sum()
     + (O.orthographic_project(ref.joints, c64) * dJ2.double()).sum()).backward()
    d = lambda x: x.to(dev)  # noqa: E731
    for mode in modes:
        m = _lib.MODES[mode]
        try:
            v, j, j2 = eng.forward(d(betas), d(pose), d(trans), d(cam), axis_angle=True, mode=m)
            torch.cuda.synchronize()
            ev = (v.cpu().double() - ref.vertices).abs()
            ej = (j.cpu().double() - ref.joints).abs()
            print("[%s] fwd: verts max %.3e (body %d vert %d) mean %.3e | joints max %.3e: 0-23 %.2e 24-44 %.2e 45+ %.2e"
                  % (mode, ev.max(), ev.amax((1, 2)).argmax(), ev[ev.amax((1, 2)).argmax()].amax(1).argmax(), ev.mean(),
                     ej.max(), ej[:, :24].max(), ej[:, 24:45].max(), ej[:, 45:].max()), flush=True)
            gb, gp, gt, gc = eng.backward(d(betas), d(pose), d(trans), d(cam), j, d(dV), d(dJ), d(dJ2),
                                          axis_angle=True, mode=m)
            torch.cuda.synchronize()
            print("[%s] bwd rel err: betas %.3e pose %.3e transl %.3e cam %.3e"
                  % (mode, rel(gb.cpu().double(), b64.grad), rel(gp.cpu().double(), p64.grad),
                     rel(gt.cpu().double(), t64.grad), rel(gc.cpu().double(), c64.grad)), flush=True)
        except Exception as e:  # noqa: BLE001
            print("[%s] FAILED: %r" % (mode, e), flush=True)
            return 1
    # timings
    B = 4096
    g = torch.Generator().manual_seed(1)
    betas = torch.randn(B, 10, generator=g).to(dev)
    pose = (torch.randn(B, 72, generator=g) * 0.3).to(dev)
    rot = O.batch_rodrigues(pose.cpu().reshape(-1, 3)).reshape(B, 24, 3, 3).to(dev)
    trans = torch.rand(B, 3, generator=g).to(dev)
    dV = torch.randn(B, 6890, 3, device=dev)
    dJ = torch.randn(B, 90, 3, device=dev)
    for mode in [x for x in modes if x != "fp32_simt"]:
        m = _lib.MODES[mode]
        for slab in (1024, 4096):
            for _ in range(2):
                sv = eng.forward(betas, rot, trans, None, mode=m, slab=slab, save=True)[3]
                eng.backward(betas, rot, trans, None, None, dV, dJ, None, mode=m, slab=slab, saved=sv)
            torch.cuda.synchronize()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            n = 5
            e0.record()
            for _ in range(n):
                sv = eng.forward(betas, rot, trans, None, mode=m, slab=slab, save=True)[3]
            e1.record()
            for _ in range(n):
                eng.backward(betas, rot, trans, None, None, dV, dJ, None, mode=m, slab=slab, saved=sv)
            e2.record()
            torch.cuda.synchronize()
            tf, tb = e0.elapsed_time(e1) / n, e1.elapsed_time(e2) / n
            print("[%s slab %d] B=%d fwd %.3f ms bwd %.3f ms -> %.2f M meshes/s fwd+bwd"
                  % (mode, slab, B, tf, tb, B / (tf + tb) / 1e3), flush=True)
        lib.b200smpl_timing_enable(1)
        for _ in range(3):
            sv = eng.forward(betas, rot, trans, None, mode=m, slab=4096, save=True)[3]
            eng.backward(betas, rot, trans, None, None, dV, dJ, None, mode=m, slab=4096, saved=sv)
        torch.cuda.synchronize()
        lib.b200smpl_timing_enable(0)
        print("[%s] per-kernel (3 steps, slab 4096): name launches total_ms\n%s" % (mode, timing_report(lib)), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
