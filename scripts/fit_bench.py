#!/usr/bin/env python
"""BASELINE.json configs[2]: per-player SMPL fitting, Adam 200 iterations on the 2D keypoint reprojection loss
+ shape prior, 1024 synthetic players, one CUDA-graph-replayed batched loop.  Prints one JSON line."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from soccerplayershapepose_b200 import config                                   # noqa: E402
from soccerplayershapepose_b200.fitting import BatchedFitter                    # noqa: E402
from soccerplayershapepose_b200.model_io import make_synthetic_smpl             # noqa: E402
from soccerplayershapepose_b200.smpl import SMPL                                # noqa: E402
from soccerplayershapepose_b200.synthetic_inputs import make_smpl_inputs        # noqa: E402
from soccerplayershapepose_b200 import ops                                      # noqa: E402

from bench import ClockSampler                                                  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 200
mode = sys.argv[3] if len(sys.argv) > 3 else "fp32"
lr = float(sys.argv[4]) if len(sys.argv) > 4 else 1e-3        # the reference's fitting rate (Python/Soccer/global_var.py:71)
dev = torch.device("cuda", 0)
smpl = SMPL(model_data=make_synthetic_smpl(1234), mode=mode).to(dev)
x = make_smpl_inputs(B, 0)
g = torch.Generator().manual_seed(1)
rot_t, betas_t = x["rotmats"].to(dev), x["betas"].to(dev)
cam_t = torch.stack([0.6 + 0.6 * torch.rand(B, generator=g), 0.4 * torch.rand(B, generator=g) - 0.2,
                     0.4 * torch.rand(B, generator=g) - 0.2], 1).to(dev)
with torch.no_grad():
    out = smpl(betas=betas_t, body_pose=rot_t[:, 1:], global_orient=rot_t[:, :1], pose2rot=False, return_verts=False)
    label = ops.orthographic_project(out.joints, cam_t, 512.0)[:, config.SMPL_TO_KPRCNN_MAP, :].contiguous()
y = make_smpl_inputs(B, 7)                    # initial guess: another random draw
rot0, betas0 = y["rotmats"].to(dev), torch.zeros_like(betas_t)
cam0 = torch.tensor([0.9, 0.0, 0.0], device=dev).repeat(B, 1)
res = {}
for use_graph in (True, False):
    fitter = BatchedFitter(smpl, lr=lr, shape_weight=1e-3, use_cuda_graph=use_graph)
    fitter.fit(rot0, betas0, cam0, label, iterations=41)      # warm-up (captures the 20-iteration graph, kept per shape)
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    t0 = time.perf_counter()
    r = fitter.fit(rot0, betas0, cam0, label, iterations=iters)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    clocks = sampler.stop()
    if use_graph:
        clocks_graph = clocks
    res["graph" if use_graph else "eager"] = dt
    if use_graph:
        l0, l1 = r["initial_loss"].mean().item(), r["best_loss"].mean().item()
kernels = None
if os.environ.get("B200_FIT_KERNELS", "0") == "1":           # per-kernel device times of the eager iteration (library timers)
    import ctypes
    from soccerplayershapepose_b200 import _lib
    lib = _lib.load()
    lib.b200smpl_timing_enable(1)
    fitter.fit(rot0, betas0, cam0, label, iterations=10)
    torch.cuda.synchronize()
    lib.b200smpl_timing_enable(0)
    buf = ctypes.create_string_buffer(1 << 16)
    lib.b200smpl_timing_report(buf, 1 << 16)
    kernels = {r[0]: round(float(r[2]) / int(r[1]) * 1e3, 1) for r in (l.split() for l in buf.value.decode().splitlines())}
print(json.dumps({"workload": "BASELINE.json configs[2]: fit %d players x %d Adam iterations (lr %g), joints2D loss + shape "
                              "prior, %s mode, joints-only SMPL path" % (B, iters, lr, mode),
                  "clocks": clocks_graph,
                  "seconds_cuda_graph": res["graph"], "seconds_eager": res["eager"],
                  "player_iterations_per_s": B * iters / res["graph"], "ms_per_iteration": res["graph"] / iters * 1e3,
                  "mean_loss_initial": l0, "mean_loss_best": l1,
                  **({"kernel_us_eager_timed": kernels} if kernels else {})}))
