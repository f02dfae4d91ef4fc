#!/usr/bin/env python
"""A few eager fitting iterations (no CUDA graph) for an ncu launch list / full capture: fit_iter_profile.py [B] [iters]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from soccerplayershapepose_b200 import config, ops                              # noqa: E402
from soccerplayershapepose_b200.fitting import BatchedFitter                    # noqa: E402
from soccerplayershapepose_b200.model_io import make_synthetic_smpl             # noqa: E402
from soccerplayershapepose_b200.smpl import SMPL                                # noqa: E402
from soccerplayershapepose_b200.synthetic_inputs import make_smpl_inputs        # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda", 0)
smpl = SMPL(model_data=make_synthetic_smpl(1234), mode="fp32").to(dev)
x = make_smpl_inputs(B, 0)
rot, betas = x["rotmats"].to(dev), x["betas"].to(dev)
cam = torch.tensor([0.9, 0.0, 0.0], device=dev).repeat(B, 1)
label = torch.rand(B, len(config.SMPL_TO_KPRCNN_MAP), 2, device=dev) * 512
fitter = BatchedFitter(smpl, lr=1e-3, shape_weight=1e-3, use_cuda_graph=False)
fitter.fit(rot, torch.zeros_like(betas), cam, label, iterations=iters)
torch.cuda.synchronize()
