#!/usr/bin/env python
"""BASELINE.json configs[4]: regressor training step -- features -> IEF head -> SMPL -> multi-task loss (verts,
joints2D, joints3D, shape, pose) -> backward -> Adam, batch 256 per GPU, data-parallel with NCCL gradient
all-reduce of the head when launched under torchrun.  The CNN encoder is the caller's: synthetic 512-d features.
  python scripts/regressor_bench.py [steps]            # 1 GPU
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/regressor_bench.py"""
import json, os, sys, time
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from soccerplayershapepose_b200 import config, ops, regressor, sharding          # noqa: E402
from soccerplayershapepose_b200.model_io import make_synthetic_smpl             # noqa: E402
from soccerplayershapepose_b200.smpl import SMPL                                # noqa: E402
from soccerplayershapepose_b200.synthetic_inputs import make_smpl_inputs        # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
smpl = SMPL(model_data=make_synthetic_smpl(1234), mode="fp32").to(dev)
torch.manual_seed(0)
head = regressor.IEFModule((1024, 1024), in_features=512).to(dev)
crit = regressor.MultiTaskLoss(("verts", "joints2D", "joints3D", "shape_params", "pose_params"),
                               {"verts": 1.0, "joints2D": 0.1, "joints3D": 1.0, "shape_params": 0.1, "pose_params": 0.1}).to(dev)


class Step(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.head, self.crit = head, crit
        self.r6, self.proj = regressor.gpu_ops()

    def forward(self, feats, labels):
        return self.crit(labels, regressor.predict(self.head, smpl, feats, self.r6, self.proj))[0]


model = Step()
if world > 1 and os.environ.get("B200_REGRESSOR_GRAPH", "1") == "0":
    model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local])
opt = torch.optim.Adam(model.parameters(), lr=1e-4)
x = make_smpl_inputs(B, rank)
rot, betas = x["rotmats"].to(dev), x["betas"].to(dev)
cam = torch.tensor([0.9, 0.0, 0.0], device=dev).repeat(B, 1)
with torch.no_grad():
    t = smpl(betas=betas, body_pose=rot[:, 1:], global_orient=rot[:, :1], pose2rot=False)
    labels = {"verts": t.vertices, "joints3D": t.joints[:, config.ALL_JOINTS_TO_COCO_MAP, :], "shape_params": betas,
              "pose_params_rot_matrices": rot,
              "joints2D": ops.orthographic_project(t.joints, cam, 512.0)[:, config.SMPL_TO_KPRCNN_MAP, :].contiguous()}
feats = torch.randn(B, 512, device=dev)


def step():
    opt.zero_grad(set_to_none=True)
    loss = model(feats, labels)
    loss.backward()
    opt.step()
    return loss


graphed = os.environ.get("B200_REGRESSOR_GRAPH", "1") != "0"
if graphed:
    # one CUDA graph per step: head, SMPL layer, loss, backward, flat-bucket NCCL all-reduce, capturable Adam
    r6, proj = regressor.gpu_ops()
    gstep = regressor.GraphedTrainStep(head, crit, smpl, feats, labels, r6, proj, lr=1e-4, world_size=world)

    def step():
        return gstep(feats, labels)

for _ in range(5):
    l0 = step().clone()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    l1 = step()
e1.record()
torch.cuda.synchronize()
ms = sharding.max_over_ranks(e0.elapsed_time(e1), dev)
nparams = sum(p.numel() for p in model.parameters())
if rank == 0:
    print(json.dumps({"workload": "regressor step: 512-d features -> IEF head (1024,1024) -> SMPL -> 5-term multi-task loss, "
                                  "batch %d per GPU x %d GPU(s)%s" % (B, world, ", one CUDA graph per step" if graphed else ", eager"),
                      "ms_per_step": ms / steps,
                      "crops_per_s": B * world * steps / (ms / 1e3), "allreduced_params": nparams if world > 1 else 0,
                      "loss_first": float(l0), "loss_last": float(l1)}))
if world > 1:
    torch.cuda.synchronize()
    dist.barrier()
    if graphed:
        # a process group whose collectives are referenced by a live CUDA graph can block in its destructor:
        # everything is flushed and synchronised, leave without running it
        sys.stdout.flush()
        os._exit(0)
    dist.destroy_process_group()
