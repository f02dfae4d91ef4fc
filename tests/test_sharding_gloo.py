"""N > 1 host logic on CPU (gloo, world_size 2): the batch shards cover every body exactly once, a sharded run
of the path (the CPU oracle stands in for the device kernels) reproduces the single-process result with no
data-path collective, and the timing / throughput aggregation takes the max over ranks."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from soccerplayershapepose_b200 import sharding


def test_shard_ranges_partition_the_batch():
    for B in (0, 1, 5, 64, 4097, 65536):
        for world in (1, 2, 3, 4, 8):
            rows = []
            for r in range(world):
                b, e = sharding.shard_range(B, r, world)
                assert 0 <= b <= e <= B
                rows += list(range(b, e))
            assert rows == list(range(B))
            sizes = [sharding.shard_range(B, r, world)[1] - sharding.shard_range(B, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.smpl_oracle import SMPLOracle
        from soccerplayershapepose_b200.model_io import make_synthetic_smpl
        from soccerplayershapepose_b200.synthetic_inputs import make_smpl_inputs
        torch.set_num_threads(2)
        B = 7                                           # ragged: 3 + 4
        x = make_smpl_inputs(B, 0)
        orc = SMPLOracle(make_synthetic_smpl(1234), dtype=torch.float32)
        b, r, t = (sharding.shard(x[k], rank, world) for k in ("betas", "rotmats", "trans"))
        out = orc.forward_flat(b, r, t, pose2rot=False)            # this rank's bodies only, no communication
        joints = sharding.gather_shards(out.joints.contiguous(), B)
        verts_sum = sharding.gather_shards(out.vertices.sum(dim=(1, 2)).contiguous(), B)
        slow = sharding.max_over_ranks(10.0 + rank)                # the slowest rank decides
        if rank == 0:
            np.savez(os.path.join(out_dir, "r0.npz"), joints=joints.numpy(), verts_sum=verts_sum.numpy(), slow=slow)
    finally:
        dist.destroy_process_group()


def test_sharded_run_matches_single_process(tmp_path, synthetic_model):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = np.load(os.path.join(str(tmp_path), "r0.npz"))
    from oracle.smpl_oracle import SMPLOracle
    from soccerplayershapepose_b200.synthetic_inputs import make_smpl_inputs
    x = make_smpl_inputs(7, 0)
    ref = SMPLOracle(synthetic_model, dtype=torch.float32).forward_flat(x["betas"], x["rotmats"], x["trans"], pose2rot=False)
    np.testing.assert_allclose(got["joints"], ref.joints.numpy(), atol=1e-6)
    np.testing.assert_allclose(got["verts_sum"], ref.vertices.sum(dim=(1, 2)).numpy(), rtol=1e-5)
    assert float(got["slow"]) == 11.0
    assert sharding.aggregate_throughput(4096, 8, 20, 1000.0) == 4096 * 8 * 20 / 1.0
