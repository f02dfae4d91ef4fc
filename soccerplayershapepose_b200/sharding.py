"""Batch sharding of the SMPL path across the GPUs of one node.

The layer has no trainable parameters, so bodies are independent units: rank g of G takes the contiguous rows
[g*B/G, (g+1)*B/G) of the batch, packs its own copy of the model and runs forward / backward on its shard with
NO data-path collective.  The only collectives are bookkeeping: max-over-ranks timing for bench.py and, when the
caller wants whole-batch results on every rank, an all-gather of the outputs.  Works with any torch.distributed
backend (NCCL on the GPUs; gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_range(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Rows [begin, end) of rank `rank`: contiguous, sizes differ by at most one, every row exactly once."""
    if not (0 <= rank < world) or batch < 0:
        raise ValueError("bad shard request")
    return (batch * rank) // world, (batch * (rank + 1)) // world


def shard(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    b, e = shard_range(t.shape[0], rank, world)
    return t[b:e]


def max_over_ranks(value: float, device=None) -> float:
    """The slowest rank decides: device time of a multi-GPU step is the max over ranks."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_shards(local: torch.Tensor, batch: int) -> torch.Tensor:
    """All-gather per-rank shards (possibly of unequal length) back into the whole batch, in row order."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_range(batch, r, world)[1] - shard_range(batch, r, world)[0] for r in range(world)]
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    outs: List[torch.Tensor] = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad)
    return torch.cat([o[:n] for o, n in zip(outs, sizes)], 0)


def aggregate_throughput(bodies_per_rank: int, world: int, steps: int, max_ms_total: float) -> float:
    """Whole-job meshes/s: units all ranks processed / the max-over-ranks time."""
    return bodies_per_rank * world * steps / (max_ms_total / 1e3)
