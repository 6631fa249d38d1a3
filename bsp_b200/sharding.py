"""Multi-GPU plumbing: one process per GPU, segments sharded by contiguous index range, map
replicated, no data-path collective.  torch.distributed (NCCL on the
GPUs, gloo in the CPU tests) carries only the hit-count all-reduce and the result gather."""
from __future__ import annotations

import torch
import torch.distributed as dist

from .api import shard_range


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def my_shard(n_total: int):
    """[first, first+count) of this rank over n_total segments (C ABI: bspgpu_shard_range)."""
    rank, size = world()
    return shard_range(n_total, rank, size)


def reduce_hits(hits: torch.Tensor) -> torch.Tensor:
    """sum of per-rank hit counters (int64 tensor, in place)."""
    if world()[1] > 1:
        dist.all_reduce(hits, op=dist.ReduceOp.SUM)
    return hits


def gather_results(local: torch.Tensor, dst: int = 0):
    """rank `dst` receives the ranks' (n_r, 4) int32 result shards, concatenated in rank order
    (= segment order, because shards are contiguous index ranges); other ranks get None.
    Shards may have different lengths (the last rank's can be shorter)."""
    rank, size = world()
    if size == 1:
        return local
    counts = torch.zeros(size, dtype=torch.int64, device=local.device)
    counts[rank] = local.shape[0]
    dist.all_reduce(counts)
    longest = int(counts.max().item())
    padded = local
    if local.shape[0] < longest:
        padded = torch.zeros((longest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded[: local.shape[0]] = local
    bufs = [torch.empty_like(padded) for _ in range(size)] if rank == dst else None
    dist.gather(padded.contiguous(), bufs, dst=dst)
    if rank != dst:
        return None
    return torch.cat([b[: int(c)] for b, c in zip(bufs, counts.tolist())], dim=0)


def max_over_ranks(value: float, device) -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if world()[1] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
