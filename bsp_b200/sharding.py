"""Multi-GPU plumbing for launchers that run one process per GPU (bench.py under torchrun; the gloo tests on CPU):
segments sharded by contiguous index range, map replicated, no data-path collective.

On the GPUs the two collectives of the path -- the hit-count all-reduce and the result gather -- go through the
library's own NCCL communicator (include/bspgpu_multi.h: bspgpu_dist_*; `init_library_comm` below ships the NCCL
unique id over torch.distributed).  torch.distributed itself only carries launcher chores: the id broadcast, barriers,
the max-over-ranks of the timings, and -- in the CPU tests, where there is no GPU library to ask -- the same two
collectives over gloo."""
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


def sample_indices(count: int, k: int = 4096) -> torch.Tensor:
    """k evenly strided local row indices of a shard of `count` rows (the rows every rank contributes to the parity sample)"""
    if count <= 0:
        return torch.zeros(0, dtype=torch.int64)
    return torch.arange(0, count, max(1, count // k))[:k]


def init_library_comm(g) -> bool:
    """give handle `g` an NCCL communicator over all ranks (bspgpu_dist_init): rank 0 creates the unique id, torch.distributed
    broadcasts its 128 bytes.  Returns False when there is only one rank (nothing to set up)."""
    from . import multi
    rank, size = world()
    if size == 1:
        return False
    dev = torch.device("cuda", torch.cuda.current_device())
    idt = torch.zeros(multi.NCCL_ID_BYTES, dtype=torch.uint8, device=dev)
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(multi.dist_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, src=0)
    multi.dist_init(g, bytes(idt.cpu().numpy().tobytes()), rank, size)
    return True


def reduce_hits(hits: torch.Tensor, g=None) -> torch.Tensor:
    """sum of per-rank hit counters (1-element int64 tensor, in place).  With a handle that owns a library communicator the
    counter of the handle's last trace is all-reduced by the library (ncclAllReduce) straight into `hits`."""
    if world()[1] == 1:
        return hits
    if g is not None:
        from . import multi
        multi.dist_hit_count(g, dev_i64=hits, want_host=False)
    else:
        dist.all_reduce(hits, op=dist.ReduceOp.SUM)
    return hits


def gather_equal(local: torch.Tensor, g=None, dst: int = 0):
    """rank `dst` receives every rank's equally shaped tensor stacked in rank order; other ranks get None.
    g: handle with a library communicator (NCCL send/recv inside libbspgpu.so) -- else torch.distributed."""
    rank, size = world()
    if size == 1:
        return local.unsqueeze(0)
    local = local.contiguous()
    if g is not None:
        from . import multi
        full = torch.empty((size,) + tuple(local.shape), dtype=local.dtype, device=local.device) if rank == dst else None
        multi.dist_gather(g, local, full, dst)
        g.sync()
        return full
    bufs = [torch.empty_like(local) for _ in range(size)] if rank == dst else None
    dist.gather(local, bufs, dst=dst)
    return torch.stack(bufs) if rank == dst else None


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
