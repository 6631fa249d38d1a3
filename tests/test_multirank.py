"""N>1 host logic on CPU: two gloo ranks shard a segment set, trace their shards (the oracle
port stands in for the GPU kernel here -- this test is about the sharding / reduction /
gather plumbing bench.py uses, not about tracing), and rank 0 must end up with exactly the
unsharded result and hit count."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_total, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from bsp_b200 import sharding
    from tools import mapgen, seggen
    m, _ = mapgen.get_map("rooms8")
    lo, hi = mapgen.seg_bounds(m)
    first, count = sharding.my_shard(n_total)
    segs = seggen.uniform(42, first, count, lo, hi)            # each rank generates exactly its shard
    res, _, ctr = oracle.Port(m).trace(segs, want_pos=False)
    local = torch.from_numpy(res.view(np.int32).reshape(-1, 4).copy())
    hits = sharding.reduce_hits(torch.tensor([ctr["hits"]], dtype=torch.int64))
    full = sharding.gather_results(local, dst=0)
    slowest = sharding.max_over_ranks(float(rank + 1), "cpu")
    # the parity sample bench.py takes at every world size: a strided sample of every rank's shard, gathered to rank 0
    # (over the library's NCCL on the GPUs, over gloo here) and re-traced there by the checker
    idx = sharding.sample_indices(count, 512)
    rows = sharding.gather_equal(local[idx].contiguous())
    srows = sharding.gather_equal(torch.from_numpy(segs)[idx].contiguous())
    if rank == 0:
        import bench
        parity = bench.check_sample({"m": m}, srows.reshape(-1, 8).numpy(), rows.reshape(-1, 4).numpy())
        assert parity["mismatches"] == 0 and parity["checked"] == world * len(idx)
        broken = rows.clone(); broken[1, 3, 0] ^= 1                       # one flipped bit in one rank's rows must be caught
        assert bench.check_sample({"m": m}, srows.reshape(-1, 8).numpy(), broken.reshape(-1, 4).numpy())["mismatches"] == 1
        np.savez(out_path, results=full.numpy(), hits=int(hits.item()), first=first, count=count, slowest=slowest, sampled=parity["checked"])
    else:
        assert full is None and rows is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [100_001, 64])
def test_two_ranks_reproduce_the_unsharded_result(tmp_path, n_total):
    import torch.multiprocessing as mp
    import oracle
    from tools import mapgen, seggen
    oracle.build()
    mapgen.get_map("rooms8")                                    # warm the on-disk cache before forking
    out = str(tmp_path / "rank0.npz")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, n_total, out), nprocs=2, join=True)
    z = np.load(out)
    m, _ = mapgen.get_map("rooms8")
    lo, hi = mapgen.seg_bounds(m)
    exp, _, ctr = oracle.Port(m).trace(seggen.uniform(42, 0, n_total, lo, hi), want_pos=False)
    assert z["results"].shape == (n_total, 4)
    assert np.array_equal(z["results"].view(np.uint32), exp.view(np.uint32).reshape(-1, 4))
    assert int(z["hits"]) == ctr["hits"]
    assert int(z["first"]) == 0 and int(z["count"]) == (n_total + 1) // 2
    assert float(z["slowest"]) == 2.0
