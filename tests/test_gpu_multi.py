"""Multi-GPU entry points behind the C ABI (include/bspgpu_multi.h) on min(2, device_count) GPUs: the in-process
bspgpu_multi_* (one handle + one host thread per GPU, NCCL communicators owned by the library) and the rank/world
bspgpu_dist_* (one process per GPU).  Ground truth: the unmodified reference object where it was built, else the port."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import bits

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _devices():
    import torch
    return list(range(min(2, torch.cuda.device_count())))


def test_multi_trace_host_arrays_against_the_reference(native, maps, oracle_mod):
    from bsp_b200 import multi
    from tools import mapgen, seggen
    m, path = maps("city10k")
    lo, hi = mapgen.seg_bounds(m)
    diag = float(np.sqrt(((hi - lo).astype(np.float64) ** 2).sum()))
    segs = np.vstack((seggen.edge_cases(m, 4096, 5), seggen.long_rays(3, 10_000_000, 500_001, lo, hi, 0.25 * diag, diag)))
    mg = multi.BspGpuMulti(path, devices=_devices())
    assert mg.world == len(_devices())
    pos = np.zeros((len(segs), 4), np.float32)
    res = mg.trace(segs, out=np.zeros(len(segs), native.RESULT_DT), pos=pos)
    exp, exp_pos, ctr = oracle_mod.Port(m).trace(segs)
    assert np.array_equal(bits(res).reshape(-1, 4), bits(exp).reshape(-1, 4))
    assert np.array_equal(bits(pos[:, :3]), bits(exp_pos))
    assert mg.hit_count() == ctr["hits"]                       # ncclAllReduce of the per-GPU counters
    if oracle_mod.have_reference():
        r = oracle_mod.Reference(path).trace(segs)
        assert np.array_equal((res["flags"] & 1).astype(np.uint8), r["hit"]) and np.array_equal(bits(res["t"]), bits(r["t"]))
    # packed 24-byte segments + compact results, ragged sizes (shards of different length, empty shards, the one-GPU small path)
    packed = np.ascontiguousarray(np.hstack((segs[:, 0:3], segs[:, 4:7])))
    for n in (0, 1, 7, 40_001, len(segs)):
        r8 = mg.trace(packed[:n], packed24=True, compact=True)
        assert np.array_equal(bits(r8["t"]), bits(exp["t"][:n])) and np.array_equal(r8["bits"] & 3, exp["flags"][:n])
        assert mg.hit_count() == int((exp["flags"][:n] & 1).sum())
        # one float per segment (BSPGPU_OUT_COMPACT4): t on a hit, T_MISS on a miss; shards land at 4-byte offsets
        r4 = mg.trace(packed[:n], packed24=True, compact=4)
        hit = (exp["flags"][:n] & 1) != 0
        assert r4.dtype == np.float32 and np.array_equal(bits(r4), bits(np.where(hit, exp["t"][:n], np.float32(native.T_MISS)).astype(np.float32)))
    mg.close()


def test_multi_resident_shards_and_nccl_gather(native, maps, oracle_mod):
    """every GPU generates and traces its own shard; the result shards are gathered over NCCL onto one GPU in segment order"""
    from bsp_b200 import multi
    from tools import mapgen, seggen
    m, path = maps("rooms8")
    lo, hi = mapgen.seg_bounds(m)
    n_total, first = 3_000_001, 5_000_000
    mg = multi.BspGpuMulti(path, devices=_devices())
    mg.generate(first, n_total, 0, 2, lo, hi)
    mg.trace_resident()
    hits = mg.hit_count()
    sharded = mg.read_results(0, n_total)                      # read from the owning shards
    assert hits == int((sharded["flags"] & 1).sum())
    mg.gather(dst_rank=mg.world - 1)
    gathered = mg.read_results(0, n_total)
    assert np.array_equal(gathered.view(np.uint32), sharded.view(np.uint32))
    idx = np.arange(0, n_total, 37)
    want_segs = seggen.uniform(2, first, n_total, lo, hi)[idx]
    exp, _, _ = oracle_mod.Port(m).trace(want_segs, want_pos=False)
    assert np.array_equal(bits(gathered[idx]).reshape(-1, 4), bits(exp).reshape(-1, 4))
    mg.close()


def test_dist_entry_points_single_rank(native, maps, oracle_mod):
    """the rank/world variant with a world of one: communicator set-up, all-reduce and gather degenerate to copies"""
    import torch
    from bsp_b200 import multi
    m, path = maps("rooms8")
    from tools import mapgen, seggen
    lo, hi = mapgen.seg_bounds(m)
    segs = seggen.uniform(4, 0, 200_000, lo, hi)
    g = native.BspGpu(path)
    multi.dist_init(g, multi.dist_unique_id(), 0, 1)
    d_segs = torch.from_numpy(segs).cuda(); out = torch.zeros((len(segs), 4), dtype=torch.int32, device="cuda")
    g.trace(d_segs, out=out)
    exp, _, ctr = oracle_mod.Port(m).trace(segs, want_pos=False)
    dev = torch.zeros(1, dtype=torch.int64, device="cuda")
    assert multi.dist_hit_count(g, dev_i64=dev) == ctr["hits"] and int(dev.item()) == ctr["hits"]
    full = torch.zeros((1,) + tuple(out.shape), dtype=torch.int32, device="cuda")
    multi.dist_gather(g, out, full, 0); g.sync()
    assert np.array_equal(full[0].cpu().numpy().view(np.uint32), exp.view(np.uint32).reshape(-1, 4))
    with pytest.raises(native.BspGpuError):
        multi.dist_init(g, multi.dist_unique_id(), 0, 1)      # already has a communicator
    multi.dist_shutdown(g)
    g.close()


_RANK_SCRIPT = r'''
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np, torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
import bsp_b200, oracle
from bsp_b200 import sharding
from tools import mapgen, seggen
m, path = mapgen.get_map("rooms8")
lo, hi = mapgen.seg_bounds(m)
n_total = 1_000_003
first, count = sharding.my_shard(n_total)
g = bsp_b200.BspGpu(path, device=rank)
assert sharding.init_library_comm(g)
segs = torch.empty((count, 8), dtype=torch.float32, device="cuda"); out = torch.empty((count, 4), dtype=torch.int32, device="cuda")
g.generate(segs, first, count, 0, 2, lo, hi)
g.trace(segs, out=out)
hits = torch.zeros(1, dtype=torch.int64, device="cuda")
sharding.reduce_hits(hits, g)                                   # ncclAllReduce inside libbspgpu.so
idx = sharding.sample_indices(count, 2048).cuda()
rows = sharding.gather_equal(out[idx].contiguous(), g)          # ncclSend/ncclRecv inside libbspgpu.so
srows = sharding.gather_equal(segs[idx].contiguous(), g)
torch.cuda.synchronize()
if rank == 0:
    exp, _, ctr = oracle.Port(m).trace(seggen.uniform(2, 0, n_total, lo, hi), want_pos=False)
    assert int(hits.item()) == ctr["hits"], (int(hits.item()), ctr["hits"])
    e2, _, _ = oracle.Port(m).trace(srows.reshape(-1, 8).cpu().numpy(), want_pos=False)
    assert np.array_equal(rows.reshape(-1, 4).cpu().numpy().view(np.uint32), e2.view(np.uint32).reshape(-1, 4))
    print("RANKS-OK", world)
dist.barrier(); g.close(); dist.destroy_process_group()
'''


def test_dist_entry_points_two_processes(native, maps, oracle_mod, tmp_path):
    """one process per GPU (torchrun shape): the NCCL unique id travels over torch.distributed, the collectives run inside libbspgpu.so"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    maps("rooms8")
    script = tmp_path / "rank.py"
    script.write_text(_RANK_SCRIPT.format(root=ROOT))
    port = 29600 + (os.getpid() % 1000)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "RANKS-OK 2" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
