"""BSPGPU_OPT_SMEM_STACK with the whole map staged in shared memory (rooms8), device-resident (development aid)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bsp_b200
import bench
n = 32 * 1024 * 1024
for wl in ("rooms8-uniform",):
    w = bench.workload_setup(wl)
    g = bsp_b200.BspGpu(w["path"], 0); g.set_stream(torch.cuda.current_stream().cuda_stream)
    segs = torch.empty((n, 8), dtype=torch.float32, device="cuda"); out = torch.empty((n, 4), dtype=torch.int32, device="cuda")
    bench.device_segments(g, w, segs, 0, n)
    line = [wl]; ref = None
    for slots in (0, 2, 4, 6):
        g.set_option(15, slots)
        g.trace(segs, out=out); best = 1e9
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.trace(segs, out=out, async_=True); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
        chk = int(out.to(torch.int64).sum().item())
        if ref is None: ref = chk
        assert chk == ref, "results differ"
        line.append("slots %d: %.0f M/s (variant %d, %d B smem)" % (slots, n / best / 1e3, g.info()["variant"], g.info()["smem_bytes"]))
    print("  ".join(line), flush=True)
    g.close()
