"""phase budgets (node steps / side steps per phase) across the bench workloads, device-resident (development aid)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bsp_b200
import bench

n = 32 * 1024 * 1024
for wl in ("city10k-long", "mega200k-short", "city10k-camera", "rooms8-uniform"):
    w = bench.workload_setup(wl)
    g = bsp_b200.BspGpu(w["path"], 0); g.set_stream(torch.cuda.current_stream().cuda_stream)
    segs = torch.empty((n, 8), dtype=torch.float32, device="cuda"); out = torch.empty((n, 4), dtype=torch.int32, device="cuda")
    bench.device_segments(g, w, segs, 0, n)
    line = [wl]
    for steps, side, thr in ((0, 6, 16), (0, 2, 16), (0, 3, 16), (0, 4, 16), (0, 5, 16), (0, 3, 12), (0, 4, 20), (12, 3, 16)):
        g.set_option(7, steps); g.set_option(10, side); g.set_option(11, thr)
        g.trace(segs, out=out); best = 1e9
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.trace(segs, out=out, async_=True); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
        line.append("%d/%d/%d: %.0f" % (steps, side, thr, n / best / 1e3))
    print("  ".join(line), "(variant %d)" % g.info()["variant"], flush=True)
    g.close()
