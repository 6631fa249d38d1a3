"""two-level node records on/off (BSPGPU_OPT_PACKED_NODES) on the bench workloads, device-resident (development aid)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bsp_b200
import bench

n = 32 * 1024 * 1024
for wl, variant in (("city10k-long", 3), ("mega200k-short", 3), ("city10k-camera", 3), ("rooms8-uniform", 3), ("rooms8-uniform", 1)):
    w = bench.workload_setup(wl)
    g = bsp_b200.BspGpu(w["path"], 0); g.set_stream(torch.cuda.current_stream().cuda_stream)
    segs = torch.empty((n, 8), dtype=torch.float32, device="cuda"); out = torch.empty((n, 4), dtype=torch.int32, device="cuda")
    bench.device_segments(g, w, segs, 0, n)
    g.set_option(1, variant)
    line = [wl, "variant %d" % variant]
    ref = None
    for packed, steps in ((0, 0), (1, 0), (1, 4), (1, 6), (1, 12)):
        if variant == 1 and (packed, steps) != (0, 0):
            continue
        g.set_option(14, packed); g.set_option(7, steps)
        g.trace(segs, out=out); best = 1e9
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.trace(segs, out=out, async_=True); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
        chk = int(out.to(torch.int64).sum().item())
        if ref is None: ref = chk
        assert chk == ref, "results differ"
        line.append("p%d/s%d: %.1f M/s (%d CTAs)" % (packed, steps, n / best / 1e3, g.info()["grid_ctas"]))
    print("  ".join(line), flush=True)
    g.close()
