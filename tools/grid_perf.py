"""start grid on/off and scheduler comparison on short / long workloads (development aid)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bsp_b200
from tools import mapgen
for name, kind in (("mega200k", "short"), ("city10k", "short"), ("rooms8", "short"), ("city10k", "long"), ("rooms8", "uniform")):
    m, path = mapgen.get_map(name)
    lo, hi = mapgen.seg_bounds(m)
    diag = float(np.sqrt(((hi - lo).astype(np.float64) ** 2).sum()))
    g = bsp_b200.BspGpu(path, 0); g.set_stream(torch.cuda.current_stream().cuda_stream)
    n = 32 * 1024 * 1024
    segs = torch.empty((n, 8), dtype=torch.float32, device="cuda"); out = torch.empty((n, 4), dtype=torch.int32, device="cuda")
    if kind == "short": g.generate(segs, 0, n, 1, 5, lo, hi, 1.0, 32.0)
    elif kind == "long": g.generate(segs, 0, n, 1, 3, lo, hi, 0.25 * diag, diag)
    else: g.generate(segs, 0, n, 0, 2, lo, hi)
    line = [name, kind]
    for sched, grid, threads in ((0, 1, 0), (0, 0, 0), (4, 1, 256), (4, 0, 256), (4, 1, 1024)):
        g.set_option(9, sched); g.set_option(13, grid); g.set_option(2, threads)
        g.trace(segs, out=out); best = 1e9
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.trace(segs, out=out, async_=True); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
        line.append("s%d/g%d/t%d: %.1f M/s" % (sched, grid, threads, n / best / 1e3))
    print("  ".join(line), flush=True)
    g.close()
