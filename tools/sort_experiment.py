"""N4 experiment: how much faster is the trace when segments arrive sorted by start cell?  (sort done
with torch, NOT timed -- this only measures the lever a binning pre-pass would have.)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bsp_b200
from tools import mapgen

def timeit(g, s, out):
    g.trace(s, out=out)
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.trace(s, out=out, async_=True); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

for name, kind in (("city10k", "long"), ("mega200k", "short"), ("city10k", "short"), ("rooms8", "uniform")):
    m, path = mapgen.get_map(name)
    lo, hi = mapgen.seg_bounds(m)
    diag = float(np.sqrt(((hi - lo).astype(np.float64) ** 2).sum()))
    g = bsp_b200.BspGpu(path, 0)
    g.set_stream(torch.cuda.current_stream().cuda_stream)
    n = 32 * 1024 * 1024
    segs = torch.empty((n, 8), dtype=torch.float32, device="cuda")
    if kind == "long": g.generate(segs, 0, n, 1, 3, lo, hi, 0.25 * diag, diag)
    elif kind == "short": g.generate(segs, 0, n, 1, 5, lo, hi, 1.0, 32.0)
    else: g.generate(segs, 0, n, 0, 2, lo, hi)
    out = torch.empty((n, 4), dtype=torch.int32, device="cuda")
    base = timeit(g, segs, out)
    lo_t = torch.tensor(lo, device="cuda"); ext = torch.tensor(hi - lo, device="cuda")
    line = [f"{name}/{kind}: unsorted {n / base / 1e3:.0f} M/s"]
    for (gx, gy, gz) in ((64, 64, 4), (256, 256, 8), (1024, 1024, 16)):
        q = ((segs[:, 0:3] - lo_t) / ext).clamp(0, 0.999999)
        key = ((q[:, 2] * gz).long() * gy + (q[:, 1] * gy).long()) * gx + (q[:, 0] * gx).long()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); order = torch.argsort(key); s2 = segs[order].contiguous(); e1.record(); torch.cuda.synchronize()
        t = timeit(g, s2, out)
        line.append(f"{gx}x{gy}x{gz}: {n / t / 1e3:.0f} M/s ({base / t:.2f}x; torch sort+gather {e0.elapsed_time(e1):.1f} ms vs trace {base:.1f} ms)")
        del order, s2, key, q
    print("  ".join(line), flush=True)
    g.close()
