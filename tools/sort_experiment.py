import sys, os, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch, bsp_b200
from tools import mapgen
m, path = mapgen.get_map("city10k")
lo, hi = mapgen.seg_bounds(m)
diag = float(np.sqrt(((hi - lo).astype(np.float64) ** 2).sum()))
g = bsp_b200.BspGpu(path, 0)
g.set_stream(torch.cuda.current_stream().cuda_stream)
n = 32 * 1024 * 1024
segs = torch.empty((n, 8), dtype=torch.float32, device="cuda")
g.generate(segs, 0, n, 1, 3, lo, hi, 0.25 * diag, diag)
out = torch.empty((n, 4), dtype=torch.int32, device="cuda")
def timeit(s):
    g.trace(s, out=out)
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.trace(s, out=out, async_=True); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
base = timeit(segs)
print("unsorted ms", base, "M/s", n / base / 1e3)
lo_t = torch.tensor(lo, device="cuda"); ext = torch.tensor(hi - lo, device="cuda")
for (gx, gy, gz) in ((16, 16, 2), (64, 64, 4), (256, 256, 8)):
    q = ((segs[:, 0:3] - lo_t) / ext).clamp(0, 0.999999)
    ix = (q[:, 0] * gx).long(); iy = (q[:, 1] * gy).long(); iz = (q[:, 2] * gz).long()
    key = (iz * gy + iy) * gx + ix
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); order = torch.argsort(key); s2 = segs[order].contiguous(); e1.record(); torch.cuda.synchronize()
    t = timeit(s2)
    print("grid", (gx, gy, gz), "sorted trace ms", t, "M/s", n / t / 1e3, "speedup", base / t, "(torch sort+gather ms", e0.elapsed_time(e1), ")")
# also sort by direction octant + start cell
d = segs[:, 4:7] - segs[:, 0:3]
octant = ((d[:, 0] > 0).long() * 4 + (d[:, 1] > 0).long() * 2 + (d[:, 2] > 0).long())
q = ((segs[:, 0:3] - lo_t) / ext).clamp(0, 0.999999)
key = octant * (64 * 64 * 4) + ((q[:, 2] * 4).long() * 64 + (q[:, 1] * 64).long()) * 64 + (q[:, 0] * 64).long()
order = torch.argsort(key); s2 = segs[order].contiguous()
t = timeit(s2)
print("octant+grid sorted trace ms", t, "M/s", n / t / 1e3, "speedup", base / t)
