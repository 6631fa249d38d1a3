"""Caller-integration demo: a frame's worth of camera rays through the batched
API.  The reference issues ONE debug segment per frame from `game_frame` (bsp.cc:357-370); here a
whole pinhole grid (eye -> eye + far * ray per pixel, the C4 workload) goes through one
`trace` call and the hit fractions become a depth image (binary PGM, no image library needed).

    python tools/render_depth.py --map city10k --width 1920 --height 1080 --out gpurun_out/depth.pgm
    python tools/render_depth.py --check      # small frame, compared bit-for-bit with the oracle
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def render(g, views, width, height, view=0, far=32768.0, device=True):
    """returns (t, hit) arrays of shape (height, width) for one view."""
    import bsp_b200
    n = width * height
    if device:
        import torch
        segs = torch.empty((n, 8), dtype=torch.float32, device="cuda")
        g.generate(segs, view * n, n, 2, 0, (0, 0, 0), (1, 1, 1), views=views, width=width, height=height, far=far)
        out = torch.empty((n, 4), dtype=torch.int32, device="cuda")
        g.trace(segs, out=out)
        res = out.cpu().numpy().view(bsp_b200.RESULT_DT).reshape(-1)
    else:
        from tools import seggen
        res = g.trace(seggen.camera(views, width, height, view * n, n, far=far))
    return res["t"].reshape(height, width), (res["flags"] & 1).reshape(height, width).astype(bool)


def write_pgm(path, t, hit):
    img = np.where(hit, 255.0 * (1.0 - np.sqrt(np.clip(t, 0.0, 1.0))), 0.0).astype(np.uint8)
    with open(path, "wb") as f:
        f.write(b"P5\n%d %d\n255\n" % (img.shape[1], img.shape[0]))
        f.write(img.tobytes())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--map", default="city10k")
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--view", type=int, default=0)
    ap.add_argument("--out", default="depth.pgm")
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    import bsp_b200
    from tools import mapgen, seggen
    bsp_b200.build()
    m, path = mapgen.get_map(a.map)
    lo, hi = mapgen.seg_bounds(m)
    views = seggen.make_views(4, 8, lo, hi)
    g = bsp_b200.BspGpu(path)
    if a.check:
        import oracle
        w, h = 320, 180
        t, hit = render(g, views, w, h, a.view)
        exp, _, _ = oracle.Port(m).trace(seggen.camera(views, w, h, a.view * w * h, w * h))
        ok = np.array_equal(t.reshape(-1).view(np.uint32), exp["t"].view(np.uint32)) and np.array_equal(hit.reshape(-1), (exp["flags"] & 1).astype(bool))
        print("depth frame %dx%d: %d hits, %s the oracle" % (w, h, int(hit.sum()), "matches" if ok else "DIFFERS FROM"))
        sys.exit(0 if ok else 1)
    t, hit = render(g, views, a.width, a.height, a.view)
    write_pgm(a.out, t, hit)
    print(f"{a.out}: {a.width}x{a.height}, {int(hit.sum())} of {hit.size} rays hit, kernel {g.info()['last_kernel_ms']:.3f} ms")


if __name__ == "__main__":
    main()
