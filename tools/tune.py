"""Knob sweep on the GPU (development aid): traces/s of the device-resident path for a map /
segment kind over variants x block sizes x refill x max_steps.  Prints one line per setting."""
from __future__ import annotations

import argparse
import itertools
import json
import sys
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import bsp_b200
    from tools import mapgen
    ap = argparse.ArgumentParser()
    ap.add_argument("--map", default="rooms8")
    ap.add_argument("--kind", default="uniform", choices=["uniform", "long", "short", "camera"])
    ap.add_argument("--n", type=int, default=16 * 1024 * 1024)
    ap.add_argument("--variants", default="0")
    ap.add_argument("--threads", default="0")
    ap.add_argument("--refill", default="8")
    ap.add_argument("--steps", default="0")
    ap.add_argument("--ctas", default="0")
    ap.add_argument("--top", default="98304")
    ap.add_argument("--block_segs", default="256")
    ap.add_argument("--sched", default="0")
    ap.add_argument("--side_steps", default="3")
    ap.add_argument("--leaf_thr", default="16")
    ap.add_argument("--slots", default="-1", help="BSPGPU_OPT_SMEM_STACK")
    ap.add_argument("--bin", type=int, default=0, help="BSPGPU_OPT_BIN_SEGMENTS")
    ap.add_argument("--general", type=int, default=0, help="flatten every node plane as a general plane")
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    bsp_b200.build()
    m, path = mapgen.get_map(a.map)
    lo, hi = mapgen.seg_bounds(m)
    diag = float(np.sqrt(((hi - lo).astype(np.float64) ** 2).sum()))
    g = bsp_b200.BspGpu(m, 0, general_planes_only=bool(a.general))
    g.set_stream(torch.cuda.current_stream().cuda_stream)
    g.set_option(16, a.bin)
    segs = torch.empty((a.n, 8), dtype=torch.float32, device="cuda")
    if a.kind == "uniform":
        g.generate(segs, 0, a.n, 0, 2, lo, hi)
    elif a.kind == "long":
        g.generate(segs, 0, a.n, 1, 3, lo, hi, 0.25 * diag, diag)
    elif a.kind == "camera":
        from tools import seggen
        views = seggen.make_views(4, 64, lo, hi)
        g.generate(segs, 0, a.n, 2, 0, lo, hi, views=views, width=3840, height=2160)
    else:
        g.generate(segs, 0, a.n, 1, 5, lo, hi, 1.0, 32.0)
    out = torch.empty((a.n, 4), dtype=torch.int32, device="cuda")
    ints = lambda s: [int(x) for x in s.split(",")]
    for variant, threads, refill, steps, ctas, top, bs, sched, ss, lt, slots in itertools.product(
            ints(a.variants), ints(a.threads), ints(a.refill), ints(a.steps), ints(a.ctas), ints(a.top), ints(a.block_segs),
            ints(a.sched), ints(a.side_steps), ints(a.leaf_thr), ints(a.slots)):
        g.set_option(9, sched); g.set_option(10, ss); g.set_option(11, lt); g.set_option(15, slots)
        g.set_option(1, variant); g.set_option(2, threads); g.set_option(6, refill); g.set_option(7, steps)
        g.set_option(3, ctas); g.set_option(4, top); g.set_option(8, bs)
        try:
            g.trace(segs, out=out)  # warm-up
            best = 1e30
            for _ in range(a.reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); g.trace(segs, out=out, async_=True); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            inf = g.info()
            print(json.dumps({"map": a.map, "kind": a.kind, "n": a.n, "variant": inf["variant"], "threads": threads, "grid": inf["grid_ctas"],
                              "smem": inf["smem_bytes"], "slots": inf["smem_stack_slots"], "prepass_ms": round(inf["last_prepass_ms"], 3), "sched": sched, "refill": refill, "max_steps": steps, "side_steps": ss, "leaf_thr": lt, "block_segs": bs, "top": top, "ms": round(best, 3),
                              "Mtraces_s": round(a.n / best / 1e3, 1), "hits": g.hit_count()}), flush=True)
        except bsp_b200.BspGpuError as e:
            print(json.dumps({"variant": variant, "threads": threads, "error": str(e)}), flush=True)


if __name__ == "__main__":
    main()
