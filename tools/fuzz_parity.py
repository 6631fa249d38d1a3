"""Differential fuzzing on the GPU: many seeded maps (axial kd trees and q3map-style tilted trees),
many segment families (mixed, short, start-grid adversarial, extreme floats), every staging variant,
start grid on and off, axis node records vs general planes only, segment binning on and off, random phase budgets / block sizes /
shared-memory stack slots, 16-byte and compact 8-byte results -- each result record compared bit for bit with the oracle port.
    python tools/fuzz_parity.py --seeds 20
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import bsp_b200
    import oracle
    from helpers import grid_adversarial_segments, mixed_segments
    from tools import mapgen, seggen
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=10)
    ap.add_argument("--first", type=int, default=100)
    a = ap.parse_args()
    oracle.build(); bsp_b200.build()
    total, bad_total, t0 = 0, 0, time.time()
    for seed in range(a.first, a.first + a.seeds):
        rs = np.random.Generator(np.random.PCG64(seed))
        maps = {"rooms8": mapgen.rooms8(seed), "tilted": mapgen.tilted(seed, n_brushes=int(rs.integers(40, 400)), leaf_max=int(rs.integers(1, 4)))}
        for name, m in maps.items():
            lo, hi = mapgen.seg_bounds(m)
            segs = [mixed_segments(m, 60_000, 40_000, 60_000, 4096, seed), seggen.long_rays(seed, 0, 60_000, lo, hi, 0.01, 8.0),
                    grid_adversarial_segments(m, 80_000, seed)]
            ext = seggen.uniform(seed + 7, 0, 20_000, lo, hi)
            specials = np.array([np.inf, -np.inf, 3.0e38, -3.0e38, 1.0e-40, -1.0e-40, 0.0, -0.0, np.nan, 1.0e30], np.float32)
            for k in range(0, len(ext), 2):
                ext[k, int(rs.choice([0, 1, 2, 4, 5, 6]))] = specials[int(rs.integers(0, len(specials)))]
            segs.append(ext)
            segs = np.vstack(segs)
            # a sprinkling of exact zeros (both signs): rays that must leave the axis shortcut (trace_core.h: ray_is_plain)
            z = rs.integers(0, len(segs), len(segs) // 20)
            segs[z, rs.choice([0, 1, 2, 4, 5, 6], len(z))] = np.array([0.0, -0.0], np.float32)[rs.integers(0, 2, len(z))]
            with np.errstate(all="ignore"):
                exp, _, _ = oracle.Port(m).trace(segs, want_pos=False)
            want8 = (exp["flags"] | ((exp["plane"] + 1).astype(np.uint32) << 2)).astype(np.uint32)
            d_segs = torch.from_numpy(segs).cuda()
            for general in (False, True):
                g = bsp_b200.BspGpu(m, general_planes_only=general)
                variants = (1, 3) if g.info()["staged_bytes"] < 200 * 1024 else (3,)
                for variant in variants:
                    for grid in (0, 1):
                        g.set_option(1, variant); g.set_option(13, grid)
                        g.set_option(2, int(rs.choice([0, 64, 256, 512, 1024])))
                        g.set_option(6, int(rs.integers(1, 33))); g.set_option(7, int(rs.integers(1, 40))); g.set_option(10, int(rs.integers(1, 40)))
                        g.set_option(11, int(rs.integers(1, 33))); g.set_option(8, int(rs.choice([0, 32, 96, 4096])))
                        g.set_option(15, int(rs.choice([0, 2, 4, 6, 8, -1]))); g.set_option(16, int(rs.integers(0, 2)))   # smem stack slots, segment binning
                        compact = bool(rs.integers(0, 2))
                        if rs.integers(0, 2):                      # device-resident path (the one binning applies to)
                            out = torch.zeros((len(segs), 2 if compact else 4), dtype=torch.int32, device="cuda")
                            g.trace(d_segs, out=out, compact=compact)
                            res = out.cpu().numpy().view(bsp_b200.RESULT8_DT if compact else bsp_b200.RESULT_DT).reshape(-1)
                        else:
                            res = g.trace(segs, compact=compact)
                        if compact:
                            wrong = (res["t"].view(np.uint32) != exp["t"].view(np.uint32)) | (res["bits"] != want8)
                        else:
                            wrong = (res.view(np.uint32).reshape(-1, 4) != exp.view(np.uint32).reshape(-1, 4)).any(axis=1)
                        bad = int(wrong.sum())
                        total += len(segs); bad_total += bad
                        if bad:
                            i = int(np.nonzero(wrong)[0][0])
                            print(f"MISMATCH seed {seed} map {name} general {general} variant {variant} grid {grid} compact {compact}: {bad}; first {i}: seg {segs[i]} got {res[i]} want {exp[i]}", flush=True)
                g.close()
        print(f"seed {seed}: {total} traces compared so far, {bad_total} mismatches, {time.time() - t0:.0f}s", flush=True)
    print("FUZZ", "OK" if bad_total == 0 else "FAILED", total, "traces,", bad_total, "mismatches")
    sys.exit(0 if bad_total == 0 else 1)


if __name__ == "__main__":
    main()
