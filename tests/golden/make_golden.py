"""Generates tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref/libbspref.so,
built from /root/reference by oracle/Makefile) -- run in the build container, where the
reference exists:   python tests/golden/make_golden.py

Each fixture is self-contained: the IBSP file bytes, the segments, and the reference's
{hit, t, pos} for them (Intersection::plane is null for every row), so
the GPU box -- which has no /root/reference -- can check against the reference's own outputs.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def main():
    import oracle
    from tools import ibsp, mapgen, seggen
    oracle.build()
    assert oracle.have_reference(), "needs /root/reference"
    for name, n_edge, n_rand in (("kat0", 512, 1024), ("kat1", 1024, 2048), ("rooms8", 4096, 8192), ("tilted", 2048, 6144)):
        m, path = mapgen.get_map(name)
        lo, hi = mapgen.seg_bounds(m)
        diag = float(np.sqrt(((hi - lo).astype(np.float64) ** 2).sum()))
        segs = [seggen.edge_cases(m, n_edge, 100), seggen.uniform(101, 0, n_rand, lo, hi),
                seggen.long_rays(102, 0, n_rand // 2, lo, hi, 0.25 * diag, diag),
                seggen.long_rays(103, 0, n_rand // 2, lo, hi, 1.0, 32.0)]
        if name == "kat0":
            vec = mapgen.KAT0_VECTORS
        elif name == "kat1":
            vec = mapgen.KAT1_VECTORS
        else:
            vec = []
        if vec:
            k = np.zeros((len(vec), 8), np.float32)
            for i, v in enumerate(vec):
                k[i, 0:3] = v[0]; k[i, 4:7] = v[1]
            segs.insert(0, k)
        segs = np.vstack(segs)
        ref = oracle.Reference(path)
        r = ref.trace(segs)
        assert int(r["plane_set"].sum()) == 0
        out = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(out, bsp=np.frombuffer(ibsp.to_bytes(m), np.uint8), segs=segs, hit=r["hit"], t=r["t"], pos=r["pos"])
        print(name, len(segs), "segments,", int(r["hit"].sum()), "hits ->", out, os.path.getsize(out), "bytes")
    extreme()


def extreme():
    """rooms8 and the odd map (a solid brush without sides) with infinite / NaN / huge / denormal coordinates: what the
    unmodified reference answers when IEEE arithmetic is simply left to run.  NaN payloads in `pos` are the host's."""
    import oracle
    sys.path.insert(0, os.path.dirname(HERE))
    from helpers import extreme_segments
    from test_hostsim import _odd_map
    from tools import ibsp, mapgen
    for name, m, extent in (("extreme_rooms8", mapgen.get_map("rooms8")[0], 2048), ("extreme_odd", _odd_map(), 80)):
        segs = extreme_segments(24_000, extent, seed=12)
        path = os.path.join("/tmp", name + ".bsp")
        ibsp.write(m, path)
        with np.errstate(all="ignore"):
            r = oracle.Reference(path).trace(segs)
        out = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(out, bsp=np.frombuffer(ibsp.to_bytes(m), np.uint8), segs=segs, hit=r["hit"], t=r["t"], pos=r["pos"])
        print(name, len(segs), "segments,", int(r["hit"].sum()), "hits ->", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    if "--extreme-only" in sys.argv:
        extreme()
        sys.exit(0)
    main()
