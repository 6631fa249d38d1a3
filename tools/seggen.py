"""Counter-based segment generators (numpy), bit-identical to the CUDA generator
kernel ``bspgpu_generate`` (bsp_b200/csrc/seggen.cuh).

Every float is derived from ``(seed, global segment index, slot)`` alone, so any
rank can generate exactly its shard, and host and device agree bit for bit:

    key      = fin(seed + GOLDEN)                                  (splitmix64 finaliser)
    u(i, k)  = fin(key + (i*SLOTS + k + 1) * GOLDEN)   mod 2^64     k in [0, SLOTS)
    f(i, k)  = float32(u >> 40) * 2^-24                             in [0,1), exact

All later arithmetic is IEEE binary32 with one rounding per operation (numpy
ufuncs here; ``-fmad=false -prec-div=true -prec-sqrt=true`` on the device), written
in the same order on both sides.  Output layout is the device layout: ``(n, 8)``
float32 = ``{start.xyz, 0, end.xyz, 0}`` (two float4 per segment).

Kinds:
    UNIFORM  both endpoints uniform in the box                        (C1, C2)
    LONG     start uniform, direction ~isotropic (rejection-sampled in the unit
             ball, 8 tries, then normalised), length uniform in [len_lo, len_hi]  (C3; C5 with short lengths)
    CAMERA   pinhole grids: segment = eye -> eye + far * (fwd + right*sx + up*sy)  (C4)
"""
from __future__ import annotations

import numpy as np

GOLDEN = np.uint64(0x9E3779B97F4A7C15)
SLOTS = 32
KIND_UNIFORM, KIND_LONG, KIND_CAMERA = 0, 1, 2


def _fin(z: np.ndarray) -> np.ndarray:
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def make_key(seed: int) -> int:
    with np.errstate(over="ignore"):
        return int(_fin(np.array([seed], np.uint64) + GOLDEN)[0])


def unit_float(key: int, idx: np.ndarray, slot: int) -> np.ndarray:
    """f(i, slot) for an array of global indices (uint64)."""
    with np.errstate(over="ignore"):
        ctr = idx * np.uint64(SLOTS) + np.uint64(slot + 1)
        u = _fin(np.uint64(key) + ctr * GOLDEN)
    return (u >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24)


def _point(key, idx, slot0, lo, ext):
    return [lo[a] + ext[a] * unit_float(key, idx, slot0 + a) for a in range(3)]


def _pack(start, end) -> np.ndarray:
    n = start[0].shape[0]
    out = np.zeros((n, 8), np.float32)
    for a in range(3):
        out[:, a] = start[a]
        out[:, 4 + a] = end[a]
    return out


def uniform(seed: int, first: int, count: int, lo, hi) -> np.ndarray:
    lo = np.asarray(lo, np.float32); hi = np.asarray(hi, np.float32)
    ext = hi - lo
    key = make_key(seed)
    idx = np.arange(first, first + count, dtype=np.uint64)
    return _pack(_point(key, idx, 0, lo, ext), _point(key, idx, 3, lo, ext))


def long_rays(seed: int, first: int, count: int, lo, hi, len_lo: float, len_hi: float) -> np.ndarray:
    lo = np.asarray(lo, np.float32); hi = np.asarray(hi, np.float32)
    ext = hi - lo
    len_lo = np.float32(len_lo); len_ext = np.float32(len_hi) - len_lo
    key = make_key(seed)
    idx = np.arange(first, first + count, dtype=np.uint64)
    start = _point(key, idx, 0, lo, ext)
    one, two = np.float32(1.0), np.float32(2.0)
    qmin = np.float32(2.0 ** -10)
    vx = np.zeros(count, np.float32); vy = np.zeros(count, np.float32); vz = np.zeros(count, np.float32)
    q = np.zeros(count, np.float32)
    done = np.zeros(count, bool)
    for j in range(8):
        cx = two * unit_float(key, idx, 3 + 3 * j) - one
        cy = two * unit_float(key, idx, 4 + 3 * j) - one
        cz = two * unit_float(key, idx, 5 + 3 * j) - one
        cq = (cx * cx + cy * cy) + cz * cz
        take = ~done  # a later try replaces an earlier one only while none was accepted
        vx = np.where(take, cx, vx); vy = np.where(take, cy, vy); vz = np.where(take, cz, vz)
        q = np.where(take, cq, q)
        done |= take & (cq <= one) & (cq >= qmin)
    bad = q < qmin  # only possible when all 8 tries failed and the last one is tiny
    vx = np.where(bad, one, vx); vy = np.where(bad, np.float32(0), vy); vz = np.where(bad, np.float32(0), vz)
    q = np.where(bad, one, q)
    s = np.sqrt(q)
    d = [vx / s, vy / s, vz / s]
    ln = len_lo + len_ext * unit_float(key, idx, 27)
    end = [start[a] + d[a] * ln for a in range(3)]
    return _pack(start, end)


def make_views(seed: int, count: int, lo, hi, free_test=None) -> np.ndarray:
    """``count`` camera poses as a (count, 12) float32 array {eye, fwd, right, up}.
    Orientations come from rational rotations (no libm), eyes are rejection-sampled
    with ``free_test(points)->bool mask`` when given."""
    from .mapgen import Rng, ROTATIONS
    rng = Rng(seed)
    lo = np.asarray(lo, np.float64); hi = np.asarray(hi, np.float64)
    views = np.zeros((count, 12), np.float32)
    for v in range(count):
        for _ in range(1000):
            eye = np.array([lo[a] + (hi[a] - lo[a]) * (rng.randint(0, 1 << 20) / float(1 << 20)) for a in range(3)])
            if free_test is None or bool(free_test(eye[None, :].astype(np.float32))[0]):
                break
        a, b, c = ROTATIONS[rng.randint(0, len(ROTATIONS) - 1)]
        co, si = a / c, b / c
        quad = rng.randint(0, 3)
        for _ in range(quad):
            co, si = -si, co
        p, q_, r = ROTATIONS[rng.randint(0, len(ROTATIONS) - 1)]
        pitch_s = -(min(p, q_) / r) * 0.5  # look a little downwards
        pitch_c = float(np.sqrt(1.0 - pitch_s * pitch_s))
        fwd = np.array([co * pitch_c, si * pitch_c, pitch_s])
        right = np.array([si, -co, 0.0])
        up = np.cross(right, fwd)
        views[v] = np.concatenate((eye, fwd, right, up)).astype(np.float32)
    return views


def camera(views: np.ndarray, width: int, height: int, first: int, count: int, tan_half: float = 1.0,
           far: float = 32768.0) -> np.ndarray:
    """row-major pinhole grids; global index = (view*height + py)*width + px."""
    views = np.asarray(views, np.float32)
    idx = np.arange(first, first + count, dtype=np.uint64)
    per = np.uint64(width * height)
    v = (idx // per).astype(np.int64)
    pix = idx % per
    px = (pix % np.uint64(width)).astype(np.float32)
    py = (pix // np.uint64(width)).astype(np.float32)
    half, one = np.float32(0.5), np.float32(1.0)
    tanx = np.float32(tan_half)
    tany = np.float32(np.float32(tan_half) * np.float32(height) / np.float32(width))
    sx = ((px + half) * np.float32(2.0 / width) - one) * tanx
    sy = (one - (py + half) * np.float32(2.0 / height)) * tany
    far = np.float32(far)
    start = [views[v, a] for a in range(3)]
    end = []
    for a in range(3):
        ray = (views[v, 3 + a] + views[v, 6 + a] * sx) + views[v, 9 + a] * sy
        end.append(start[a] + ray * far)
    return _pack(start, end)


# --------------------------------------------------------------------------- edge cases (host only)

def edge_cases(m, count: int, seed: int = 7) -> np.ndarray:
    """Hand-built families that exercise ties, +-0 and on-plane paths:
    zero-length, axis-parallel, start/end exactly on a brush face, both endpoints in a
    node plane, grazing along a face, -0.0 coordinates, segments through brush corners."""
    from .mapgen import seg_bounds
    rs = np.random.Generator(np.random.PCG64(seed))
    lo, hi = seg_bounds(m)
    lo = lo.astype(np.float64); hi = hi.astype(np.float64)
    planes_n = m.planes["n"].astype(np.float64); planes_d = m.planes["d"].astype(np.float64)
    axial = np.nonzero((np.abs(planes_n) == 1.0).sum(1) == 1)[0]
    side_planes = m.brush_sides["plane"]
    side_axial = side_planes[np.isin(side_planes, axial)]
    node_planes = m.nodes["plane"]
    node_axial = node_planes[np.isin(node_planes, axial)]

    def rnd(n):
        return lo + (hi - lo) * rs.random((n, 3))

    fam = max(1, count // 8)
    out = []
    # 0 zero length
    p = rnd(fam); out.append(np.hstack((p, p)))
    # 1 axis parallel
    s = rnd(fam); e = s.copy(); ax = rs.integers(0, 3, fam)
    e[np.arange(fam), ax] = (lo + (hi - lo) * rs.random((fam, 3)))[np.arange(fam), ax]
    out.append(np.hstack((s, e)))

    def on_plane(pts, plane_ids):
        pts = pts.copy()
        n = planes_n[plane_ids]; d = planes_d[plane_ids]
        ax = np.argmax(np.abs(n), axis=1)
        pts[np.arange(len(pts)), ax] = d * n[np.arange(len(pts)), ax]
        return pts, ax
    if len(side_axial):
        # 2 start exactly on a brush face, 3 end exactly on a brush face
        ids = rs.choice(side_axial, fam); s, _ = on_plane(rnd(fam), ids); out.append(np.hstack((s, rnd(fam))))
        ids = rs.choice(side_axial, fam); e, _ = on_plane(rnd(fam), ids); out.append(np.hstack((rnd(fam), e)))
        # 5 grazing: both endpoints on the same brush face plane
        ids = rs.choice(side_axial, fam); s, _ = on_plane(rnd(fam), ids); e, _ = on_plane(rnd(fam), ids)
        out.append(np.hstack((s, e)))
    if len(node_axial):
        # 4 both endpoints in a node plane
        ids = rs.choice(node_axial, fam); s, _ = on_plane(rnd(fam), ids); e, _ = on_plane(rnd(fam), ids)
        out.append(np.hstack((s, e)))
    # 6 negative zeros / exact zeros
    s = rnd(fam); e = rnd(fam)
    z = rs.integers(0, 3, fam); s[np.arange(fam), z] = -0.0
    z = rs.integers(0, 3, fam); e[np.arange(fam), z] = 0.0
    out.append(np.hstack((s, e)))
    # 7 through brush corners: segment from a random point through a corner of a random brush
    if len(m.brushes):
        b = rs.integers(0, len(m.brushes), fam)
        corner = np.zeros((fam, 3))
        for k in range(fam):
            br = m.brushes[b[k]]
            pl = m.brush_sides["plane"][br["first_side"]:br["first_side"] + br["num_sides"]]
            pn = planes_n[pl]; pd = planes_d[pl]
            c = np.zeros(3)
            for a in range(3):
                cand = [pd[j] * pn[j, a] for j in range(len(pl)) if abs(pn[j, a]) == 1.0]
                c[a] = cand[rs.integers(0, len(cand))] if cand else rs.uniform(lo[a], hi[a])
            corner[k] = c
        s = rnd(fam)
        e = s + (corner - s) * 2.0
        out.append(np.hstack((s, e)))
    segs6 = np.vstack(out).astype(np.float32)[:count]
    full = np.zeros((len(segs6), 8), np.float32)
    full[:, 0:3] = segs6[:, 0:3]
    full[:, 4:7] = segs6[:, 3:6]
    return full
