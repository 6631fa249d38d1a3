"""Synthetic Quake-3 IBSP maps: brush primitives + a small axial-split BSP compiler.

The reference ships no map (``.gitignore:7`` ignores ``*.bsp``; ``bsp.cc:303``
hard-codes a Warsow map that is not in its tree), so every workload map is made
here, deterministically, from integer seeds:

    kat0 / kat1   the survey's known-answer maps, hand-laid
    rooms8        4x2 hollow rooms, door gaps, crates, non-axial pillars/ramps (~500 brushes)
    city10k       90x90 blocks of box buildings, ~25% with pitched (non-axial) roofs (~10k)
    mega200k      447x447 low blocks (~200k brushes; map >> 228 KB shared memory)

Conventions follow the reference's reader: a plane is ``dot(n,p) - d``; the positive
side is "in front" = outside a brush (bsp.cc:131-137); ``children[0]`` is the front
child, a negative child ``c`` is leaf ``-(c+1)`` (bsp.cc:209,248); a brush is solid
iff ``textures[brush.texture].content_flags & 1`` (bsp.cc:190-193).

No libm transcendental is used anywhere (only + - * / sqrt, all correctly rounded),
so the bytes of a generated map do not depend on the machine that generated it.
"""
from __future__ import annotations

import hashlib
import os

import numpy as np

from . import ibsp

GOLDEN = 0x9E3779B97F4A7C15
MASK64 = (1 << 64) - 1


class Rng:
    """splitmix64 stream of python ints (host-side map layout decisions only)."""

    def __init__(self, seed: int):
        self.state = seed & MASK64

    def u64(self) -> int:
        self.state = (self.state + GOLDEN) & MASK64
        z = self.state
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
        return z ^ (z >> 31)

    def randint(self, lo: int, hi: int) -> int:
        """uniform integer in [lo, hi] (inclusive)."""
        return lo + self.u64() % (hi - lo + 1)

    def chance(self, num: int, den: int) -> bool:
        return self.u64() % den < num


# --------------------------------------------------------------------------- brushes

TEX_SOLID, TEX_NONSOLID, TEX_SOLID_PLUS, TEX_WATER = 0, 1, 2, 3
DEFAULT_TEXTURES = [
    (b"textures/synth/solid", 0, 1),             # CONTENTS_SOLID
    (b"textures/synth/clip_nonsolid", 0, 0),     # never collides
    (b"textures/synth/solid_plus", 0x80, 0x10001),  # solid bit + another content bit
    (b"textures/synth/water", 0, 0x20),          # CONTENTS_WATER: skipped by the & 1 test
]


class BrushList:
    """Accumulates convex brushes: side planes (n.xyz, d) as float32 + a conservative AABB."""

    def __init__(self, dedup_planes: bool = True):
        self.dedup = dedup_planes
        self.planes: list[tuple[float, float, float, float]] = []
        self._plane_index: dict[bytes, int] = {}
        self.side_planes: list[int] = []          # flat, per side
        self.first_side: list[int] = []
        self.num_sides: list[int] = []
        self.texture: list[int] = []
        self.bmin: list[tuple[float, float, float]] = []
        self.bmax: list[tuple[float, float, float]] = []

    def plane(self, nx, ny, nz, d) -> int:
        p = np.array([nx, ny, nz, d], dtype=np.float32)
        if self.dedup:
            key = p.tobytes()
            idx = self._plane_index.get(key)
            if idx is not None:
                return idx
            self._plane_index[key] = len(self.planes)
        self.planes.append(tuple(float(v) for v in p))
        return len(self.planes) - 1

    def add(self, planes, bmin, bmax, texture=TEX_SOLID) -> int:
        self.first_side.append(len(self.side_planes))
        self.num_sides.append(len(planes))
        for (nx, ny, nz, d) in planes:
            self.side_planes.append(self.plane(nx, ny, nz, d))
        self.texture.append(texture)
        self.bmin.append(tuple(float(v) for v in bmin))
        self.bmax.append(tuple(float(v) for v in bmax))
        return len(self.texture) - 1

    # -- primitives ------------------------------------------------------------
    def box(self, x0, x1, y0, y1, z0, z1, texture=TEX_SOLID) -> int:
        """six axial sides in the order (+x,-x,+y,-y,+z,-z) (the KAT-1 ``box`` order)."""
        assert x0 < x1 and y0 < y1 and z0 < z1
        planes = [(1, 0, 0, x1), (-1, 0, 0, -x0), (0, 1, 0, y1), (0, -1, 0, -y0), (0, 0, 1, z1), (0, 0, -1, -z0)]
        return self.add(planes, (x0, y0, z0), (x1, y1, z1), texture)

    def pillar(self, cx, cy, half, z0, z1, rot, texture=TEX_SOLID) -> int:
        """square prism rotated about z by the rational rotation ``rot=(a,b,c)``,
        cos=a/c, sin=b/c (e.g. (3,4,5)): four non-axial sides + top + bottom."""
        a, b, c = rot
        co, si = a / c, b / c
        planes = []
        for (ux, uy) in ((co, si), (-co, -si), (-si, co), (si, -co)):
            d = ux * cx + uy * cy + half
            planes.append((ux, uy, 0.0, d))
        planes.append((0, 0, 1, z1))
        planes.append((0, 0, -1, -z0))
        r = half * (abs(co) + abs(si)) + 0.5  # analytic half-extent of the rotated square + slack
        return self.add(planes, (cx - r, cy - r, z0), (cx + r, cy + r, z1), texture)

    def ramp(self, x0, x1, y0, y1, z0, z1, along_x=True, texture=TEX_SOLID) -> int:
        """wedge rising from z0 at (x0|y0) to z1 at (x1|y1): one sloped top plane + 4 sides."""
        if along_x:
            run, rise = x1 - x0, z1 - z0
            ln = float(np.sqrt(np.float64(run * run + rise * rise)))
            nx, nz = -rise / ln, run / ln
            slope = (nx, 0.0, nz, nx * x0 + nz * z0)
            planes = [slope, (1, 0, 0, x1), (0, 1, 0, y1), (0, -1, 0, -y0), (0, 0, -1, -z0)]
        else:
            run, rise = y1 - y0, z1 - z0
            ln = float(np.sqrt(np.float64(run * run + rise * rise)))
            ny, nz = -rise / ln, run / ln
            slope = (0.0, ny, nz, ny * y0 + nz * z0)
            planes = [slope, (0, 1, 0, y1), (1, 0, 0, x1), (-1, 0, 0, -x0), (0, 0, -1, -z0)]
        return self.add(planes, (x0, y0, z0), (x1, y1, z1 + 0.5), texture)

    def roof(self, x0, x1, y0, y1, z0, z1, ridge_along_x=True, texture=TEX_SOLID) -> int:
        """pitched roof prism on the rectangle, eaves at z0, ridge at z1: two sloped
        (non-axial) planes + two gable ends + bottom."""
        if ridge_along_x:
            half, rise = (y1 - y0) / 2.0, z1 - z0
            ln = float(np.sqrt(np.float64(half * half + rise * rise)))
            ny, nz = rise / ln, half / ln
            ym = (y0 + y1) / 2.0
            planes = [(0.0, ny, nz, ny * ym + nz * z1), (0.0, -ny, nz, -ny * ym + nz * z1),
                      (1, 0, 0, x1), (-1, 0, 0, -x0), (0, 0, -1, -z0)]
        else:
            half, rise = (x1 - x0) / 2.0, z1 - z0
            ln = float(np.sqrt(np.float64(half * half + rise * rise)))
            nx, nz = rise / ln, half / ln
            xm = (x0 + x1) / 2.0
            planes = [(nx, 0.0, nz, nx * xm + nz * z1), (-nx, 0.0, nz, -nx * xm + nz * z1),
                      (0, 1, 0, y1), (0, -1, 0, -y0), (0, 0, -1, -z0)]
        return self.add(planes, (x0 - 0.5, y0 - 0.5, z0), (x1 + 0.5, y1 + 0.5, z1 + 0.5), texture)

    def __len__(self):
        return len(self.texture)


# --------------------------------------------------------------------------- compiler

def _textures_array(textures) -> np.ndarray:
    arr = np.zeros(len(textures), ibsp.TEXTURE_DT)
    for i, (name, sflags, cflags) in enumerate(textures):
        arr[i] = (name, sflags, cflags)
    return arr


def _tree_numpy(bmin, bmax, lo0, hi0, leaf_max, cost_traverse, cost_brush, max_depth, empty_bonus):
    """reference implementation of the SAH split search (slow: ~150 us of numpy overhead per node)."""
    nodes: list[list] = []       # [axis, c, child0, child1, lo, hi]
    leaves: list[tuple] = []     # (ids, lo, hi)
    stats = {"max_depth": 0, "forced_leaves": 0}

    def make_leaf(ids, lo, hi, depth) -> int:
        leaves.append((np.sort(ids), lo, hi))
        stats["max_depth"] = max(stats["max_depth"], depth)
        return -len(leaves)  # -(leaf_index + 1)

    def best_split(ids, lo, hi):
        n = len(ids)
        ext = hi - lo
        sa = 2.0 * (ext[0] * ext[1] + ext[1] * ext[2] + ext[0] * ext[2])
        best = (np.inf, -1, 0.0, 0, 0)
        for a in range(3):
            mn = np.clip(bmin[ids, a], lo[a], hi[a])
            mx = np.clip(bmax[ids, a], lo[a], hi[a])
            cand = np.unique(np.concatenate((mn, mx)))
            cand = cand[(cand > lo[a]) & (cand < hi[a])]
            if cand.size == 0:
                continue
            n_back = np.searchsorted(np.sort(mn), cand, side="left")          # brushes with min < c
            n_front = n - np.searchsorted(np.sort(mx), cand, side="right")    # brushes with max > c
            o1, o2 = ext[(a + 1) % 3], ext[(a + 2) % 3]
            sa_back = 2.0 * (o1 * o2 + (cand - lo[a]) * (o1 + o2))
            sa_front = 2.0 * (o1 * o2 + (hi[a] - cand) * (o1 + o2))
            cost = cost_traverse + cost_brush * (sa_back * n_back + sa_front * n_front) / sa
            cost = np.where((n_back == 0) | (n_front == 0), cost * empty_bonus, cost)
            i = int(np.argmin(cost))
            if cost[i] < best[0]:
                best = (float(cost[i]), a, float(cand[i]), int(n_back[i]), int(n_front[i]))
        return best

    root_ids = np.nonzero(np.all(bmax > lo0, axis=1) & np.all(bmin < hi0, axis=1))[0]
    work = [(root_ids, lo0, hi0, 0, -1, 0)]  # ids, lo, hi, depth, parent, slot
    while work:
        ids, lo, hi, depth, parent, slot = work.pop()
        n = len(ids)
        if n == 0 or depth >= max_depth:
            if n:
                stats["forced_leaves"] += 1
            ref = make_leaf(ids, lo, hi, depth)
        else:
            cost, axis, c, n_back, n_front = best_split(ids, lo, hi)
            leaf_cost = cost_brush * n
            progress = axis >= 0 and max(n_back, n_front) < n
            if axis < 0 or (cost >= leaf_cost and n <= leaf_max) or (n > leaf_max and not progress and cost >= leaf_cost):
                ref = make_leaf(ids, lo, hi, depth)
            else:
                ref = len(nodes)
                nodes.append([axis, c, 0, 0, lo, hi])
                hi_back = hi.copy(); hi_back[axis] = c
                lo_front = lo.copy(); lo_front[axis] = c
                # push back first so the front child is compiled first (DFS preorder, front first)
                work.append((ids[bmin[ids, axis] < c], lo, hi_back, depth + 1, ref, 1))
                work.append((ids[bmax[ids, axis] > c], lo_front, hi, depth + 1, ref, 0))
        if parent >= 0:
            nodes[parent][2 + slot] = ref
    nn, nl = len(nodes), len(leaves)
    out = {
        "axis": np.array([n[0] for n in nodes], np.int32), "c": np.array([n[1] for n in nodes], np.float64),
        "children": np.array([[n[2], n[3]] for n in nodes], np.int32).reshape(nn, 2),
        "node_lo": np.array([n[4] for n in nodes], np.float64).reshape(nn, 3), "node_hi": np.array([n[5] for n in nodes], np.float64).reshape(nn, 3),
        "leaf_count": np.array([len(l[0]) for l in leaves], np.int64),
        "leaf_lo": np.array([l[1] for l in leaves], np.float64).reshape(nl, 3), "leaf_hi": np.array([l[2] for l in leaves], np.float64).reshape(nl, 3),
        "leaf_ids": np.concatenate([l[0] for l in leaves]).astype(np.int32) if leaves else np.zeros(0, np.int32),
    }
    out["leaf_first"] = np.concatenate(([0], np.cumsum(out["leaf_count"])[:-1])).astype(np.int64) if nl else np.zeros(0, np.int64)
    return out, stats


_bspc = None


def _load_bspc():
    """tools/bspc.c built on first use (gcc -O2 -ffp-contract=off); None if no compiler is available."""
    global _bspc
    if _bspc is not None:
        return _bspc or None
    import ctypes as C
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    src, so = os.path.join(here, "bspc.c"), os.path.join(here, "libbspc.so")
    try:
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            tmp = so + ".tmp%d" % os.getpid()
            subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-o", tmp, src], check=True)
            os.replace(tmp, so)
        lib = C.CDLL(so)
        dp = C.POINTER(C.c_double)
        lib.bspc_build.restype = C.c_void_p
        lib.bspc_build.argtypes = [C.c_int64, dp, dp, dp, dp, C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_double]
        lib.bspc_counts.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        lib.bspc_export.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.bspc_free.argtypes = [C.c_void_p]
        _bspc = lib
    except Exception:
        _bspc = False
    return _bspc or None


def _tree_c(lib, bmin, bmax, lo0, hi0, leaf_max, cost_traverse, cost_brush, max_depth, empty_bonus):
    import ctypes as C
    dp = C.POINTER(C.c_double)
    bmin = np.ascontiguousarray(bmin, np.float64); bmax = np.ascontiguousarray(bmax, np.float64)
    lo0 = np.ascontiguousarray(lo0, np.float64); hi0 = np.ascontiguousarray(hi0, np.float64)
    t = lib.bspc_build(len(bmin), bmin.ctypes.data_as(dp), bmax.ctypes.data_as(dp), lo0.ctypes.data_as(dp), hi0.ctypes.data_as(dp),
                       leaf_max, cost_traverse, cost_brush, max_depth, empty_bonus)
    cnt = (C.c_int64 * 5)()
    lib.bspc_counts(t, cnt)
    nn, nl, ni = int(cnt[0]), int(cnt[1]), int(cnt[2])
    node_i = np.zeros((nn, 3), np.int32); node_f = np.zeros((nn, 7), np.float64)
    leaf_i = np.zeros((nl, 2), np.int64); leaf_f = np.zeros((nl, 6), np.float64); ids = np.zeros(ni, np.int32)
    lib.bspc_export(t, node_i.ctypes.data, node_f.ctypes.data, leaf_i.ctypes.data, leaf_f.ctypes.data, ids.ctypes.data)
    lib.bspc_free(t)
    out = {"axis": node_i[:, 0].copy(), "c": node_f[:, 0].copy(), "children": node_i[:, 1:3].copy(),
           "node_lo": node_f[:, 1:4].copy(), "node_hi": node_f[:, 4:7].copy(),
           "leaf_first": leaf_i[:, 0].copy(), "leaf_count": leaf_i[:, 1].copy(),
           "leaf_lo": leaf_f[:, 0:3].copy(), "leaf_hi": leaf_f[:, 3:6].copy(), "leaf_ids": ids}
    return out, {"max_depth": int(cnt[3]), "forced_leaves": int(cnt[4])}


def compile_bsp(bl: BrushList, world_min, world_max, textures=None, leaf_max: int = 6,
                cost_traverse: float = 1.0, cost_brush: float = 2.0, max_depth: int = 56,
                empty_bonus: float = 0.8, builder: str = "auto") -> ibsp.IBSP:
    """Axial kd-style BSP over the brush AABBs with a surface-area heuristic.

    Every leaf lists *every* brush whose AABB overlaps the leaf cell with positive
    volume (over-inclusion is harmless: the reference clips the segment against
    the brush itself, bsp.cc:120-183).  Nodes are emitted in depth-first order,
    node 0 is the root (the reference starts at node 0, bsp.cc:263).
    ``builder``: "c" (tools/bspc.c), "numpy" (the slow reference implementation), or "auto";
    both produce the same bytes (tests/test_tools.py).
    """
    textures = textures or DEFAULT_TEXTURES
    nb = len(bl)
    bmin = np.asarray(bl.bmin, np.float64).reshape(nb, 3)
    bmax = np.asarray(bl.bmax, np.float64).reshape(nb, 3)
    lo0 = np.asarray(world_min, np.float64)
    hi0 = np.asarray(world_max, np.float64)
    lib = _load_bspc() if builder in ("auto", "c") else None
    if builder == "c" and lib is None:
        raise RuntimeError("tools/bspc.c could not be built")
    if lib is not None:
        tree, stats = _tree_c(lib, bmin, bmax, lo0, hi0, leaf_max, cost_traverse, cost_brush, max_depth, empty_bonus)
    else:
        tree, stats = _tree_numpy(bmin, bmax, lo0, hi0, leaf_max, cost_traverse, cost_brush, max_depth, empty_bonus)

    nn, nl = len(tree["axis"]), len(tree["leaf_count"])
    # split planes join the planes lump in node order (de-duplicated against the brush planes)
    node_plane = np.zeros(nn, np.uint32)
    for i in range(nn):
        normal = [0.0, 0.0, 0.0]
        normal[int(tree["axis"][i])] = 1.0
        node_plane[i] = bl.plane(normal[0], normal[1], normal[2], float(tree["c"][i]))
    children = tree["children"]
    leaf_first, leaf_count = tree["leaf_first"], tree["leaf_count"]
    leaf_lo, leaf_hi = tree["leaf_lo"], tree["leaf_hi"]
    node_lo, node_hi = tree["node_lo"], tree["node_hi"]
    if nn == 0:
        # the reference always starts at node 0 (bsp.cc:263): wrap the lone leaf in a node
        node_plane = np.array([bl.plane(1.0, 0.0, 0.0, float(lo0[0]))], np.uint32)
        children = np.array([[-1, -(nl + 1)]], np.int32)
        node_lo, node_hi = lo0[None, :], hi0[None, :]
        leaf_first = np.concatenate((leaf_first, [int(leaf_count.sum())])).astype(np.int64)
        leaf_count = np.concatenate((leaf_count, [0])).astype(np.int64)
        leaf_lo = np.vstack((leaf_lo, lo0[None, :])); leaf_hi = np.vstack((leaf_hi, hi0[None, :]))
        nn, nl = 1, nl + 1

    node_arr = np.zeros(nn, ibsp.NODE_DT)
    node_arr["plane"] = node_plane
    node_arr["children"] = children
    node_arr["mins"] = np.floor(node_lo).astype(np.int32)
    node_arr["maxs"] = np.ceil(node_hi).astype(np.int32)
    leaf_arr = np.zeros(nl, ibsp.LEAF_DT)
    leaf_arr["mins"] = np.floor(leaf_lo).astype(np.int32)
    leaf_arr["maxs"] = np.ceil(leaf_hi).astype(np.int32)
    leaf_arr["first_brush"] = leaf_first.astype(np.uint32)
    leaf_arr["num_brushes"] = leaf_count.astype(np.uint32)
    leaf_brushes = tree["leaf_ids"].astype(np.uint32)
    off = int(leaf_count.sum())

    plane_arr = np.zeros(len(bl.planes), ibsp.PLANE_DT)
    if bl.planes:
        p = np.asarray(bl.planes, np.float32)
        plane_arr["n"] = p[:, :3]
        plane_arr["d"] = p[:, 3]
    brush_arr = np.zeros(nb, ibsp.BRUSH_DT)
    brush_arr["first_side"] = np.asarray(bl.first_side, np.uint32)
    brush_arr["num_sides"] = np.asarray(bl.num_sides, np.uint32)
    brush_arr["texture"] = np.asarray(bl.texture, np.int32)
    side_arr = np.zeros(len(bl.side_planes), ibsp.BRUSHSIDE_DT)
    side_arr["plane"] = np.asarray(bl.side_planes, np.uint32)
    side_arr["texture"] = np.repeat(np.asarray(bl.texture, np.int32), np.asarray(bl.num_sides, np.int64)) if nb else 0

    meta = dict(stats)
    meta.update(world_min=[float(v) for v in lo0], world_max=[float(v) for v in hi0],
                leaf_brush_refs=int(off), avg_leaf_brushes=float(off) / max(1, nl))
    return ibsp.IBSP(_textures_array(textures), plane_arr, node_arr, leaf_arr, leaf_brushes, brush_arr, side_arr, meta)


# --------------------------------------------------------------------------- KAT maps

def _raw_map(textures, planes, nodes, leaves, leaf_brushes, brushes, sides) -> ibsp.IBSP:
    tex = _textures_array(textures)
    pl = np.zeros(len(planes), ibsp.PLANE_DT)
    for i, (nx, ny, nz, d) in enumerate(planes):
        pl[i] = ((np.float32(nx), np.float32(ny), np.float32(nz)), np.float32(d))
    nd = np.zeros(len(nodes), ibsp.NODE_DT)
    for i, (p, c0, c1) in enumerate(nodes):
        nd[i] = (p, (c0, c1), (0, 0, 0), (0, 0, 0))
    lf = np.zeros(len(leaves), ibsp.LEAF_DT)
    for i, (first, num) in enumerate(leaves):
        lf[i] = (0, 0, (0, 0, 0), (0, 0, 0), 0, 0, first, num)
    br = np.zeros(len(brushes), ibsp.BRUSH_DT)
    for i, b in enumerate(brushes):
        br[i] = b
    sd = np.zeros(len(sides), ibsp.BRUSHSIDE_DT)
    for i, s in enumerate(sides):
        sd[i] = s
    return ibsp.IBSP(tex, pl, nd, lf, np.asarray(leaf_brushes, np.uint32), br, sd, {"world_min": [-64] * 3, "world_max": [64] * 3})


def kat0() -> ibsp.IBSP:
    """one node, two leaves, one 32^3 box brush."""
    planes = [(1, 0, 0, 16), (-1, 0, 0, 16), (0, 1, 0, 16), (0, -1, 0, 16), (0, 0, 1, 16), (0, 0, -1, 16), (1, 0, 0, 0)]
    return _raw_map(
        textures=[(b"solid", 0, 1), (b"empty", 0, 0)], planes=planes,
        nodes=[(6, -1, -2)], leaves=[(0, 1), (1, 1)], leaf_brushes=[0, 0],
        brushes=[(0, 6, 0)], sides=[(i, 0) for i in range(6)])


def _kat_box(x0, x1, y0, y1, z0, z1):
    return [(1, 0, 0, x1), (-1, 0, 0, -x0), (0, 1, 0, y1), (0, -1, 0, -y0), (0, 0, 1, z1), (0, 0, -1, -z0)]


def kat1() -> ibsp.IBSP:
    """two nodes, three leaves, six brushes (one non-solid, one wedge)."""
    s = float(np.float32(0.70710678))
    brush_planes = [
        _kat_box(-16, 16, -16, 16, -16, 16),      # bA
        _kat_box(100, 120, -16, 16, -16, 16),     # bB
        _kat_box(90, 130, -8, 8, -8, 8),          # bC
        _kat_box(200, 220, -16, 16, -16, 16),     # bN (non-solid)
        _kat_box(230, 250, -16, 16, -16, 16),     # bD
        [(s, s, 0, 300 * s), (-s, s, 0, 20), (0, -1, 0, -150), (0, 0, 1, 16), (0, 0, -1, 16)],  # bR
    ]
    brush_tex = [0, 0, 0, 1, 0, 0]
    planes, brushes, sides = [], [], []
    for bp, tex in zip(brush_planes, brush_tex):
        brushes.append((len(sides), len(bp), tex))
        for p in bp:
            sides.append((len(planes), tex))
            planes.append(p)
    nx0 = len(planes); planes.append((1, 0, 0, 0))
    ny50 = len(planes); planes.append((0, 1, 0, 50))
    # n0{nx0, [1, -1]}  n1{ny50, [-2, -3]};  L0=[bA]  L1=[bR]  L2=[bB,bC,bN,bD]
    m = _raw_map(
        textures=[(b"solid", 0, 1), (b"water", 0, 0x20)], planes=planes,
        nodes=[(nx0, 1, -1), (ny50, -2, -3)], leaves=[(0, 1), (1, 1), (2, 4)],
        leaf_brushes=[0, 5, 1, 2, 3, 4], brushes=brushes, sides=sides)
    m.meta.update(world_min=[-80, -60, -32], world_max=[320, 230, 32])
    return m


# (start, end, hit, t_bits, pos) captured from the UNMODIFIED reference by the survey
KAT0_VECTORS = [
    ((64, 0, 0), (-64, 0, 0), 1, 0x3EC00000, (16, 0, 0)),
    ((-64, 1, 2), (64, 3, 4), 1, 0x3EC00000, (-16, 1.75, 2.75)),
    ((64, 0, 40), (-64, 0, 40), 0, 0x3F800000, (0, 0, 0)),
    ((1, 0, 0), (64, 0, 0), 0, 0x3F800000, (0, 0, 0)),
    ((1, 0, 0), (-64, 0, 0), 0, 0x3F800000, (0, 0, 0)),
    ((64, 0, 0), (4, 0, 0), 1, 0x3F4CCCCD, (16, 0, 0)),
    ((16, 0, 0), (-64, 0, 0), 0, 0x3F800000, (0, 0, 0)),
    ((64, 0, 0), (16, 0, 0), 1, 0x3F800000, (16, 0, 0)),
    ((40, 40, 40), (40, 40, 40), 0, 0x3F800000, (0, 0, 0)),
    ((50, 40, 30), (-50, -40, -30), 1, 0x3EAE147B, (16, 12.8, 9.6)),
    ((0, 64, 0), (0, -64, 0), 1, 0x3EC00000, (0, 16, 0)),
]
KAT1_VECTORS = [
    ((64, 0, 0), (-64, 0, 0), 1, 0x3F000000, (0, 0, 0)),            # A: hit reported at tmin
    ((-64, 0, 0), (64, 0, 0), 1, 0x3EC00000, (-16, 0, 0)),          # B
    ((180, 0, 0), (60, 0, 0), 1, 0x3ED55555, (130, 0, 0)),          # C: min over bB/bC
    ((180, 12, 0), (60, 12, 0), 1, 0x3F000000, (120, 12, 0)),       # D
    ((190, 0, 0), (260, 0, 0), 1, 0x3F124925, (230, 0, 0)),         # E: through non-solid bN
    ((190, 0, 0), (225, 0, 0), 0, 0x3F800000, (0, 0, 0)),           # F
    ((-10, 45, 0), (207.5, 210, 0), 1, 0x3F22E8BA, None),           # G: u<tmin flip at n1
    ((200, 157, 3), (100, 157, -5), 1, 0x3F11EB85, None),           # H
    ((100, 158, 0), (160, 156, 1), 1, 0x3EF564F6, None),            # I
    ((110, 40, 0), (110, -40, 0), 1, 0x3E99999A, (110, 16, 0)),     # J
    ((10, 0, 0), (-64, 0, 0), 0, 0x3F800000, (0, 0, 0)),            # K: starts inside bA
    ((140, 36, 0), (100, -4, 0), 1, 0x3F000000, (120, 16, 0)),      # L: two entering planes tie
]


# --------------------------------------------------------------------------- workload maps

ROTATIONS = [(3, 4, 5), (4, 3, 5), (5, 12, 13), (12, 5, 13), (8, 15, 17), (15, 8, 17), (7, 24, 25), (20, 21, 29)]


def rooms8(seed: int = 1) -> ibsp.IBSP:
    """C1/C2 map: 4x2 hollow rooms (1024 x 2048 x 1024 each) inside x,y in [-2048,2048],
    z in [-512,512]; 32-unit walls with door gaps, ~50 crates / columns, non-axial pillars
    and ramps per room, a few non-solid volumes.  ~500 brushes."""
    rng = Rng(seed)
    bl = BrushList()
    W, T = 2048, 32
    z0, z1 = -512, 512
    # outer shell: floor, ceiling, 4 walls
    bl.box(-W, W, -W, W, z0, z0 + T)
    bl.box(-W, W, -W, W, z1 - T, z1)
    bl.box(-W, -W + T, -W, W, z0 + T, z1 - T)
    bl.box(W - T, W, -W, W, z0 + T, z1 - T)
    bl.box(-W + T, W - T, -W, -W + T, z0 + T, z1 - T)
    bl.box(-W + T, W - T, W - T, W, z0 + T, z1 - T)
    door_w, door_h = 192, 320
    # interior walls along x (between the 4 columns of rooms), each with one door per room row
    for k in (-1024, 0, 1024):
        for (ya, yb) in ((-W + T, -T // 2), (T // 2, W - T)):
            dc = rng.randint(ya + 300, yb - 300) // 16 * 16
            bl.box(k - T // 2, k + T // 2, ya, dc - door_w // 2, z0 + T, z1 - T)
            bl.box(k - T // 2, k + T // 2, dc + door_w // 2, yb, z0 + T, z1 - T)
            bl.box(k - T // 2, k + T // 2, dc - door_w // 2, dc + door_w // 2, z0 + T + door_h, z1 - T)
    # the wall along y=0 between the two rows, one door per column
    for c in range(4):
        xa = -W + T if c == 0 else -W + c * 1024 + T // 2
        xb = W - T if c == 3 else -W + (c + 1) * 1024 - T // 2
        dc = rng.randint(xa + 250, xb - 250) // 16 * 16
        bl.box(xa, dc - door_w // 2, -T // 2, T // 2, z0 + T, z1 - T)
        bl.box(dc + door_w // 2, xb, -T // 2, T // 2, z0 + T, z1 - T)
        bl.box(dc - door_w // 2, dc + door_w // 2, -T // 2, T // 2, z0 + T + door_h, z1 - T)
    for k in (-1024, 0, 1024):  # posts where the interior walls cross
        bl.box(k - T // 2, k + T // 2, -T // 2, T // 2, z0 + T, z1 - T)
    # room contents
    for c in range(4):
        for r in range(2):
            rx0, rx1 = -W + c * 1024 + 48, -W + (c + 1) * 1024 - 48
            ry0, ry1 = (-W + 48, -48) if r == 0 else (48, W - 48)
            fz = z0 + T
            for _ in range(46):
                sx, sy = rng.randint(16, 96), rng.randint(16, 96)
                sz = rng.randint(32, 384)
                x = rng.randint(rx0, rx1 - sx); y = rng.randint(ry0, ry1 - sy)
                zb = fz if rng.chance(3, 4) else rng.randint(fz + 64, z1 - T - sz - 1)
                tex = TEX_SOLID_PLUS if rng.chance(1, 10) else TEX_SOLID
                bl.box(x, x + sx, y, y + sy, zb, zb + sz, tex)
            for _ in range(5):
                half = rng.randint(12, 40)
                x = rng.randint(rx0 + 64, rx1 - 64); y = rng.randint(ry0 + 64, ry1 - 64)
                bl.pillar(x, y, half, fz, z1 - T, ROTATIONS[rng.randint(0, len(ROTATIONS) - 1)])
            for _ in range(2):
                sx, sy, sz = rng.randint(96, 256), rng.randint(64, 160), rng.randint(48, 200)
                x = rng.randint(rx0, rx1 - sx); y = rng.randint(ry0, ry1 - sy)
                bl.ramp(x, x + sx, y, y + sy, fz, fz + sz, along_x=rng.chance(1, 2))
            # one non-solid volume (water / clip) per room: must be skipped by the content test
            sx, sy = rng.randint(128, 320), rng.randint(128, 320)
            x = rng.randint(rx0, rx1 - sx); y = rng.randint(ry0, ry1 - sy)
            bl.box(x, x + sx, y, y + sy, fz, fz + rng.randint(32, 128), TEX_WATER if rng.chance(1, 2) else TEX_NONSOLID)
    m = compile_bsp(bl, (-W, -W, z0), (W, W, z1))
    m.meta["name"] = "rooms8"
    return m


def _city(grid: int, pitch: int, world_xy: int, world_z: int, hmin: int, hmax: int, roof_num: int,
          roof_den: int, seed: int, name: str) -> ibsp.IBSP:
    rng = Rng(seed)
    bl = BrushList()
    W, Z, T = world_xy, world_z, 64
    bl.box(-W, W, -W, W, -T, 0)                 # ground
    bl.box(-W, W, -W, W, Z, Z + T)              # sky lid
    bl.box(-W - T, -W, -W, W, 0, Z)
    bl.box(W, W + T, -W, W, 0, Z)
    bl.box(-W, W, -W - T, -W, 0, Z)
    bl.box(-W, W, W, W + T, 0, Z)
    half = grid * pitch // 2
    for gy in range(grid):
        for gx in range(grid):
            bx0, by0 = -half + gx * pitch, -half + gy * pitch
            margin = pitch // 8
            fx = rng.randint(pitch // 2, pitch - 2 * margin)
            fy = rng.randint(pitch // 2, pitch - 2 * margin)
            x0 = bx0 + margin + rng.randint(0, pitch - 2 * margin - fx)
            y0 = by0 + margin + rng.randint(0, pitch - 2 * margin - fy)
            # skewed heights: mostly low, a few towers
            u = rng.randint(0, 1023)
            h = hmin + (hmax - hmin) * u * u * u // (1023 * 1023 * 1023)
            h = max(hmin, h // 8 * 8)
            bl.box(x0, x0 + fx, y0, y0 + fy, 0, h)
            if rng.chance(roof_num, roof_den):
                rise = rng.randint(max(8, min(fx, fy) // 4), max(9, min(fx, fy) // 2))
                bl.roof(x0, x0 + fx, y0, y0 + fy, h, h + rise, ridge_along_x=rng.chance(1, 2))
    m = compile_bsp(bl, (-W - T, -W - T, -T), (W + T, W + T, Z + T))
    m.meta["name"] = name
    # segments are drawn inside the playable volume, not inside the shell slabs
    m.meta["seg_min"] = [-W, -W, 0]
    m.meta["seg_max"] = [W, W, Z]
    return m


def city10k(seed: int = 3) -> ibsp.IBSP:
    """C3/C4 map: 90x90 blocks (pitch 352) of box buildings, 25% with pitched roofs, in
    x,y in [-16384,16384], z in [0,4096]; ground + shell.  ~10.1k brushes."""
    return _city(grid=90, pitch=352, world_xy=16384, world_z=4096, hmin=64, hmax=3000,
                 roof_num=1, roof_den=4, seed=seed, name="city10k")


def mega200k(seed: int = 5) -> ibsp.IBSP:
    """C5 map: 447x447 blocks (pitch 64) of low box buildings, 2% roofs.  ~204k brushes."""
    return _city(grid=447, pitch=64, world_xy=14336, world_z=1024, hmin=16, hmax=768,
                 roof_num=1, roof_den=50, seed=seed, name="mega200k")


def _brush_corners(bmin, bmax):
    xs = np.array([[bmin[0], bmax[0]][i & 1] for i in range(8)])
    ys = np.array([[bmin[1], bmax[1]][(i >> 1) & 1] for i in range(8)])
    zs = np.array([[bmin[2], bmax[2]][(i >> 2) & 1] for i in range(8)])
    return np.stack((xs, ys, zs), axis=1)


def tilted(seed: int = 9, n_brushes: int = 260, leaf_max: int = 1, max_depth: int = 60) -> ibsp.IBSP:
    """Q3-style tree: node planes are taken from BRUSH SIDES, so most of them are non-axial (the kd
    compiler above only ever emits axial splits, real q3map output does not), the tree is unbalanced
    and deeper than 32 levels (exercises the 64-entry traversal stack).  Brushes: rotated pillars,
    ramps, roofs and boxes scattered in a 2048^3 cube; classification against a split plane uses the
    eight AABB corners (conservative)."""
    rng = Rng(seed)
    bl = BrushList()
    W = 1024
    for k in range(n_brushes):
        cx, cy, cz = rng.randint(-W + 100, W - 100), rng.randint(-W + 100, W - 100), rng.randint(-W + 100, W - 100)
        kind = rng.randint(0, 5) % 4              # pillars twice as likely: most side planes non-axial
        tex = TEX_NONSOLID if rng.chance(1, 12) else (TEX_SOLID_PLUS if rng.chance(1, 8) else TEX_SOLID)
        if kind == 0:
            bl.pillar(cx, cy, rng.randint(8, 60), cz - rng.randint(8, 90), cz + rng.randint(8, 90), ROTATIONS[rng.randint(0, len(ROTATIONS) - 1)], tex)
        elif kind == 1:
            sx, sy, sz = rng.randint(20, 120), rng.randint(20, 120), rng.randint(10, 100)
            bl.ramp(cx, cx + sx, cy, cy + sy, cz, cz + sz, along_x=rng.chance(1, 2), texture=tex)
        elif kind == 2:
            sx, sy, sz = rng.randint(20, 120), rng.randint(20, 120), rng.randint(10, 80)
            bl.roof(cx, cx + sx, cy, cy + sy, cz, cz + sz, ridge_along_x=rng.chance(1, 2), texture=tex)
        else:
            sx, sy, sz = rng.randint(10, 90), rng.randint(10, 90), rng.randint(10, 90)
            bl.box(cx, cx + sx, cy, cy + sy, cz, cz + sz, tex)
    nb = len(bl)
    bmin = np.asarray(bl.bmin, np.float64); bmax = np.asarray(bl.bmax, np.float64)
    corners = np.stack([_brush_corners(bmin[i], bmax[i]) for i in range(nb)])          # (nb, 8, 3)
    planes64 = np.asarray(bl.planes, np.float64)
    first = np.asarray(bl.first_side); nsides = np.asarray(bl.num_sides); side_planes = np.asarray(bl.side_planes)

    nodes, leaves = [], []
    stats = {"max_depth": 0}

    def leaf(ids, depth):
        leaves.append(np.sort(ids)); stats["max_depth"] = max(stats["max_depth"], depth)
        return -len(leaves)

    work = [(np.arange(nb), 0, -1, 0)]
    while work:
        ids, depth, parent, slot = work.pop()
        ref = None
        if len(ids) <= leaf_max or depth >= max_depth:
            ref = leaf(ids, depth)
        else:
            chosen = None
            lopsided = rng.chance(7, 8)               # most splits peel off as little as possible: a deep, unbalanced tree
            for _ in range(24):                       # a few random candidate side planes (non-axial preferred)
                b = int(ids[rng.randint(0, len(ids) - 1)])
                pidx = int(side_planes[first[b] + rng.randint(0, int(nsides[b]) - 1)])
                n, d = planes64[pidx, :3], planes64[pidx, 3]
                if np.abs(n).max() == 1.0 and rng.chance(3, 4):
                    continue
                dist = corners[ids] @ n - d            # (k, 8)
                in_front = dist.max(axis=1) > 1e-6
                in_back = dist.min(axis=1) < -1e-6
                nf, nbk = int(in_front.sum()), int(in_back.sum())
                if max(nf, nbk) < len(ids):
                    score = max(nf, nbk) + (nf + nbk - len(ids))
                    if lopsided:
                        score = -max(nf, nbk)
                    if chosen is None or score < chosen[0]:
                        chosen = (score, pidx, ids[in_front], ids[in_back | ~in_front])
            if chosen is None:
                ref = leaf(ids, depth)
            else:
                ref = len(nodes)
                nodes.append([chosen[1], 0, 0])
                work.append((chosen[3], depth + 1, ref, 1))
                work.append((chosen[2], depth + 1, ref, 0))
        if parent >= 0:
            nodes[parent][1 + slot] = ref
    node_arr = np.zeros(len(nodes), ibsp.NODE_DT)
    for i, (pidx, c0, c1) in enumerate(nodes):
        node_arr[i] = (pidx, (c0, c1), (0, 0, 0), (0, 0, 0))
    node_arr[0]["mins"] = (-W - 150,) * 3            # q3map writes node bounds; the root's are the world box
    node_arr[0]["maxs"] = (W + 150,) * 3
    leaf_arr = np.zeros(len(leaves), ibsp.LEAF_DT)
    off = 0
    for i, ids in enumerate(leaves):
        leaf_arr[i]["first_brush"] = off; leaf_arr[i]["num_brushes"] = len(ids); off += len(ids)
    plane_arr = np.zeros(len(bl.planes), ibsp.PLANE_DT)
    p32 = np.asarray(bl.planes, np.float32)
    plane_arr["n"] = p32[:, :3]; plane_arr["d"] = p32[:, 3]
    brush_arr = np.zeros(nb, ibsp.BRUSH_DT)
    brush_arr["first_side"] = first; brush_arr["num_sides"] = nsides; brush_arr["texture"] = np.asarray(bl.texture, np.int32)
    side_arr = np.zeros(len(bl.side_planes), ibsp.BRUSHSIDE_DT)
    side_arr["plane"] = side_planes
    side_arr["texture"] = np.repeat(np.asarray(bl.texture, np.int32), nsides)
    meta = {"name": "tilted", "max_depth": stats["max_depth"], "world_min": [-W - 150.0] * 3, "world_max": [W + 150.0] * 3}
    return ibsp.IBSP(_textures_array(DEFAULT_TEXTURES), plane_arr, node_arr, leaf_arr,
                     np.concatenate(leaves).astype(np.uint32), brush_arr, side_arr, meta)


def q3style(seed: int = 12) -> ibsp.IBSP:
    from tools import mapgen_q3
    return mapgen_q3.q3style(seed)


MAPS = {"kat0": kat0, "kat1": kat1, "rooms8": rooms8, "city10k": city10k, "mega200k": mega200k, "tilted": tilted, "q3style": q3style}


def cache_dir() -> str:
    d = os.environ.get("BSP_B200_CACHE", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), ".mapcache"))
    os.makedirs(d, exist_ok=True)
    return d


def _source_digest() -> str:
    h = hashlib.sha256()
    here = os.path.dirname(os.path.abspath(__file__))
    for fn in (__file__, ibsp.__file__, os.path.join(here, "bspc.c"), os.path.join(here, "mapgen_q3.py")):
        with open(fn, "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:12]


def get_map(name: str, use_cache: bool = True):
    """returns (IBSP, path-to-.bsp).  Generated once per source revision and cached on disk."""
    path = os.path.join(cache_dir(), f"{name}-{_source_digest()}.bsp")
    metap = path + ".meta.npy"
    if use_cache and os.path.exists(path) and os.path.exists(metap):
        m = ibsp.read(path)
        m.meta.update(np.load(metap, allow_pickle=True).item())
        return m, path
    m = MAPS[name]()
    for old in os.listdir(cache_dir()):           # files of earlier source revisions
        if old.startswith(name + "-") and not old.startswith(os.path.basename(path)):
            try:
                os.remove(os.path.join(cache_dir(), old))
            except OSError:
                pass
    ibsp.write(m, path + ".tmp%d" % os.getpid(), num_lumps=int(m.meta.get("num_lumps", ibsp.NUM_LUMPS)))
    os.replace(path + ".tmp%d" % os.getpid(), path)
    np.save(metap + ".tmp%d.npy" % os.getpid(), np.array(m.meta, dtype=object), allow_pickle=True)
    os.replace(metap + ".tmp%d.npy" % os.getpid(), metap)
    return m, path


def seg_bounds(m: ibsp.IBSP):
    lo = m.meta.get("seg_min", m.meta["world_min"])
    hi = m.meta.get("seg_max", m.meta["world_max"])
    return np.asarray(lo, np.float32), np.asarray(hi, np.float32)


if __name__ == "__main__":
    import sys
    import time
    for nm in sys.argv[1:] or ["kat0", "kat1", "rooms8"]:
        t0 = time.time()
        mm, p = get_map(nm, use_cache=False)
        print(nm, mm.counts(), {k: v for k, v in mm.meta.items() if not isinstance(v, list)}, f"{time.time() - t0:.1f}s", p)
