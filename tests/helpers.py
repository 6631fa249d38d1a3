"""shared helpers for the test-suite (no pytest fixtures here)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def load_golden(name):
    """returns (IBSP, segs, hit, t, pos) of tests/golden/<name>.npz (reference outputs)."""
    from tools import ibsp
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return ibsp.from_bytes(z["bsp"].tobytes()), z["segs"], z["hit"], z["t"], z["pos"]


def check_against_extreme_golden(hit, t, pos, g_hit, g_t, g_pos, what):
    """hit and t bit for bit; pos bit for bit where the reference's is finite, NaN where it is NaN (payloads are the host's)"""
    assert np.array_equal(np.asarray(hit, np.uint8), g_hit), what
    assert np.array_equal(bits(t), bits(g_t)), what
    finite = np.isfinite(g_pos).all(axis=1)
    assert np.array_equal(bits(pos[finite]), bits(g_pos[finite])), what
    assert np.array_equal(np.asarray(pos), g_pos, equal_nan=True), what


def mixed_segments(m, n_uniform=200_000, n_long=200_000, n_short=100_000, n_edge=16384, seed=21):
    from tools import mapgen, seggen
    lo, hi = mapgen.seg_bounds(m)
    diag = float(np.sqrt(((hi - lo).astype(np.float64) ** 2).sum()))
    return np.vstack((
        seggen.edge_cases(m, n_edge, seed),
        seggen.uniform(seed, 0, n_uniform, lo, hi),
        seggen.long_rays(seed + 1, 0, n_long, lo, hi, 0.25 * diag, diag),
        seggen.long_rays(seed + 2, 0, n_short, lo, hi, 1.0, 32.0),
    ))


def grid_adversarial_segments(m, n, seed=5):
    """segments built to break the start grid if its margins were wrong: endpoints on / a few ulps or a
    fraction of a unit beside start-grid cell borders, lengths from 0 to a few cells, and a third of them
    hugging axial node planes on either side."""
    rs = np.random.Generator(np.random.PCG64(seed))
    lo = np.array(m.nodes["mins"][0], np.float64); hi = np.array(m.nodes["maxs"][0], np.float64)
    ext = hi - lo
    assert (ext > 0).all(), "map has no world bounds in its root node"
    cell = np.cbrt(ext.prod() / 1048576.0)
    dim = np.clip(np.ceil(ext / cell), 1, 256)
    cs = ext / dim
    p = lo + ext * rs.random((n, 3))
    ax = rs.integers(0, 3, n)
    k = np.round((p[np.arange(n), ax] - lo[ax]) / cs[ax])
    p[np.arange(n), ax] = lo[ax] + k * cs[ax] + rs.choice([0.0, 1e-3, -1e-3, 1e-5, -1e-5, 0.4, -0.4], n)
    d = rs.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1)[:, None]
    e = p + d * rs.choice([0.0, 1e-3, 0.5, 4.0, 30.0, 200.0], n)[:, None]
    segs = np.zeros((n, 8), np.float32)
    segs[:, 0:3] = p; segs[:, 4:7] = e
    planes = m.planes[m.nodes["plane"]]
    axial = np.nonzero((np.abs(planes["n"]) == 1.0).any(axis=1))[0]
    if len(axial):
        cnt = n // 3
        sel = rs.choice(axial, cnt)
        a = np.argmax(np.abs(planes["n"][sel]), axis=1)
        val = planes["d"][sel] * planes["n"][sel][np.arange(cnt), a]
        off = [0.0, 1e-3, -1e-3, 0.3, -0.3, 1.1, -1.1]
        segs[np.arange(cnt), a] = (val + rs.choice(off, cnt)).astype(np.float32)
        segs[np.arange(cnt), 4 + a] = (val + rs.choice(off, cnt)).astype(np.float32)
    return segs


def extreme_segments(n, extent, seed=11):
    """uniform segments in [-extent, extent]^3; every third one gets 1-3 special coordinates (inf, NaN, huge, denormal, +-0)"""
    rs = np.random.Generator(np.random.PCG64(seed))
    segs = np.zeros((n, 8), np.float32)
    segs[:, 0:3] = rs.uniform(-extent, extent, (n, 3)); segs[:, 4:7] = rs.uniform(-extent, extent, (n, 3))
    specials = np.array([np.inf, -np.inf, 3.0e38, -3.0e38, 1.0e-40, -1.0e-40, 1.0e-45, 0.0, -0.0, np.nan, 1.0e30, -1.0e30], np.float32)
    for k in range(0, n, 3):
        for _ in range(int(rs.integers(1, 4))):
            segs[k, int(rs.choice([0, 1, 2, 4, 5, 6]))] = specials[int(rs.integers(0, len(specials)))]
    return segs


class HostSim:
    """tests/hostsim: a g++ build of bsp_b200/csrc/trace_core.h + flatten.cc (test-only)."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            so = os.path.join(ROOT, "tests", "hostsim", "libhostsim.so")
            srcs = [os.path.join(ROOT, "tests", "hostsim", "hostsim.cc"), os.path.join(ROOT, "bsp_b200", "csrc", "flatten.cc")]
            deps = srcs + [os.path.join(ROOT, "bsp_b200", "csrc", f) for f in ("trace_core.h", "map_layout.h", "flatten.h")]
            if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
                subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-Wall", "-Wextra", "-fPIC", "-shared", "-o", so] + srcs, check=True)
            lib = C.CDLL(so)
            lib.hostsim_create_opts.restype = C.c_void_p
            lib.hostsim_create_opts.argtypes = [C.c_void_p, C.c_int, C.c_int64]
            lib.hostsim_error.restype = C.c_char_p
            lib.hostsim_destroy.argtypes = [C.c_void_p]
            lib.hostsim_layout.argtypes = [C.c_void_p, C.c_void_p]
            lib.hostsim_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
            lib.hostsim_trace_phased.argtypes = lib.hostsim_trace.argtypes + [C.c_uint32, C.c_uint32, C.c_int]
            lib.hostsim_trace_slots.argtypes = lib.hostsim_trace_phased.argtypes
            lib.hostsim_leaf.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]
            cls._lib = lib
        return cls._lib

    def __init__(self, m, general_planes_only=False, grid_cells=-1):
        """general_planes_only: flatten every node plane as a general plane (no axis shortcut); grid_cells as bspgpu_create_opts"""
        from bsp_b200.api import MapDesc
        self.l = self.lib()
        desc = MapDesc()
        self._keep = []
        for nm in ("textures", "planes", "nodes", "leaves", "leaf_brushes", "brushes", "brush_sides"):
            arr = np.ascontiguousarray(getattr(m, nm))
            self._keep.append(arr)
            setattr(desc, nm, arr.ctypes.data if arr.size else None)
            setattr(desc, "num_" + nm, len(arr))
        self.h = self.l.hostsim_create_opts(C.byref(desc), 1 if general_planes_only else 0, grid_cells)
        if not self.h:
            raise ValueError(self.l.hostsim_error().decode())

    def layout(self):
        out = np.zeros(9, np.uint64)
        self.l.hostsim_layout(self.h, out.ctypes.data)
        return dict(zip(("nodes", "refs", "sides", "depth", "staged_bytes", "total_bytes", "grid_cells", "general_planes", "unused"),
                        (int(v) for v in out)))

    def trace(self, segs, budgets=None, grid=False, split_sides=False, split_stack=False, walk="slots"):
        """budgets=None: trace_one (the simple form); else the resumable walk with those step budgets -- walk="slots": over child
        slots, what the shipped kernels run; walk="nodes": over node records"""
        from bsp_b200.api import RESULT_DT
        segs = np.ascontiguousarray(segs, np.float32)
        res = np.zeros(len(segs), RESULT_DT)
        pos = np.zeros((len(segs), 4), np.float32)
        if budgets is None:
            self.l.hostsim_trace(self.h, segs.ctypes.data, 8, 4, len(segs), res.ctypes.data, pos.ctypes.data)
        else:
            fn = self.l.hostsim_trace_slots if walk == "slots" else self.l.hostsim_trace_phased
            fn(self.h, segs.ctypes.data, 8, 4, len(segs), res.ctypes.data, pos.ctypes.data, budgets[0], budgets[1], (1 if grid else 0) | (2 if split_sides else 0) | (4 if split_stack else 0))
        return res, pos

    def leaf(self, pts):
        pts = np.ascontiguousarray(pts, np.float32)
        out = np.zeros(len(pts), np.int32)
        self.l.hostsim_leaf(self.h, pts.ctypes.data, pts.shape[1], len(pts), out.ctypes.data)
        return out
