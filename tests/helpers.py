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
            lib.hostsim_create.restype = C.c_void_p
            lib.hostsim_create.argtypes = [C.c_void_p]
            lib.hostsim_error.restype = C.c_char_p
            lib.hostsim_destroy.argtypes = [C.c_void_p]
            lib.hostsim_layout.argtypes = [C.c_void_p, C.c_void_p]
            lib.hostsim_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
            lib.hostsim_trace_phased.argtypes = lib.hostsim_trace.argtypes + [C.c_uint32, C.c_uint32]
            lib.hostsim_leaf.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]
            cls._lib = lib
        return cls._lib

    def __init__(self, m):
        from bsp_b200.api import MapDesc
        self.l = self.lib()
        desc = MapDesc()
        self._keep = []
        for nm in ("textures", "planes", "nodes", "leaves", "leaf_brushes", "brushes", "brush_sides"):
            arr = np.ascontiguousarray(getattr(m, nm))
            self._keep.append(arr)
            setattr(desc, nm, arr.ctypes.data if arr.size else None)
            setattr(desc, "num_" + nm, len(arr))
        self.h = self.l.hostsim_create(C.byref(desc))
        if not self.h:
            raise ValueError(self.l.hostsim_error().decode())

    def layout(self):
        out = np.zeros(6, np.uint64)
        self.l.hostsim_layout(self.h, out.ctypes.data)
        return dict(zip(("nodes", "refs", "sides", "depth", "staged_bytes", "total_bytes"), (int(v) for v in out)))

    def trace(self, segs, budgets=None):
        from bsp_b200.api import RESULT_DT
        segs = np.ascontiguousarray(segs, np.float32)
        res = np.zeros(len(segs), RESULT_DT)
        pos = np.zeros((len(segs), 4), np.float32)
        if budgets is None:
            self.l.hostsim_trace(self.h, segs.ctypes.data, 8, 4, len(segs), res.ctypes.data, pos.ctypes.data)
        else:
            self.l.hostsim_trace_phased(self.h, segs.ctypes.data, 8, 4, len(segs), res.ctypes.data, pos.ctypes.data, budgets[0], budgets[1])
        return res, pos

    def leaf(self, pts):
        pts = np.ascontiguousarray(pts, np.float32)
        out = np.zeros(len(pts), np.int32)
        self.l.hostsim_leaf(self.h, pts.ctypes.data, pts.shape[1], len(pts), out.ctypes.data)
        return out
