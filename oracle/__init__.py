"""CPU checkers for the BSP segment-trace path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  Nothing under ``bsp_b200/``
does; the product path fails loudly without its CUDA library instead of falling
back to anything here.

Two checkers:

* :class:`Reference` -- ``oracle/_ref/libbspref.so``: the reference's own ``bsp.cc``,
  ``memory_arena.cc`` and ``work_queue.cc`` compiled UNMODIFIED (oracle/Makefile)
  behind ``oracle/ref_harness.cc``.  This is the ground truth for ``{hit, t, pos}``.
* :class:`Port` -- ``oracle/libbsporacle.so``: ``oracle/restate.c``, a plain-C
  restatement with visit counters and the extension outputs (plane / brush /
  startsolid).  Pinned bit-for-bit against :class:`Reference` and the KAT vectors
  in ``tests/test_oracle.py``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libbspref.so")
PORT_SO = os.path.join(HERE, "libbsporacle.so")

RESULT_DT = np.dtype([("t", "<f4"), ("plane", "<i4"), ("brush", "<i4"), ("flags", "<u4")])
COUNTER_FIELDS = ("traces", "nodes", "leaves", "leaf_brushes", "solid", "sides", "crossings", "hits", "max_depth")


def build(verbose: bool = False) -> None:
    """(re)build the checkers; the reference object only where /root/reference exists."""
    r = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
    if verbose or r.returncode:
        print(r.stdout, r.stderr)
    if r.returncode:
        raise RuntimeError("oracle build failed")
    build_patched_reference(verbose)


def build_patched_reference(verbose: bool = False) -> None:
    """oracle/_ref/patched_ref_kat: the reference's own bsp.cc / bsp.h with INTEGRATION.md's patch applied
    (tools/apply_integration.py) + its unmodified bsp_renderer.cc, linked against libbspgpu.so -- the checker of the drop-in
    boundary (tests/test_integration_patch.py).  Only where the reference sources and the built product library exist."""
    import sys
    import tempfile
    root = os.path.dirname(HERE)
    ref = os.environ.get("BSP_REF_DIR", "/root/reference")
    lib = os.path.join(root, "bsp_b200", "libbspgpu.so")
    exe = os.path.join(HERE, "_ref", "patched_ref_kat")
    if not os.path.exists(os.path.join(ref, "bsp.cc")) or not os.path.exists(lib):
        return
    deps = [lib, os.path.join(root, "tools", "apply_integration.py"), os.path.join(root, "tests", "cpp", "patched_ref_kat.cc"), os.path.join(root, "include", "bspgpu.h")]
    if os.path.exists(exe) and all(os.path.getmtime(d) <= os.path.getmtime(exe) for d in deps):
        return
    sys.path.insert(0, os.path.join(root, "tests"))
    try:
        import test_integration_patch
        with tempfile.TemporaryDirectory() as tmp:
            test_integration_patch.build_patched(tmp)
    except Exception as e:  # the checker of one test: never fatal for the build
        if verbose:
            print("patched reference not built:", e)
    finally:
        sys.path.pop(0)


def have_reference() -> bool:
    return os.path.exists(REF_SO)


def _seg_view(segs: np.ndarray):
    """accepts (n,6) packed vec3 pairs or (n,8) float4 pairs; returns (ptr, stride, end_off, n)."""
    segs = np.ascontiguousarray(segs, np.float32)
    if segs.ndim != 2 or segs.shape[1] not in (6, 8):
        raise ValueError("segments must be (n,6) or (n,8) float32")
    w = segs.shape[1]
    return segs, segs.ctypes.data_as(C.POINTER(C.c_float)), w, w // 2, segs.shape[0]


class Reference:
    """The unmodified reference ``BSP`` loaded from a .bsp file (bsp.cc:58)."""

    _lib = None

    def __init__(self, path: str):
        if Reference._lib is None:
            if not have_reference():
                raise RuntimeError(f"{REF_SO} missing: run `make -C oracle ref` where /root/reference exists")
            lib = C.CDLL(REF_SO)
            lib.bspref_load.restype = C.c_void_p
            lib.bspref_load.argtypes = [C.c_char_p]
            lib.bspref_free.argtypes = [C.c_void_p]
            lib.bspref_counts.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
            lib.bspref_sizes.argtypes = [C.POINTER(C.c_uint32)]
            fp, u8p = C.POINTER(C.c_float), C.POINTER(C.c_uint8)
            lib.bspref_trace.restype = C.c_double
            lib.bspref_trace.argtypes = [C.c_void_p, fp, C.c_uint64, C.c_uint64, C.c_uint64, u8p, fp, fp, u8p]
            lib.bspref_trace_mt.restype = C.c_double
            lib.bspref_trace_mt.argtypes = [C.c_void_p, fp, C.c_uint64, C.c_uint64, C.c_uint64, u8p, fp, fp, u8p, C.c_uint32]
            lib.bspref_leaf.argtypes = [C.c_void_p, fp, C.c_uint64, C.c_uint64, C.POINTER(C.c_int32)]
            Reference._lib = lib
        self.lib = Reference._lib
        self.h = self.lib.bspref_load(path.encode())
        self.seconds = 0.0

    def close(self):
        if self.h:
            self.lib.bspref_free(self.h)
            self.h = None

    def counts(self):
        out = (C.c_uint32 * 7)()
        self.lib.bspref_counts(self.h, out)
        return dict(zip(("textures", "planes", "nodes", "leaves", "leaf_brushes", "brushes", "brush_sides"), out))

    @classmethod
    def struct_sizes(cls):
        Reference("/dev/null") if False else None
        lib = cls._lib or C.CDLL(REF_SO)
        lib.bspref_sizes.argtypes = [C.POINTER(C.c_uint32)]
        out = (C.c_uint32 * 10)()
        lib.bspref_sizes(out)
        return dict(zip(("header", "texture", "plane", "node", "leaf", "brush", "brushside", "vertex", "face", "intersection"), out))

    def trace(self, segs: np.ndarray, threads: int = 1):
        """returns dict(hit u8, t f32, pos (n,3) f32, plane_set u8); ``self.seconds`` = loop time."""
        segs, ptr, stride, end_off, n = _seg_view(segs)
        hit = np.zeros(n, np.uint8); t = np.zeros(n, np.float32)
        pos = np.zeros((n, 3), np.float32); pset = np.zeros(n, np.uint8)
        fp, u8p = C.POINTER(C.c_float), C.POINTER(C.c_uint8)
        args = [self.h, ptr, stride, end_off, n, hit.ctypes.data_as(u8p), t.ctypes.data_as(fp),
                pos.ctypes.data_as(fp), pset.ctypes.data_as(u8p)]
        if threads <= 1:
            self.seconds = self.lib.bspref_trace(*args)
        else:
            self.seconds = self.lib.bspref_trace_mt(*args, threads)
            if self.seconds < 0:
                raise RuntimeError("reference work_queue already initialised with a different thread count")
        return {"hit": hit, "t": t, "pos": pos, "plane_set": pset}

    def time_only(self, segs: np.ndarray, threads: int = 1) -> float:
        """trace without keeping outputs (hit/t still written once); returns loop seconds."""
        self.trace(segs, threads)
        return self.seconds

    def leaf(self, pts: np.ndarray) -> np.ndarray:
        pts = np.ascontiguousarray(pts, np.float32)
        out = np.zeros(pts.shape[0], np.int32)
        self.lib.bspref_leaf(self.h, pts.ctypes.data_as(C.POINTER(C.c_float)), pts.shape[1], pts.shape[0],
                             out.ctypes.data_as(C.POINTER(C.c_int32)))
        return out


class _OrcMap(C.Structure):
    _fields_ = sum(([(nm, C.c_void_p), ("num_" + nm, C.c_uint32)] for nm in
                    ("textures", "planes", "nodes", "leaves", "leaf_brushes", "brushes", "brush_sides")), [])


class _OrcCounters(C.Structure):
    _fields_ = [(nm, C.c_uint64) for nm in COUNTER_FIELDS]


class Port:
    """``oracle/restate.c`` over the raw lumps of a :class:`tools.ibsp.IBSP`."""

    _lib = None

    def __init__(self, m):
        if Port._lib is None:
            if not os.path.exists(PORT_SO):
                build()
            lib = C.CDLL(PORT_SO)
            fp = C.POINTER(C.c_float)
            lib.orc_trace.argtypes = [C.POINTER(_OrcMap), fp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, fp, C.POINTER(_OrcCounters)]
            lib.orc_leaf.argtypes = [C.POINTER(_OrcMap), fp, C.c_uint64, C.c_uint64, C.POINTER(C.c_int32)]
            Port._lib = lib
        self.lib = Port._lib
        self._keep = {}
        om = _OrcMap()
        for nm in ("textures", "planes", "nodes", "leaves", "leaf_brushes", "brushes", "brush_sides"):
            arr = np.ascontiguousarray(getattr(m, nm))
            self._keep[nm] = arr
            setattr(om, nm, arr.ctypes.data if arr.size else None)
            setattr(om, "num_" + nm, len(arr))
        self.om = om

    def trace(self, segs: np.ndarray, want_pos: bool = True):
        """returns (results RESULT_DT[n], pos (n,3) or None, counters dict)."""
        segs, ptr, stride, end_off, n = _seg_view(segs)
        res = np.zeros(n, RESULT_DT)
        pos = np.zeros((n, 3), np.float32) if want_pos else None
        ctr = _OrcCounters()
        self.lib.orc_trace(C.byref(self.om), ptr, stride, end_off, n, res.ctypes.data,
                           pos.ctypes.data_as(C.POINTER(C.c_float)) if want_pos else None, C.byref(ctr))
        return res, pos, {nm: int(getattr(ctr, nm)) for nm in COUNTER_FIELDS}

    def leaf(self, pts: np.ndarray) -> np.ndarray:
        pts = np.ascontiguousarray(pts, np.float32)
        out = np.zeros(pts.shape[0], np.int32)
        self.lib.orc_leaf(C.byref(self.om), pts.ctypes.data_as(C.POINTER(C.c_float)), pts.shape[1], pts.shape[0],
                          out.ctypes.data_as(C.POINTER(C.c_int32)))
        return out


def bytes_per_trace(counters: dict, b_io: int = 48) -> dict:
    """B = B_io + 28 n_node + 8 n_leaf + 12 n_lb + 8 n_solid + 20 n_side (means per trace)."""
    n = max(1, counters["traces"])
    mean = {k: counters[k] / n for k in ("nodes", "leaves", "leaf_brushes", "solid", "sides")}
    map_bytes = 28 * mean["nodes"] + 8 * mean["leaves"] + 12 * mean["leaf_brushes"] + 8 * mean["solid"] + 20 * mean["sides"]
    return {"b_io": b_io, "b_map": map_bytes, "b_total": b_io + map_bytes, "means": mean,
            "hit_rate": counters["hits"] / n}
