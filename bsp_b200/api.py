"""ctypes binding of include/bspgpu.h (one function per C entry point, no logic of its own)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

RESULT_DT = np.dtype([("t", "<f4"), ("plane", "<i4"), ("brush", "<i4"), ("flags", "<u4")])
SEGMENT_FLOATS = 8
HIT, STARTSOLID = 1, 2

RESULT8_DT = np.dtype([("t", "<f4"), ("bits", "<u4")])      # bspgpu_result8: bits = flags | (plane + 1) << 2
SEGS_DEVICE, OUT_DEVICE, SEGS_PACKED24, ASYNC, OUT_COMPACT8, OUT_COMPACT4 = 1, 2, 4, 8, 16, 32
T_MISS = 2.0    # BSPGPU_T_MISS: what BSPGPU_OUT_COMPACT4 stores for a miss
VARIANT_AUTO, VARIANT_SMEM, VARIANT_TOPTREE, VARIANT_GLOBAL = 0, 1, 2, 3
OPT_VARIANT, OPT_BLOCK_THREADS, OPT_CTAS_PER_SM, OPT_TOPTREE_BYTES, OPT_CHUNK_SEGMENTS, OPT_REFILL, \
    OPT_MAX_STEPS, OPT_BLOCK_SEGS, OPT_SCHEDULER, OPT_SIDE_STEPS, OPT_LEAF_THRESHOLD, OPT_MAX_LAUNCH_SEGMENTS, OPT_START_GRID = range(1, 14)
OPT_SMEM_STACK, OPT_BIN_SEGMENTS, OPT_PAGEABLE = 15, 16, 17

# every symbol include/bspgpu.h declares (tests/test_abi.py checks the library exports them all)
ABI_SYMBOLS = (
    "bspgpu_create", "bspgpu_create_opts", "bspgpu_create_from_file", "bspgpu_destroy", "bspgpu_trace", "bspgpu_trace_pos", "bspgpu_trace_stream",
    "bspgpu_hit_count", "bspgpu_hit_count_device", "bspgpu_leaf", "bspgpu_leaf_clusters", "bspgpu_set_vis", "bspgpu_visible_leaves",
    "bspgpu_generate", "bspgpu_set_option", "bspgpu_set_stream",
    "bspgpu_reset_stream", "bspgpu_get_info", "bspgpu_sync", "bspgpu_host_alloc", "bspgpu_host_free", "bspgpu_shard_range",
    "bspgpu_last_error", "bspgpu_abi_version",
)


class BspGpuError(RuntimeError):
    pass


class MapDesc(C.Structure):
    _fields_ = sum(([(nm, C.c_void_p), ("num_" + nm, C.c_uint32)] for nm in
                    ("textures", "planes", "nodes", "leaves", "leaf_brushes", "brushes", "brush_sides")), [])


class GenDesc(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("seed", C.c_uint64), ("lo", C.c_float * 3), ("hi", C.c_float * 3),
                ("len_lo", C.c_float), ("len_hi", C.c_float), ("views", C.c_void_p),
                ("num_views", C.c_uint32), ("width", C.c_uint32), ("height", C.c_uint32),
                ("tan_half", C.c_float), ("far_dist", C.c_float)]


class Info(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("device", C.c_int32), ("sm_count", C.c_uint32), ("variant", C.c_uint32),
                ("block_threads", C.c_uint32), ("grid_ctas", C.c_uint32), ("smem_bytes", C.c_uint32),
                ("tree_depth", C.c_uint32), ("num_nodes", C.c_uint32), ("num_brush_refs", C.c_uint32),
                ("num_sides", C.c_uint32), ("map_bytes", C.c_uint64), ("last_kernel_ms", C.c_float),
                ("last_launches", C.c_uint32), ("start_grid_cells", C.c_uint32), ("staged_bytes", C.c_uint32),
                ("num_general_planes", C.c_uint32), ("num_leaves", C.c_uint32), ("num_clusters", C.c_uint32),
                ("smem_stack_slots", C.c_uint32), ("last_prepass_ms", C.c_float)]


class CreateOptions(C.Structure):
    _fields_ = [("start_grid_cells", C.c_int64), ("general_planes_only", C.c_int32), ("reserved", C.c_int32)]


_lib = None


def lib_path() -> str:
    return os.path.join(HERE, "libbspgpu.so")


def load_library() -> C.CDLL:
    """dlopen libbspgpu.so and declare the prototypes.  Raises if the library was not built --
    there is nothing to fall back to."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise BspGpuError(f"{path} not built: run `python -m bsp_b200.build` (or __graft_entry__.build())")
    lib = C.CDLL(path)
    vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
    lib.bspgpu_create.argtypes = [C.POINTER(vp), C.POINTER(MapDesc), i32]
    lib.bspgpu_create_opts.argtypes = [C.POINTER(vp), C.POINTER(MapDesc), i32, C.POINTER(CreateOptions)]
    lib.bspgpu_create_from_file.argtypes = [C.POINTER(vp), C.c_char_p, i32]
    lib.bspgpu_destroy.argtypes = [vp]; lib.bspgpu_destroy.restype = None
    lib.bspgpu_trace.argtypes = [vp, vp, u64, vp, u32]
    lib.bspgpu_trace_pos.argtypes = [vp, vp, u64, vp, vp, u32]
    lib.bspgpu_trace_stream.argtypes = [vp, vp, u64, vp, vp, u32, vp]
    lib.bspgpu_hit_count.argtypes = [vp, C.POINTER(u64)]
    lib.bspgpu_hit_count_device.argtypes = [vp, vp]
    lib.bspgpu_leaf.argtypes = [vp, vp, u32, u64, vp, u32]
    lib.bspgpu_leaf_clusters.argtypes = [vp, vp, u32, u64, vp, vp, C.c_int32, vp, u32]
    lib.bspgpu_set_vis.argtypes = [vp, vp, u64]
    lib.bspgpu_visible_leaves.argtypes = [vp, vp, vp, C.POINTER(C.c_int32)]
    lib.bspgpu_generate.argtypes = [vp, C.POINTER(GenDesc), u64, u64, vp]
    lib.bspgpu_set_option.argtypes = [vp, i32, C.c_int64]
    lib.bspgpu_set_stream.argtypes = [vp, vp]
    lib.bspgpu_reset_stream.argtypes = [vp]
    lib.bspgpu_get_info.argtypes = [vp, C.POINTER(Info)]
    lib.bspgpu_sync.argtypes = [vp]
    lib.bspgpu_host_alloc.argtypes = [C.c_size_t]; lib.bspgpu_host_alloc.restype = vp
    lib.bspgpu_host_free.argtypes = [vp]; lib.bspgpu_host_free.restype = None
    lib.bspgpu_shard_range.argtypes = [u64, u32, u32, C.POINTER(u64), C.POINTER(u64)]; lib.bspgpu_shard_range.restype = None
    lib.bspgpu_last_error.restype = C.c_char_p
    lib.bspgpu_abi_version.restype = u32
    _lib = lib
    return lib


def shard_range(n: int, rank: int, world: int):
    lib = load_library()
    first, count = C.c_uint64(), C.c_uint64()
    lib.bspgpu_shard_range(n, rank, world, C.byref(first), C.byref(count))
    return first.value, count.value


def _ptr(x):
    """device pointer of a torch tensor / host pointer of a numpy array / raw int."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray):
        if not x.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        if not x.is_contiguous():
            raise ValueError("tensor must be contiguous")
        return x.data_ptr()
    raise TypeError(type(x))


def _is_device(x) -> bool:
    return hasattr(x, "is_cuda") and bool(x.is_cuda)


class BspGpu:
    """One ``bspgpu_t`` handle: a map flattened into HBM on one device."""

    def __init__(self, source, device: int = -1, start_grid_cells: int = 0, general_planes_only: bool = False):
        """``source``: path to an IBSP file, or an object with the seven lump arrays (``tools.ibsp.IBSP``; its ``vis``
        lump, when present, is uploaded too).  ``start_grid_cells`` / ``general_planes_only``: bspgpu_create_options
        (lump objects only)."""
        self.lib = load_library()
        self.h = C.c_void_p()
        if isinstance(source, (str, bytes, os.PathLike)):
            if start_grid_cells or general_planes_only:
                raise ValueError("create options need a lump object, not a file path")
            rc = self.lib.bspgpu_create_from_file(C.byref(self.h), os.fsencode(source), device)
        else:
            desc = MapDesc()
            keep = []
            for nm in ("textures", "planes", "nodes", "leaves", "leaf_brushes", "brushes", "brush_sides"):
                arr = np.ascontiguousarray(getattr(source, nm))
                keep.append(arr)
                setattr(desc, nm, arr.ctypes.data if arr.size else None)
                setattr(desc, "num_" + nm, len(arr))
            opts = CreateOptions(int(start_grid_cells), 1 if general_planes_only else 0, 0)
            rc = self.lib.bspgpu_create_opts(C.byref(self.h), C.byref(desc), device, C.byref(opts))
            if rc == 0:
                vis = getattr(source, "vis", None)
                if vis is not None and len(vis) >= 8:
                    vis = np.ascontiguousarray(vis, np.uint8)
                    self._check(self.lib.bspgpu_set_vis(self.h, vis.ctypes.data, len(vis)))
        self._check(rc)

    def _check(self, rc: int):
        if rc != 0:
            raise BspGpuError(f"bspgpu error {rc}: {self.lib.bspgpu_last_error().decode()}")

    def close(self):
        if self.h:
            self.lib.bspgpu_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- options / plumbing
    def set_option(self, opt: int, value: int):
        self._check(self.lib.bspgpu_set_option(self.h, opt, int(value)))

    def set_stream(self, cuda_stream: int):
        self._check(self.lib.bspgpu_set_stream(self.h, cuda_stream))

    def reset_stream(self):
        self._check(self.lib.bspgpu_reset_stream(self.h))

    def sync(self):
        self._check(self.lib.bspgpu_sync(self.h))

    def info(self) -> dict:
        inf = Info()
        self._check(self.lib.bspgpu_get_info(self.h, C.byref(inf)))
        return {k: getattr(inf, k) for k, _ in Info._fields_}

    def hit_count(self) -> int:
        v = C.c_uint64()
        self._check(self.lib.bspgpu_hit_count(self.h, C.byref(v)))
        return v.value

    def hit_count_device(self, dev_i64):
        """copy the hit counter into a 1-element int64 cuda tensor (async on the handle's stream)."""
        self._check(self.lib.bspgpu_hit_count_device(self.h, _ptr(dev_i64)))

    # -- the hot path
    def trace(self, segs, out=None, pos=None, n=None, packed24=False, async_=False, compact=False, stream=None):
        """segs: (n,8) float32 numpy array / torch tensor (host or cuda), or (n,6) with packed24.
        out: RESULT_DT array or (n,4) int32/float32 tensor (compact=True or 8: RESULT8_DT / (n,2); compact=4: float32 (n,) -- t on a
        hit, T_MISS = 2.0 on a miss, BSPGPU_OUT_COMPACT4); allocated (numpy) when segs is a
        numpy array and neither out nor pos is given.  Returns out (or pos when only pos was asked for).
        CUDA tensors: the call is enqueued on `stream` (a raw cudaStream_t) -- by default torch's CURRENT stream, so it is
        ordered after the kernels that produced `segs`; pass stream=False to use whatever bspgpu_set_stream installed."""
        floats = 6 if packed24 else SEGMENT_FLOATS
        if not hasattr(segs, "shape") or len(segs.shape) != 2 or int(segs.shape[1]) != floats:
            raise ValueError(f"segs must be (n, {floats}) float32" + (" (packed24)" if packed24 else ""))
        if str(segs.dtype).replace("torch.", "") != "float32":
            raise ValueError(f"segs must be float32, not {segs.dtype}")
        if n is None:
            n = int(segs.shape[0])
        if n > int(segs.shape[0]):
            raise ValueError("n exceeds the number of segments passed")
        if out is None and pos is None:
            if not isinstance(segs, np.ndarray):
                raise ValueError("pass `out` for tensor inputs")
            out = np.zeros(n, np.float32) if compact == 4 else np.zeros(n, RESULT8_DT if compact else RESULT_DT)
        for name, x, need in (("out", out, n * (4 if compact == 4 else 8 if compact else 16)), ("pos", pos, n * 16)):
            if x is not None:
                have = x.nbytes if isinstance(x, np.ndarray) else x.numel() * x.element_size()
                if have < need:
                    raise ValueError(f"{name} holds {have} bytes, {need} needed for {n} segments")
        flags = (SEGS_DEVICE if _is_device(segs) else 0) | (SEGS_PACKED24 if packed24 else 0) | (ASYNC if async_ else 0) | (OUT_COMPACT4 if compact == 4 else OUT_COMPACT8 if compact else 0)
        outs_dev = [_is_device(x) for x in (out, pos) if x is not None]
        if any(outs_dev) and not all(outs_dev):
            raise ValueError("out and pos must both be host or both be device")
        if outs_dev and outs_dev[0]:
            flags |= OUT_DEVICE
        if stream is None and (flags & (SEGS_DEVICE | OUT_DEVICE)):
            import torch
            stream = torch.cuda.current_stream(segs.device if _is_device(segs) else out.device if out is not None else pos.device).cuda_stream
        if stream is None or stream is False:
            self._check(self.lib.bspgpu_trace_pos(self.h, _ptr(segs), n, _ptr(out), _ptr(pos), flags))
        else:
            self._check(self.lib.bspgpu_trace_stream(self.h, _ptr(segs), n, _ptr(out), _ptr(pos), flags, stream))
        return out if out is not None else pos

    def leaf(self, pts, out=None):
        if len(pts.shape) != 2 or int(pts.shape[1]) < 3 or str(pts.dtype).replace("torch.", "") != "float32":
            raise ValueError("pts must be (n, >=3) float32")
        n = int(pts.shape[0])
        stride = int(pts.shape[1])
        if out is None:
            out = np.zeros(n, np.int32)
        flags = (SEGS_DEVICE if _is_device(pts) else 0) | (OUT_DEVICE if _is_device(out) else 0)
        self._check(self.lib.bspgpu_leaf(self.h, _ptr(pts), stride, n, _ptr(out), flags))
        return out

    def leaf_clusters(self, pts, from_cluster: int = -1, want_visible: bool = False):
        """host points -> (leaf, cluster[, visible from `from_cluster`]) per point (bspgpu_leaf_clusters)"""
        pts = np.ascontiguousarray(pts, np.float32)
        n = len(pts)
        leaf, cluster = np.zeros(n, np.int32), np.zeros(n, np.int32)
        vis = np.zeros(n, np.uint8) if want_visible else None
        self._check(self.lib.bspgpu_leaf_clusters(self.h, pts.ctypes.data, pts.shape[1], n, leaf.ctypes.data, cluster.ctypes.data,
                                                  from_cluster, vis.ctypes.data if want_visible else None, 0))
        return (leaf, cluster, vis) if want_visible else (leaf, cluster)

    def visible_leaves(self, pos):
        """the loop of bspr_render (bsp_renderer.cc:147-168): (visible flag per leaf, viewer cluster)"""
        p = np.ascontiguousarray(pos, np.float32)
        vis = np.zeros(self.info()["num_leaves"], np.uint8)
        c = C.c_int32()
        self._check(self.lib.bspgpu_visible_leaves(self.h, p.ctypes.data, vis.ctypes.data, C.byref(c)))
        return vis, c.value

    def generate(self, dev_segs, first: int, count: int, kind: int, seed: int, lo, hi, len_lo=0.0, len_hi=0.0,
                 views=None, width=0, height=0, tan_half=1.0, far=32768.0):
        g = GenDesc()
        g.kind = kind; g.seed = seed
        for a in range(3):
            g.lo[a] = float(lo[a]); g.hi[a] = float(hi[a])
        g.len_lo = float(np.float32(len_lo)); g.len_hi = float(np.float32(len_hi))
        keep = None
        if views is not None:
            keep = np.ascontiguousarray(views, np.float32)
            g.views = keep.ctypes.data; g.num_views = keep.shape[0]
        g.width = width; g.height = height; g.tan_half = float(np.float32(tan_half)); g.far_dist = float(np.float32(far))
        self._check(self.lib.bspgpu_generate(self.h, C.byref(g), first, count, _ptr(dev_segs)))
