"""ctypes binding of include/bspgpu_multi.h (one function per C entry point, no logic of its own).

`BspGpuMulti`  one process, several GPUs (bspgpu_multi_*): the library shards, owns the NCCL communicators.
`dist_*`       one process per GPU (bspgpu_dist_*): the launcher supplies rank / world and ships the NCCL unique id.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .api import BspGpuError, GenDesc, MapDesc, RESULT_DT, RESULT8_DT, OUT_COMPACT8, OUT_COMPACT4, SEGS_PACKED24, load_library

MULTI_SYMBOLS = (
    "bspgpu_multi_create", "bspgpu_multi_create_from_file", "bspgpu_multi_destroy", "bspgpu_multi_device_count", "bspgpu_multi_handle",
    "bspgpu_multi_trace", "bspgpu_multi_hit_count", "bspgpu_multi_generate", "bspgpu_multi_trace_resident", "bspgpu_multi_gather",
    "bspgpu_multi_read_results",
    "bspgpu_dist_unique_id", "bspgpu_dist_init", "bspgpu_dist_hit_count", "bspgpu_dist_gather", "bspgpu_dist_shutdown",
)
NCCL_ID_BYTES = 128

_declared = False


def _lib():
    global _declared
    lib = load_library()
    if not _declared:
        vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
        lib.bspgpu_multi_create.argtypes = [C.POINTER(vp), C.POINTER(MapDesc), C.POINTER(i32), i32]
        lib.bspgpu_multi_create_from_file.argtypes = [C.POINTER(vp), C.c_char_p, C.POINTER(i32), i32]
        lib.bspgpu_multi_destroy.argtypes = [vp]; lib.bspgpu_multi_destroy.restype = None
        lib.bspgpu_multi_device_count.argtypes = [vp]
        lib.bspgpu_multi_handle.argtypes = [vp, i32]; lib.bspgpu_multi_handle.restype = vp
        lib.bspgpu_multi_trace.argtypes = [vp, vp, u64, vp, vp, u32]
        lib.bspgpu_multi_hit_count.argtypes = [vp, C.POINTER(u64)]
        lib.bspgpu_multi_generate.argtypes = [vp, C.POINTER(GenDesc), u64, u64]
        lib.bspgpu_multi_trace_resident.argtypes = [vp]
        lib.bspgpu_multi_gather.argtypes = [vp, i32, C.POINTER(vp)]
        lib.bspgpu_multi_read_results.argtypes = [vp, u64, u64, vp]
        lib.bspgpu_dist_unique_id.argtypes = [vp]
        lib.bspgpu_dist_init.argtypes = [vp, vp, i32, i32]
        lib.bspgpu_dist_hit_count.argtypes = [vp, C.POINTER(u64), vp]
        lib.bspgpu_dist_gather.argtypes = [vp, vp, u64, vp, i32]
        lib.bspgpu_dist_shutdown.argtypes = [vp]
        _declared = True
    return lib


def _check(lib, rc):
    if rc != 0:
        raise BspGpuError(f"bspgpu error {rc}: {lib.bspgpu_last_error().decode()}")


class BspGpuMulti:
    """One ``bspgpu_multi_t``: the map replicated on several GPUs of this box, segments sharded by contiguous range."""

    def __init__(self, source, devices=None):
        self.lib = _lib()
        self.h = C.c_void_p()
        devs = (C.c_int * len(devices))(*devices) if devices else None
        nd = len(devices) if devices else 0
        if isinstance(source, (str, bytes, os.PathLike)):
            rc = self.lib.bspgpu_multi_create_from_file(C.byref(self.h), os.fsencode(source), devs, nd)
        else:
            desc = MapDesc()
            keep = []
            for nm in ("textures", "planes", "nodes", "leaves", "leaf_brushes", "brushes", "brush_sides"):
                arr = np.ascontiguousarray(getattr(source, nm))
                keep.append(arr)
                setattr(desc, nm, arr.ctypes.data if arr.size else None)
                setattr(desc, "num_" + nm, len(arr))
            rc = self.lib.bspgpu_multi_create(C.byref(self.h), C.byref(desc), devs, nd)
        _check(self.lib, rc)

    def close(self):
        if self.h:
            self.lib.bspgpu_multi_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def world(self) -> int:
        return self.lib.bspgpu_multi_device_count(self.h)

    def trace(self, segs: np.ndarray, out=None, pos=None, packed24=False, compact=False):
        """host arrays in, host arrays out (results in place, shard by shard)"""
        floats = 6 if packed24 else 8
        if not isinstance(segs, np.ndarray) or segs.ndim != 2 or segs.shape[1] != floats or segs.dtype != np.float32 or not segs.flags["C_CONTIGUOUS"]:
            raise ValueError(f"segs must be a C-contiguous (n, {floats}) float32 numpy array")
        n = len(segs)
        if out is None and pos is None:
            out = np.zeros(n, np.float32) if compact == 4 else np.zeros(n, RESULT8_DT if compact else RESULT_DT)
        flags = (SEGS_PACKED24 if packed24 else 0) | (OUT_COMPACT4 if compact == 4 else OUT_COMPACT8 if compact else 0)
        _check(self.lib, self.lib.bspgpu_multi_trace(self.h, segs.ctypes.data, n, out.ctypes.data if out is not None else None,
                                                     pos.ctypes.data if pos is not None else None, flags))
        return out if out is not None else pos

    def hit_count(self) -> int:
        v = C.c_uint64()
        _check(self.lib, self.lib.bspgpu_multi_hit_count(self.h, C.byref(v)))
        return v.value

    def generate(self, first, n_total, kind, seed, lo, hi, len_lo=0.0, len_hi=0.0):
        g = GenDesc()
        g.kind = kind; g.seed = seed
        for a in range(3):
            g.lo[a] = float(lo[a]); g.hi[a] = float(hi[a])
        g.len_lo = float(np.float32(len_lo)); g.len_hi = float(np.float32(len_hi))
        _check(self.lib, self.lib.bspgpu_multi_generate(self.h, C.byref(g), first, n_total))

    def trace_resident(self):
        _check(self.lib, self.lib.bspgpu_multi_trace_resident(self.h))

    def gather(self, dst_rank=0) -> int:
        p = C.c_void_p()
        _check(self.lib, self.lib.bspgpu_multi_gather(self.h, dst_rank, C.byref(p)))
        return p.value

    def read_results(self, first, count) -> np.ndarray:
        out = np.zeros(count, RESULT_DT)
        _check(self.lib, self.lib.bspgpu_multi_read_results(self.h, first, count, out.ctypes.data))
        return out


# ---- one process per GPU: the launcher ships the id (e.g. a torch.distributed broadcast of a uint8 tensor)

def dist_unique_id() -> bytes:
    lib = _lib()
    buf = (C.c_ubyte * NCCL_ID_BYTES)()
    _check(lib, lib.bspgpu_dist_unique_id(buf))
    return bytes(buf)


def dist_init(g, id_bytes: bytes, rank: int, world: int):
    lib = _lib()
    buf = (C.c_ubyte * NCCL_ID_BYTES).from_buffer_copy(id_bytes)
    _check(lib, lib.bspgpu_dist_init(g.h, buf, rank, world))


def dist_hit_count(g, dev_i64=None, want_host=True):
    """all-reduced hit count of the last trace (collective).  dev_i64: optional 1-element int64 cuda tensor that receives it."""
    lib = _lib()
    v = C.c_uint64()
    _check(lib, lib.bspgpu_dist_hit_count(g.h, C.byref(v) if want_host else None, dev_i64.data_ptr() if dev_i64 is not None else None))
    return v.value if want_host else None


def dist_gather(g, src, dst_tensor, dst: int = 0):
    """gather equal-sized cuda tensors to rank `dst` (collective, asynchronous on the handle's stream)"""
    lib = _lib()
    _check(lib, lib.bspgpu_dist_gather(g.h, src.data_ptr(), src.numel() * src.element_size(),
                                       dst_tensor.data_ptr() if dst_tensor is not None else None, dst))


def dist_shutdown(g):
    lib = _lib()
    _check(lib, lib.bspgpu_dist_shutdown(g.h))
