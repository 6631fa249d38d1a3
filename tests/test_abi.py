"""The C-ABI library loads on a GPU-less box, exports every symbol include/bspgpu.h declares,
validates maps before touching CUDA, and fails loudly (no CPU fallback) without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported(native):
    lib = native.load_library()
    text = open(os.path.join(ROOT, "include", "bspgpu.h")).read()
    declared = sorted(set(re.findall(r"\b(bspgpu_[a-z_]+)\s*\(", text)))
    assert len(declared) >= 17
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in include/bspgpu.h but not exported by libbspgpu.so"
    from bsp_b200 import api
    assert sorted(api.ABI_SYMBOLS) == declared


def test_abi_version_and_shard_helper(native):
    lib = native.load_library()
    assert lib.bspgpu_abi_version() == 1
    assert native.shard_range(10, 1, 3) == (4, 4)
    assert native.shard_range(10, 2, 3) == (8, 2)


def test_wire_formats(native):
    assert native.RESULT_DT.itemsize == 16
    from bsp_b200 import api
    assert C.sizeof(api.MapDesc) == 7 * 16
    assert C.sizeof(api.Info) >= 60


def _has_gpu():
    import torch
    return torch.cuda.is_available()


def test_no_cpu_fallback(native, maps):
    if _has_gpu():
        pytest.skip("a device is present")
    m, path = maps("kat0")
    with pytest.raises(native.BspGpuError, match="no CUDA device"):
        native.BspGpu(path)
    with pytest.raises(native.BspGpuError, match="no CUDA device"):
        native.BspGpu(m)


def test_map_validation_happens_before_cuda(native, maps):
    import copy
    m, _ = maps("kat1")
    bad = copy.deepcopy(m)
    bad.nodes["children"][1] = (0, -2)            # cycle back to the root
    with pytest.raises(native.BspGpuError, match="reached twice"):
        native.BspGpu(bad)
    bad = copy.deepcopy(m)
    bad.brush_sides["plane"][0] = 9999
    with pytest.raises(native.BspGpuError, match="out of range"):
        native.BspGpu(bad)
    with pytest.raises(native.BspGpuError, match="cannot open"):
        native.BspGpu("/nonexistent/map.bsp")


def test_null_arguments(native):
    lib = native.load_library()
    assert lib.bspgpu_create(None, None, 0) == -1
    assert b"null" in lib.bspgpu_last_error()
    assert lib.bspgpu_trace(None, None, 1, None, 0) < 0
    assert lib.bspgpu_hit_count(None, None) < 0
    lib.bspgpu_destroy(None)                      # harmless
