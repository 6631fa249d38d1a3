"""The C-ABI library loads on a GPU-less box, exports every symbol include/bspgpu.h declares,
validates maps before touching CUDA, and fails loudly (no CPU fallback) without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)           # prototypes only, not the prose around them
    return sorted(set(re.findall(r"\b(bspgpu_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported(native):
    lib = native.load_library()
    declared = _declared("bspgpu.h")
    assert len(declared) >= 23
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in include/bspgpu.h but not exported by libbspgpu.so"
    from bsp_b200 import api
    assert sorted(api.ABI_SYMBOLS) == declared


def test_multi_gpu_header_symbols_are_exported(native):
    """include/bspgpu_multi.h: the in-process multi-GPU entry points and the rank/world (one process per GPU) ones;
    NCCL itself is bound at run time, so the library loads on a box without it"""
    lib = native.load_library()
    declared = _declared("bspgpu_multi.h")
    assert len(declared) == 16, declared
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in include/bspgpu_multi.h but not exported by libbspgpu.so"
    out = subprocess_ldd(native.lib_path())
    assert "libnccl" not in out and "libcudart" not in out      # no link-time dependency on NCCL; static CUDA runtime
    from bsp_b200 import multi
    assert sorted(multi.MULTI_SYMBOLS) == declared


def subprocess_ldd(path):
    import subprocess
    return subprocess.run(["ldd", path], capture_output=True, text=True).stdout


def test_abi_version_and_shard_helper(native):
    lib = native.load_library()
    assert lib.bspgpu_abi_version() == 2
    assert native.shard_range(10, 1, 3) == (4, 4)
    assert native.shard_range(10, 2, 3) == (8, 2)


def test_wire_formats(native):
    assert native.RESULT_DT.itemsize == 16
    from bsp_b200 import api
    assert C.sizeof(api.MapDesc) == 7 * 16
    assert C.sizeof(api.Info) >= 80
    assert native.RESULT8_DT.itemsize == 8 and C.sizeof(api.CreateOptions) == 16


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
    assert lib.bspgpu_multi_create(None, None, None, 0) == -1
    assert lib.bspgpu_dist_init(None, None, 0, 1) == -1
    lib.bspgpu_multi_destroy(None)


def test_multi_create_fails_loudly_without_a_gpu(native, maps):
    if _has_gpu():
        pytest.skip("a device is present")
    from bsp_b200 import multi
    _, path = maps("kat0")
    with pytest.raises(native.BspGpuError, match="no CUDA device"):
        multi.BspGpuMulti(path)


def test_flag_and_option_values_match_the_header(native):
    """the Python bindings spell the trace flags and the miss marker as numbers: they must be the header's"""
    from bsp_b200 import api
    text = open(os.path.join(ROOT, "include", "bspgpu.h")).read()
    defs = {m.group(1): m.group(2) for m in re.finditer(r"#define\s+(BSPGPU_[A-Z0-9_]+)\s+(0x[0-9a-fA-F]+u|[0-9.]+f)", text)}
    for name, val in (("BSPGPU_SEGS_DEVICE", api.SEGS_DEVICE), ("BSPGPU_OUT_DEVICE", api.OUT_DEVICE), ("BSPGPU_SEGS_PACKED24", api.SEGS_PACKED24),
                      ("BSPGPU_ASYNC", api.ASYNC), ("BSPGPU_OUT_COMPACT8", api.OUT_COMPACT8), ("BSPGPU_OUT_COMPACT4", api.OUT_COMPACT4)):
        assert int(defs[name].rstrip("u"), 16) == val, name
    assert float(defs["BSPGPU_T_MISS"].rstrip("f")) == native.T_MISS == 2.0
    enums = dict(re.findall(r"\b(BSPGPU_OPT_[A-Z_]+)\s*=\s*(\d+)", text))
    assert enums["BSPGPU_OPT_TINY_CALLS"] == "18" and enums["BSPGPU_OPT_PAGEABLE"] == "17" and enums["BSPGPU_OPT_SMEM_STACK"] == "15"
    assert len(set(enums.values())) == len(enums)                # no two options share a number
