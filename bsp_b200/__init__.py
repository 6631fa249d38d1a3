"""bsp_b200 -- B200-native batched BSP segment tracing (drop-in for mikejsavage/bsp's
``BSP::trace_seg`` path, reference bsp.cc:120-271).

The product is native: ``libbspgpu.so`` (hand-written sm_100a CUDA kernels behind the C ABI
of ``include/bspgpu.h``) and ``libbsphost.so`` (the C++ host mirror of the reference's ``BSP``
interface, ``csrc/host``).  This Python package is only the thin ctypes binding the tests and
``bench.py`` drive that C ABI with; it contains no tracing code and no CPU fallback -- if the
CUDA library is missing or there is no sm_100 device, calls raise.
"""
from .api import (  # noqa: F401
    BspGpu, BspGpuError, RESULT_DT, RESULT8_DT, SEGMENT_FLOATS, HIT, STARTSOLID, T_MISS, lib_path, load_library, shard_range,
    VARIANT_AUTO, VARIANT_SMEM, VARIANT_TOPTREE, VARIANT_GLOBAL,
)
from .build import build  # noqa: F401
