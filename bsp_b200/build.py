"""Builds the in-tree native libraries (explicit nvcc / g++, no JIT cache).

libbspgpu.so   -- CUDA kernels + C ABI (include/bspgpu.h), sm_100a only.
libbsphost.so  -- the C++ host mirror of the reference's BSP interface (csrc/host).

Numerics flags are part of the contract (results must be bit-identical to the reference's
x86-64 SSE2 arithmetic): no FMA contraction, IEEE division and square root, denormals kept.
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_GPU = os.path.join(HERE, "libbspgpu.so")
LIB_HOST = os.path.join(HERE, "libbsphost.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false", "-ftz=false", "-prec-div=true", "-prec-sqrt=true",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall", "-shared", "-cudart", "static",
]
CXX_FLAGS = ["-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-Wall", "-Wextra", "-shared"]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found")
    return exe


def _stale(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources(*dirs_) -> list[str]:
    out = []
    for d in dirs_:
        for root, _, files in os.walk(d):
            out += [os.path.join(root, f) for f in files if f.endswith((".cu", ".cuh", ".cc", ".h"))]
    return out


def build(force: bool = False, verbose: bool = False, ptxas_info: bool = False, experiments: bool = False) -> None:
    """experiments=True adds -DBSPGPU_EXPERIMENTS: the measured-slower round-1 alternatives (one thread per segment, top of the
    tree in shared memory) become selectable again; the product build leaves them out."""
    deps = _sources(CSRC, os.path.join(os.path.dirname(HERE), "include"))
    if force or _stale(LIB_GPU, deps):
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if ptxas_info else []) + (["-DBSPGPU_EXPERIMENTS"] if experiments else []) + [
            "-o", LIB_GPU, os.path.join(CSRC, "bspgpu.cu"), os.path.join(CSRC, "bspgpu_multi.cu"), os.path.join(CSRC, "flatten.cc"), "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or ptxas_info or r.returncode:
            print(" ".join(cmd)); print(r.stdout); print(r.stderr)
        if r.returncode:
            raise RuntimeError("nvcc failed building libbspgpu.so")
    host_src = os.path.join(CSRC, "host", "bsp_host.cc")
    if os.path.exists(host_src) and (force or _stale(LIB_HOST, deps + [LIB_GPU])):
        cmd = ["g++"] + CXX_FLAGS + ["-o", LIB_HOST, host_src, "-L" + HERE, "-lbspgpu", "-Wl,-rpath,$ORIGIN"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode:
            print(" ".join(cmd)); print(r.stdout); print(r.stderr)
        if r.returncode:
            raise RuntimeError("g++ failed building libbsphost.so")


def build_microbench(force: bool = False, verbose: bool = False) -> str:
    """tools/microbench/microbench: measures the L2 / shared-memory / scattered-gather peaks on the box."""
    root = os.path.dirname(HERE)
    src = os.path.join(root, "tools", "microbench", "microbench.cu")
    exe = os.path.join(root, "tools", "microbench", "microbench")
    if force or _stale(exe, [src]):
        cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-o", exe, src]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode:
            print(" ".join(cmd)); print(r.stdout); print(r.stderr)
        if r.returncode:
            raise RuntimeError("nvcc failed building tools/microbench")
    return exe


if __name__ == "__main__":
    import sys
    build(force="--force" in sys.argv or "--experiments" in sys.argv, verbose=True, ptxas_info="--ptxas" in sys.argv, experiments="--experiments" in sys.argv)
