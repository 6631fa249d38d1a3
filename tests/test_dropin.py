"""The C++ host mirror of the reference interface (bsp_b200/csrc/host/bsp_host.h): bsp_init /
BSP::trace_seg / BSP::trace_segs / BSP::position_to_leaf, exercised by a C++ program that
reads like the reference's own call site (bsp.cc:363-370)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(native, tmp_path):
    native.build()
    exe = str(tmp_path / "test_dropin")
    libdir = os.path.join(ROOT, "bsp_b200")
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-Wall", "-Wextra", "-o", exe,
                    os.path.join(ROOT, "tests", "cpp", "test_dropin.cc"), "-L" + libdir, "-lbsphost", "-lbspgpu",
                    "-Wl,-rpath," + libdir], check=True)
    return exe


def test_dropin_builds_and_fails_loudly_without_a_gpu(native, maps, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    exe = _build(native, tmp_path)
    _, path = maps("kat1")
    r = subprocess.run([exe, path], capture_output=True, text=True)
    assert r.returncode != 0
    assert "no CUDA device" in r.stderr      # aborts like the reference's assert; never falls back to a CPU trace


@pytest.mark.gpu
def test_dropin_on_gpu(native, maps, tmp_path):
    exe = _build(native, tmp_path)
    _, path = maps("kat1")
    r = subprocess.run([exe, path], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout + r.stderr
