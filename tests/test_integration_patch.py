"""INTEGRATION.md's patch applied MECHANICALLY to the reference's own sources (tools/apply_integration.py), compiled
with the reference's own class BSP -- every member of bsp.h:160-193 present, bsp_init / load_lump / load_vis /
position_to_leaf unmodified -- and the reference's bsp_renderer.cc (unmodified) linked next to it: the drop-in
boundary proven against the real reference rather than against the stand-alone mirror in bsp_b200/csrc/host.

The patched tree is built here (where /root/reference exists) into oracle/_ref/patched_ref_kat -- git-ignored, but
it travels to the GPU box with the snapshot like oracle/_ref/libbspref.so -- and run there on KAT-0 / KAT-1."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("BSP_REF_DIR", "/root/reference")
EXE = os.path.join(ROOT, "oracle", "_ref", "patched_ref_kat")


def build_patched(tmp):
    sys.path.insert(0, ROOT)
    from tools import apply_integration
    apply_integration.main(REF, tmp)
    flags = ["-std=c++11", "-O2", "-pthread", "-ffp-contract=off", "-w", "-I" + tmp, "-I" + os.path.join(ROOT, "oracle", "shim"), "-I" + os.path.join(ROOT, "include")]
    objs = []
    for src in (os.path.join(tmp, "bsp.cc"), os.path.join(tmp, "bsp_renderer.cc"), os.path.join(tmp, "memory_arena.cc"),
                os.path.join(ROOT, "tests", "cpp", "patched_ref_kat.cc")):
        obj = os.path.join(tmp, os.path.basename(src) + ".o")
        subprocess.run(["g++"] + flags + ["-c", src, "-o", obj], check=True)
        objs.append(obj)
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    libdir = os.path.join(ROOT, "bsp_b200")
    subprocess.run(["g++", "-o", EXE] + objs + ["-L" + libdir, "-lbspgpu", "-pthread", "-Wl,-rpath,$ORIGIN/../../bsp_b200"], check=True)


def _maps(maps):
    return [maps("kat0")[1], maps("kat1")[1]]


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "bsp.cc")), reason="the reference sources are not on this box")
def test_patch_applies_and_the_patched_reference_builds(native, maps, tmp_path):
    native.build()
    build_patched(str(tmp_path))
    patched_h = open(tmp_path / "bsp.h").read()
    # nothing of the reference's class was dropped, the two additions are there
    for member in ("leaf_faces", "vertices", "mesh_verts", "faces", "BSP_Vis * vis", "bspgpu_t * gpu", "trace_segs"):
        assert member in patched_h, member
    assert os.path.islink(tmp_path / "bsp_renderer.cc")           # the renderer is the reference's own file, unmodified
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([EXE] + _maps(maps), capture_output=True, text=True)
        assert r.returncode != 0 and "no CUDA device" in r.stderr   # the reference's assert traps; no CPU fallback behind it


@pytest.mark.gpu
def test_patched_reference_runs_the_kat_vectors_on_the_gpu(native, maps):
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/patched_ref_kat was not built (needs /root/reference at build time)")
    native.build()
    r = subprocess.run([EXE] + _maps(maps), capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout + r.stderr
    assert "vis 0 clusters" in r.stdout
