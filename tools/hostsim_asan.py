"""Development aid: the CPU build of the device trace core (tests/hostsim: trace_core.h + flatten.cc) compiled with
-fsanitize=address,undefined and driven over every map / walk / option -- compute-sanitizer is not available on the GPU pool,
and an out-of-range slot, brush reference or stack index would show up here just as well.  Prints one line per map."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = "/tmp/libhostsim_asan.so"

CHILD = r'''
import sys, ctypes as C, numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
import helpers
lib = C.CDLL(%(so)r)
lib.hostsim_create_opts.restype = C.c_void_p
lib.hostsim_create_opts.argtypes = [C.c_void_p, C.c_int, C.c_int64]
lib.hostsim_error.restype = C.c_char_p
lib.hostsim_destroy.argtypes = [C.c_void_p]
lib.hostsim_layout.argtypes = [C.c_void_p, C.c_void_p]
lib.hostsim_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
lib.hostsim_trace_phased.argtypes = lib.hostsim_trace.argtypes + [C.c_uint32, C.c_uint32, C.c_int]
lib.hostsim_trace_slots.argtypes = lib.hostsim_trace_phased.argtypes
lib.hostsim_leaf.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]
helpers.HostSim._lib = lib
from tools import mapgen
for name in ("kat0", "kat1", "rooms8", "tilted", "q3style", "city10k"):
    m, _ = mapgen.get_map(name)
    sets = [helpers.mixed_segments(m, 20000, 20000, 20000, 2048), helpers.extreme_segments(5000, 2048)]
    for gen in (False, True):
        sim = helpers.HostSim(m, general_planes_only=gen)
        has_grid = sim.layout()["grid_cells"] > 0
        if has_grid and not gen:
            sets.append(helpers.grid_adversarial_segments(m, 20000))
        for sg in sets:
            for kw in ({}, {"grid": has_grid}, {"split_stack": True}, {"walk": "nodes", "grid": has_grid}, {"split_sides": True}):
                with np.errstate(all="ignore"):
                    sim.trace(sg, (0, 0), **kw)
            sim.trace(sg)
    print(name, "clean", flush=True)
'''


def main():
    subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer",
                    "-ffp-contract=off", "-fPIC", "-shared", "-o", SO, os.path.join(ROOT, "tests", "hostsim", "hostsim.cc"),
                    os.path.join(ROOT, "bsp_b200", "csrc", "flatten.cc")], check=True)
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0")
    r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT, "so": SO}], env=env)
    sys.exit(r.returncode)


if __name__ == "__main__":
    main()
