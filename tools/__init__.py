"""Workload tooling for the batched BSP segment tracer (not the product).

- ``ibsp``   : Quake-3 IBSP (v46) reader / writer, byte-compatible with the
               reference's on-disk structs (reference bsp.h:29-139).
- ``mapgen`` : brush primitives + a small axial-split BSP compiler and the
               committed synthetic maps (kat0, kat1, rooms8, city10k, mega200k).
- ``seggen`` : counter-based (splitmix64) segment generators whose floats are
               bit-identical to the CUDA generator kernel in bsp_b200/csrc.
"""
