"""Quake-3 IBSP v46 reader/writer (numpy structured dtypes).

Layouts are the reference's on-disk structs (reference bsp.h:29-139; sizes
measured with the reference headers: header 152, texture 72, plane 16, node 36,
leaf 48, brush 12, brushside 8).  The reference header has 18 lump slots
(bsp.h:26,38 -- one more than stock Q3, the extra LIGHTARRAY slot is never
loaded), so files are written with 18 slots; stock 17-slot files read fine
because slot 17 is never dereferenced.

Loader constraints a file must satisfy or the *reference* traps (bsp.cc:97,111):
every typed lump's length is a multiple of its struct size and the visibility
lump is exactly ``8 + num_clusters*cluster_size`` bytes (we write ``{0,0}``).
"""
from __future__ import annotations

import dataclasses
import numpy as np

LUMP_ENTITIES, LUMP_TEXTURES, LUMP_PLANES, LUMP_NODES, LUMP_LEAVES, LUMP_LEAFFACES, \
    LUMP_LEAFBRUSHES, LUMP_MODELS, LUMP_BRUSHES, LUMP_BRUSHSIDES, LUMP_VERTICES, \
    LUMP_MESHVERTS, LUMP_FOGS, LUMP_FACES, LUMP_LIGHTING, LUMP_LIGHTGRID, \
    LUMP_VISIBILITY, LUMP_LIGHTARRAY = range(18)
NUM_LUMPS = 18
HEADER_BYTES = 8 + 8 * NUM_LUMPS  # 152

CONTENTS_SOLID = 1  # the "magic number" tested at reference bsp.cc:193

TEXTURE_DT = np.dtype([("name", "S64"), ("surface_flags", "<i4"), ("content_flags", "<i4")])
PLANE_DT = np.dtype([("n", "<f4", (3,)), ("d", "<f4")])
NODE_DT = np.dtype([("plane", "<u4"), ("children", "<i4", (2,)), ("mins", "<i4", (3,)), ("maxs", "<i4", (3,))])
LEAF_DT = np.dtype([("cluster", "<i4"), ("area", "<i4"), ("mins", "<i4", (3,)), ("maxs", "<i4", (3,)),
                    ("first_face", "<u4"), ("num_faces", "<u4"), ("first_brush", "<u4"), ("num_brushes", "<u4")])
BRUSH_DT = np.dtype([("first_side", "<u4"), ("num_sides", "<u4"), ("texture", "<i4")])
BRUSHSIDE_DT = np.dtype([("plane", "<u4"), ("texture", "<i4")])
LEAFBRUSH_DT = np.dtype("<u4")

assert TEXTURE_DT.itemsize == 72 and PLANE_DT.itemsize == 16 and NODE_DT.itemsize == 36
assert LEAF_DT.itemsize == 48 and BRUSH_DT.itemsize == 12 and BRUSHSIDE_DT.itemsize == 8


@dataclasses.dataclass
class IBSP:
    """The seven lumps the trace path reads (reference bsp.cc:120-271)."""
    textures: np.ndarray
    planes: np.ndarray
    nodes: np.ndarray
    leaves: np.ndarray
    leaf_brushes: np.ndarray
    brushes: np.ndarray
    brush_sides: np.ndarray
    meta: dict = dataclasses.field(default_factory=dict)
    vis: np.ndarray = None   # the raw visibility lump (uint8): {u32 num_clusters, u32 cluster_size, pvs rows}, reference bsp.h:129-134

    def counts(self) -> dict:
        return {k: int(len(getattr(self, k))) for k in
                ("textures", "planes", "nodes", "leaves", "leaf_brushes", "brushes", "brush_sides")}

    def map_bytes(self) -> int:
        return sum(getattr(self, k).nbytes for k in
                   ("textures", "planes", "nodes", "leaves", "leaf_brushes", "brushes", "brush_sides"))


def to_bytes(m: IBSP, num_lumps: int = NUM_LUMPS) -> bytes:
    lumps = {
        LUMP_TEXTURES: np.ascontiguousarray(m.textures, TEXTURE_DT).tobytes(),
        LUMP_PLANES: np.ascontiguousarray(m.planes, PLANE_DT).tobytes(),
        LUMP_NODES: np.ascontiguousarray(m.nodes, NODE_DT).tobytes(),
        LUMP_LEAVES: np.ascontiguousarray(m.leaves, LEAF_DT).tobytes(),
        LUMP_LEAFBRUSHES: np.ascontiguousarray(m.leaf_brushes, LEAFBRUSH_DT).tobytes(),
        LUMP_BRUSHES: np.ascontiguousarray(m.brushes, BRUSH_DT).tobytes(),
        LUMP_BRUSHSIDES: np.ascontiguousarray(m.brush_sides, BRUSHSIDE_DT).tobytes(),
        # {num_clusters=0, cluster_size=0} unless the map carries clusters
        LUMP_VISIBILITY: np.zeros(2, "<u4").tobytes() if m.vis is None else np.ascontiguousarray(m.vis, np.uint8).tobytes(),
    }
    header_bytes = 8 + 8 * num_lumps
    dirent = np.zeros((num_lumps, 2), "<u4")
    body = bytearray()
    off = header_bytes
    for lump in range(num_lumps):
        data = lumps.get(lump, b"")
        dirent[lump] = (off, len(data))
        body += data
        pad = (-len(data)) % 4
        body += b"\0" * pad
        off += len(data) + pad
    return b"IBSP" + np.uint32(46).tobytes() + dirent.tobytes() + bytes(body)


def write(m: IBSP, path: str, num_lumps: int = NUM_LUMPS) -> None:
    with open(path, "wb") as f:
        f.write(to_bytes(m, num_lumps))


def from_bytes(blob: bytes) -> IBSP:
    if len(blob) < 8 + 8 * 17:
        raise ValueError("IBSP blob shorter than a 17-lump header")
    # slot 17 is only read if present; the trace path needs lumps <= 9
    dirent = np.frombuffer(blob, "<u4", count=2 * 17, offset=8).reshape(17, 2)

    def lump(idx, dt):
        off, ln = int(dirent[idx, 0]), int(dirent[idx, 1])
        if ln % dt.itemsize:
            raise ValueError(f"lump {idx}: length {ln} not a multiple of {dt.itemsize}")
        if off + ln > len(blob):
            raise ValueError(f"lump {idx}: runs past end of file")
        return np.frombuffer(blob, dt, count=ln // dt.itemsize, offset=off).copy()

    return IBSP(
        textures=lump(LUMP_TEXTURES, TEXTURE_DT), planes=lump(LUMP_PLANES, PLANE_DT),
        nodes=lump(LUMP_NODES, NODE_DT), leaves=lump(LUMP_LEAVES, LEAF_DT),
        leaf_brushes=lump(LUMP_LEAFBRUSHES, LEAFBRUSH_DT), brushes=lump(LUMP_BRUSHES, BRUSH_DT),
        brush_sides=lump(LUMP_BRUSHSIDES, BRUSHSIDE_DT),
        meta={"magic": blob[:4], "version": int(np.frombuffer(blob, "<u4", 1, 4)[0])},
        vis=lump(LUMP_VISIBILITY, np.dtype("u1")),
    )


def read(path: str) -> IBSP:
    with open(path, "rb") as f:
        return from_bytes(f.read())
