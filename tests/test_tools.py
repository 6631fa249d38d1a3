"""Workload tooling: IBSP round trip, map compiler invariants, segment generators."""
import hashlib

import numpy as np
import pytest

from helpers import bits


def test_ibsp_roundtrip_and_17_slot_header(maps):
    from tools import ibsp
    m, path = maps("kat1")
    blob = ibsp.to_bytes(m)
    assert blob[:4] == b"IBSP" and len(blob) % 4 == 0
    back = ibsp.from_bytes(blob)
    for nm in ("textures", "planes", "nodes", "leaves", "leaf_brushes", "brushes", "brush_sides"):
        assert np.array_equal(getattr(back, nm), getattr(m, nm)), nm
    # stock Quake 3 files have 17 directory slots (reference bsp.h:26 has 18; slot 17 is never read)
    stock = ibsp.from_bytes(ibsp.to_bytes(m, num_lumps=17))
    assert np.array_equal(stock.nodes, m.nodes) and np.array_equal(stock.brush_sides, m.brush_sides)
    # the visibility lump satisfies the reference loader's assert (bsp.cc:111): 8 bytes, {0,0}
    dirent = np.frombuffer(blob, "<u4", count=36, offset=8).reshape(18, 2)
    assert dirent[ibsp.LUMP_VISIBILITY, 1] == 8


def test_maps_are_deterministic():
    from tools import ibsp, mapgen
    a = ibsp.to_bytes(mapgen.rooms8()); b = ibsp.to_bytes(mapgen.rooms8())
    assert hashlib.sha256(a).hexdigest() == hashlib.sha256(b).hexdigest()
    assert ibsp.to_bytes(mapgen.rooms8(seed=2)) != a


def test_c_and_numpy_tree_builders_agree():
    """tools/bspc.c is the numpy split search transcribed; both must emit the same bytes."""
    import functools
    from tools import ibsp, mapgen
    fast = mapgen.rooms8()
    orig = mapgen.compile_bsp
    mapgen.compile_bsp = functools.partial(orig, builder="numpy")
    try:
        slow = mapgen.rooms8()
    finally:
        mapgen.compile_bsp = orig
    assert ibsp.to_bytes(fast) == ibsp.to_bytes(slow)


@pytest.mark.parametrize("name,lo_brushes,hi_brushes", [("rooms8", 400, 600), ("city10k", 9000, 11500)])
def test_compiled_tree_invariants(maps, name, lo_brushes, hi_brushes):
    m, _ = maps(name)
    assert lo_brushes <= len(m.brushes) <= hi_brushes
    n_nodes, n_leaves = len(m.nodes), len(m.leaves)
    seen_nodes = np.zeros(n_nodes, bool); seen_leaves = np.zeros(n_leaves, bool)
    stack = [0]
    while stack:
        i = stack.pop()
        assert not seen_nodes[i]
        seen_nodes[i] = True
        for c in m.nodes["children"][i]:
            if c >= 0:
                stack.append(int(c))
            else:
                assert not seen_leaves[-(c + 1)]
                seen_leaves[-(c + 1)] = True
    assert seen_nodes.all() and seen_leaves.all()          # a proper binary tree: every node / leaf reached once
    # all split planes are axial with unit normals; brush plane normals are unit length
    nrm = np.linalg.norm(m.planes["n"].astype(np.float64), axis=1)
    assert np.allclose(nrm, 1.0, atol=1e-6)
    assert (m.brushes["first_side"] + m.brushes["num_sides"]).max() <= len(m.brush_sides)
    assert m.leaf_brushes.max() < len(m.brushes)
    # non-axial planes exist (so the arithmetic is not all exact) and so do non-solid brushes
    axial = (np.abs(m.planes["n"]) == 1.0).any(axis=1)
    assert (~axial).sum() > 10
    if name == "rooms8":
        solid = m.textures["content_flags"][m.brushes["texture"]] & 1
        assert 0 < (solid == 0).sum() < len(m.brushes)


def test_every_leaf_lists_every_overlapping_brush(maps):
    """brute force: a random point strictly inside a brush's AABB core must find that brush in its leaf."""
    from tools import mapgen
    import oracle
    m, _ = maps("rooms8")
    port = oracle.Port(m)
    rs = np.random.Generator(np.random.PCG64(3))
    planes = m.planes
    for b in rs.choice(len(m.brushes), 60, replace=False):
        br = m.brushes[b]
        pl = m.brush_sides["plane"][br["first_side"]:br["first_side"] + br["num_sides"]]
        n = planes["n"][pl].astype(np.float64); d = planes["d"][pl].astype(np.float64)
        # rejection-sample a point inside all half spaces
        lo, hi = mapgen.seg_bounds(m)
        axial = [(i, j) for i in range(len(pl)) for j in range(3) if abs(n[i, j]) == 1.0]
        box_lo = np.array(lo, np.float64); box_hi = np.array(hi, np.float64)
        for i, j in axial:
            if n[i, j] > 0:
                box_hi[j] = min(box_hi[j], d[i])
            else:
                box_lo[j] = max(box_lo[j], -d[i])
        for _ in range(200):
            p = box_lo + (box_hi - box_lo) * rs.random(3)
            if np.all(n @ p - d < -1e-3):
                break
        else:
            continue
        leaf = int(port.leaf(np.asarray([p], np.float32))[0])
        lf = m.leaves[leaf]
        assert b in m.leaf_brushes[lf["first_brush"]:lf["first_brush"] + lf["num_brushes"]], (b, leaf)


def test_seggen_is_counter_based():
    from tools import seggen
    lo, hi = np.array([-10, -20, -30], np.float32), np.array([10, 20, 30], np.float32)
    whole = seggen.uniform(7, 100, 1000, lo, hi)
    parts = np.vstack((seggen.uniform(7, 100, 400, lo, hi), seggen.uniform(7, 500, 600, lo, hi)))
    assert np.array_equal(bits(whole), bits(parts))       # any rank can generate exactly its shard
    assert not np.array_equal(whole, seggen.uniform(8, 100, 1000, lo, hi))
    assert (whole[:, 0:3] >= lo).all() and (whole[:, 0:3] <= hi).all()
    assert (whole[:, 3] == 0).all() and (whole[:, 7] == 0).all()
    big = seggen.uniform(7, (1 << 40) + 5, 10, lo, hi)     # 64-bit indices
    assert np.isfinite(big).all()


def test_long_rays_lengths_and_isotropy():
    from tools import seggen
    lo, hi = np.array([-100, -100, -100], np.float32), np.array([100, 100, 100], np.float32)
    s = seggen.long_rays(3, 0, 200_000, lo, hi, 50.0, 200.0)
    d = (s[:, 4:7] - s[:, 0:3]).astype(np.float64)
    ln = np.linalg.norm(d, axis=1)
    assert ln.min() > 49.9 and ln.max() < 200.1
    u = d / ln[:, None]
    assert np.abs(u.mean(axis=0)).max() < 0.01            # no preferred direction
    assert np.abs((u ** 2).mean(axis=0) - 1 / 3).max() < 0.01


def test_camera_grid():
    from tools import seggen
    lo, hi = np.array([-100, -100, 0], np.float32), np.array([100, 100, 50], np.float32)
    views = seggen.make_views(4, 2, lo, hi)
    segs = seggen.camera(views, 64, 36, 0, 2 * 64 * 36)
    assert np.array_equal(segs[:64 * 36, 0:3], np.broadcast_to(views[0, 0:3], (64 * 36, 3)))
    centre = segs[18 * 64 + 32]
    ray = (centre[4:7] - centre[0:3]) / 32768.0
    assert np.dot(ray, views[0, 3:6]) > 0.99


def test_shard_ranges_cover_everything(native):
    for n in (0, 1, 7, 1000, 1 << 33):
        for world in (1, 2, 3, 8):
            got = [native.shard_range(n, r, world) for r in range(world)]
            assert got[0][0] == 0 and sum(c for _, c in got) == n
            for (f0, c0), (f1, _) in zip(got, got[1:]):
                assert f0 + c0 == f1
