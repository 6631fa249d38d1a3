"""`q3style`: a map with the content classes a q3map-compiled level carries and the kd-style synthetic maps lack
(the reference loads a Warsow map, bsp.cc:303, which is not redistributable):

  * bevel planes: every brush with a non-axial side also carries the axial planes of its bounding box (q3map adds
    them for the engine's box traces; to the reference's trace they are just more sides -- some of them redundant,
    some touching the brush only along an edge, which is where the >= / <= ties of bsp.cc:148,154 live);
  * mixed content flags: SOLID (1), SOLID|DETAIL (1 | 0x8000000), WATER (0x20), PLAYERCLIP (0x10000): only bit 0 is
    tested (bsp.cc:193), so water and clip must be skipped although they sit in the same leaves;
  * structural walls and floors referenced from many leaves (a beam through the whole hall: >= 8 leaves), detail
    clutter listed in the leaves it touches, and leaves with >= 32 brushes (a high traversal cost keeps the compiler from splitting small sets);
  * leaves with clusters (16 rooms; leaves buried in structural walls have cluster -1) and a real visibility lump:
    a room sees itself and every room within two doorways;
  * written as a stock 17-slot IBSP file (the reference's header has 18 slots, bsp.h:26,38; slot 17 is never read).
"""
from __future__ import annotations

import numpy as np

from tools import ibsp
from tools.mapgen import ROTATIONS, BrushList, Rng, compile_bsp

TEXTURES = [
    (b"textures/q3style/wall", 0, 1),                       # CONTENTS_SOLID, structural
    (b"textures/q3style/detail", 0, 1 | 0x8000000),         # CONTENTS_SOLID | CONTENTS_DETAIL
    (b"textures/q3style/water", 0, 0x20),                   # CONTENTS_WATER
    (b"textures/common/clip", 0x80, 0x10000),               # CONTENTS_PLAYERCLIP
    (b"textures/q3style/trim", 0, 1),
]
T_WALL, T_DETAIL, T_WATER, T_CLIP, T_TRIM = range(5)


def _with_bevels(bl: BrushList, index: int) -> None:
    """append the axial planes of the brush's bounding box that are not among its sides yet (q3map's AddBrushBevels,
    axial part).  The brush stays the same point set; only its side list grows."""
    first, n = bl.first_side[index], bl.num_sides[index]
    assert first + n == len(bl.side_planes), "bevels are added right after the brush"
    have = {bl.planes[bl.side_planes[first + k]][:3] for k in range(n)}
    (x0, y0, z0), (x1, y1, z1) = bl.bmin[index], bl.bmax[index]
    for nrm, d in (((1.0, 0.0, 0.0), x1), ((-1.0, 0.0, 0.0), -x0), ((0.0, 1.0, 0.0), y1), ((0.0, -1.0, 0.0), -y0),
                   ((0.0, 0.0, 1.0), z1), ((0.0, 0.0, -1.0), -z0)):
        if nrm not in have:
            bl.side_planes.append(bl.plane(nrm[0], nrm[1], nrm[2], d))
            bl.num_sides[index] += 1


def q3style(seed: int = 12) -> ibsp.IBSP:
    rng = Rng(seed)
    bl = BrushList()
    R, G, T = 1024, 4, 32                      # room pitch, rooms per side, wall thickness
    W = R * G // 2
    z0, z1 = 0, 768
    structural = []                            # AABBs of structural brushes: leaves buried in them get cluster -1

    def wall(x0, x1, y0, y1, za, zb, tex=T_WALL):
        bl.box(x0, x1, y0, y1, za, zb, tex)
        structural.append((x0, y0, za, x1, y1, zb))

    wall(-W - T, W + T, -W - T, W + T, z0 - T, z0)
    wall(-W - T, W + T, -W - T, W + T, z1, z1 + T)
    wall(-W - T, -W, -W, W, z0, z1); wall(W, W + T, -W, W, z0, z1)
    wall(-W, W, -W - T, -W, z0, z1); wall(-W, W, W, W + T, z0, z1)
    door_w, door_h = 256, 384
    doors = set()                              # (room a, room b) joined by a doorway
    for k in range(1, G):                      # interior walls, one doorway per room boundary (a few boundaries stay closed)
        c = -W + k * R
        for j in range(G):
            a0, a1 = -W + j * R, -W + (j + 1) * R
            for axis in (0, 1):
                closed = rng.chance(1, 5)
                dc = (a0 + a1) // 2 + rng.randint(-200, 200) // 16 * 16
                spans = [(a0, a1, z0, z1)] if closed else [(a0, dc - door_w // 2, z0, z1), (dc + door_w // 2, a1, z0, z1), (dc - door_w // 2, dc + door_w // 2, z0 + door_h, z1)]
                for (s0, s1, za, zb) in spans:
                    if axis == 0:
                        wall(c - T // 2, c + T // 2, s0, s1, za, zb)
                    else:
                        wall(s0, s1, c - T // 2, c + T // 2, za, zb)
                if not closed:
                    ra = (k - 1) + j * G if axis == 0 else j + (k - 1) * G
                    rb = k + j * G if axis == 0 else j + k * G
                    doors.add((ra, rb))
    # beams through the whole hall: each is listed in every leaf it crosses
    for i in range(6):
        y = -W + 300 + i * 640
        bl.box(-W, W, y, y + 24, z1 - 160, z1 - 128, T_TRIM)
        bl.box(y, y + 24, -W, W, z1 - 96, z1 - 64, T_TRIM)
    # room contents: detail clutter (bevelled non-axial brushes), piles, water pools, clip volumes
    for ry in range(G):
        for rx in range(G):
            x0, x1 = -W + rx * R + 64, -W + (rx + 1) * R - 64
            y0, y1 = -W + ry * R + 64, -W + (ry + 1) * R - 64
            for _ in range(28):
                kind = rng.randint(0, 3)
                cx, cy = rng.randint(x0 + 80, x1 - 80), rng.randint(y0 + 80, y1 - 80)
                if kind == 0:
                    i = bl.pillar(cx, cy, rng.randint(10, 48), z0, z0 + rng.randint(64, 700), ROTATIONS[rng.randint(0, len(ROTATIONS) - 1)], T_DETAIL)
                elif kind == 1:
                    sx, sy, sz = rng.randint(64, 200), rng.randint(48, 160), rng.randint(32, 220)
                    i = bl.ramp(cx - sx // 2, cx + sx // 2, cy - sy // 2, cy + sy // 2, z0, z0 + sz, along_x=rng.chance(1, 2), texture=T_DETAIL)
                elif kind == 2:
                    sx, sy, sz = rng.randint(48, 160), rng.randint(48, 160), rng.randint(24, 120)
                    zb = rng.randint(z0, z0 + 300)
                    i = bl.roof(cx - sx // 2, cx + sx // 2, cy - sy // 2, cy + sy // 2, zb, zb + sz, ridge_along_x=rng.chance(1, 2), texture=T_DETAIL)
                else:
                    sx, sy, sz = rng.randint(16, 96), rng.randint(16, 96), rng.randint(16, 256)
                    i = bl.box(cx - sx // 2, cx + sx // 2 + 1, cy - sy // 2, cy + sy // 2 + 1, z0, z0 + sz, T_DETAIL if rng.chance(1, 2) else T_WALL)
                if kind != 3:
                    # the conservative AABB the compiler uses may be half a unit larger than the brush; bevels must touch the brush
                    _with_bevels(bl, i)
            # a pile of small crates in one corner: a leaf with dozens of brushes
            px, py = rng.randint(x0, x1 - 200), rng.randint(y0, y1 - 200)
            for _ in range(44):
                sx, sy, sz = rng.randint(8, 40), rng.randint(8, 40), rng.randint(8, 60)
                x, y, zb = px + rng.randint(0, 160), py + rng.randint(0, 160), z0 + rng.randint(0, 120)
                bl.box(x, x + sx, y, y + sy, zb, zb + sz, T_DETAIL if rng.chance(2, 3) else T_WALL)
            sx, sy = rng.randint(160, 400), rng.randint(160, 400)
            x, y = rng.randint(x0, x1 - sx), rng.randint(y0, y1 - sy)
            bl.box(x, x + sx, y, y + sy, z0, z0 + rng.randint(24, 96), T_WATER)
            x, y = rng.randint(x0, x1 - 128), rng.randint(y0, y1 - 128)
            bl.box(x, x + 128, y, y + 128, z0, z1, T_CLIP)
    m = compile_bsp(bl, (-W - T, -W - T, z0 - T), (W + T, W + T, z1 + T), textures=TEXTURES, leaf_max=48, max_depth=40, cost_traverse=24.0)

    # ---- clusters: one per room; a leaf whose centre lies inside a structural brush has none (q3map: opaque leaves)
    lo = m.leaves["mins"].astype(np.float64); hi = m.leaves["maxs"].astype(np.float64)
    ctr = (lo + hi) / 2
    s = np.asarray(structural, np.float64)
    buried = ((ctr[:, None, :] >= s[None, :, :3]) & (ctr[:, None, :] <= s[None, :, 3:])).all(axis=2).any(axis=1)
    rx = np.clip(((ctr[:, 0] + W) // R).astype(np.int64), 0, G - 1)
    ry = np.clip(((ctr[:, 1] + W) // R).astype(np.int64), 0, G - 1)
    # (the compiler's leaves are coarse, few are buried: every 9th leaf is declared opaque as well so that the cluster -1
    # paths -- bsp_renderer.cc:150 and leaves no PVS row names -- are exercised)
    buried |= (np.arange(len(ctr)) % 9) == 4
    m.leaves["cluster"] = np.where(buried, -1, rx + ry * G).astype(np.int32)
    m.leaves["area"] = 0

    # ---- visibility: a room sees itself and the rooms within two doorways
    nclusters = G * G
    adj = np.eye(nclusters, dtype=bool)
    for a, b in doors:
        adj[a, b] = adj[b, a] = True
    vis2 = (adj.astype(np.int64) @ adj.astype(np.int64)) > 0
    csize = (nclusters + 7) // 8
    rows = np.zeros((nclusters, csize), np.uint8)
    for a in range(nclusters):
        for b in range(nclusters):
            if vis2[a, b]:
                rows[a, b >> 3] |= np.uint8(1 << (b & 7))
    m.vis = np.concatenate((np.array([nclusters, csize], "<u4").view(np.uint8), rows.reshape(-1)))

    refs = np.bincount(m.leaf_brushes, minlength=len(m.brushes))
    m.meta.update(name="q3style", num_lumps=17, max_leaf_brushes=int(m.leaves["num_brushes"].max()), max_leaves_per_brush=int(refs.max()),
                  seg_min=[-W, -W, z0], seg_max=[W, W, z1], clusters=nclusters, doors=len(doors))
    return m
