"""The oracle is pinned: the unmodified reference object and the C restatement reproduce the
KAT-0 / KAT-1 vectors the survey captured from the reference, the committed
golden fixtures (reference outputs), and each other bit for bit."""
import numpy as np
import pytest

from helpers import bits, load_golden, mixed_segments


def _kat_segments(vecs):
    segs = np.zeros((len(vecs), 8), np.float32)
    for i, v in enumerate(vecs):
        segs[i, 0:3] = v[0]; segs[i, 4:7] = v[1]
    return segs


@pytest.mark.parametrize("name", ["kat0", "kat1"])
def test_kat_vectors_port_and_reference(oracle_mod, maps, name):
    from tools import mapgen
    vecs = mapgen.KAT0_VECTORS if name == "kat0" else mapgen.KAT1_VECTORS
    m, path = maps(name)
    segs = _kat_segments(vecs)
    res, pos, _ = oracle_mod.Port(m).trace(segs)
    for i, v in enumerate(vecs):
        assert (res["flags"][i] & 1) == v[2], (name, i)
        assert int(bits(res["t"][i:i + 1])[0]) == v[3], (name, i)
        if v[4] is not None:
            np.testing.assert_allclose(pos[i], np.asarray(v[4], np.float32), rtol=1e-6, atol=1e-6)
    if oracle_mod.have_reference():
        r = oracle_mod.Reference(path).trace(segs)
        for i, v in enumerate(vecs):
            assert r["hit"][i] == v[2] and int(bits(r["t"][i:i + 1])[0]) == v[3], (name, i)
        assert int(r["plane_set"].sum()) == 0     # Intersection::plane is never assigned (bsp.cc:256,268)


def test_kat_quirks_are_kept(oracle_mod, maps):
    """reference behaviours a 'fixed' tracer would get differently."""
    m, _ = maps("kat0")
    port = oracle_mod.Port(m)
    segs = _kat_segments([((1, 0, 0), (64, 0, 0)), ((16, 0, 0), (-64, 0, 0)), ((64, 0, 0), (16, 0, 0))])
    res, _, _ = port.trace(segs)
    assert (res["flags"][0] & 1) == 0 and (res["flags"][0] & 2)      # starts inside: no hit, startsolid extension set
    assert (res["flags"][1] & 1) == 0                                 # starts on a face: no hit
    assert (res["flags"][2] & 1) == 1 and res["t"][2] == 1.0          # ends on a face: hit with t == 1 exactly
    m1, _ = maps("kat1")
    res, _, _ = oracle_mod.Port(m1).trace(_kat_segments([((64, 0, 0), (-64, 0, 0))]))
    assert (res["flags"][0] & 1) and res["t"][0] == 0.5 and res["plane"][0] == -1   # hit reported at tmin, no side raised tfar


@pytest.mark.parametrize("name", ["kat0", "kat1", "rooms8", "tilted"])
def test_port_matches_golden_reference_outputs(oracle_mod, name):
    m, segs, hit, t, pos = load_golden(name)
    res, ppos, ctr = oracle_mod.Port(m).trace(segs)
    assert np.array_equal((res["flags"] & 1).astype(np.uint8), hit)
    assert np.array_equal(bits(res["t"]), bits(t))
    assert np.array_equal(bits(ppos), bits(pos))
    assert ctr["hits"] == int(hit.sum())


@pytest.mark.parametrize("name", ["extreme_rooms8", "extreme_odd"])
def test_port_matches_golden_reference_outputs_for_extreme_floats(oracle_mod, name):
    """infinite / NaN / huge / denormal coordinates, and a solid brush without sides: the port answers what the unmodified
    reference answered (fixtures made by tests/golden/make_golden.py) -- this pins the oracle the GPU is held to there."""
    from helpers import check_against_extreme_golden
    m, segs, hit, t, pos = load_golden(name)
    with np.errstate(all="ignore"):
        res, ppos, _ = oracle_mod.Port(m).trace(segs)
    check_against_extreme_golden(res["flags"] & 1, res["t"], ppos, hit, t, pos, name)


@pytest.mark.parametrize("name", ["kat1", "rooms8", "city10k", "tilted"])
def test_port_matches_reference_object(oracle_mod, maps, name):
    if not oracle_mod.have_reference():
        pytest.skip("reference object not built on this box")
    m, path = maps(name)
    segs = mixed_segments(m, 150_000, 100_000, 50_000, 8192)
    res, pos, _ = oracle_mod.Port(m).trace(segs)
    r = oracle_mod.Reference(path).trace(segs)
    assert np.array_equal(r["hit"], (res["flags"] & 1).astype(np.uint8))
    assert np.array_equal(bits(r["t"]), bits(res["t"]))
    assert np.array_equal(bits(r["pos"]), bits(pos))


def test_reference_work_queue_matches_single_thread(oracle_mod, maps):
    if not oracle_mod.have_reference():
        pytest.skip("reference object not built on this box")
    m, path = maps("rooms8")
    segs = mixed_segments(m, 200_000, 50_000, 50_000, 4096)
    ref = oracle_mod.Reference(path)
    a = ref.trace(segs, threads=1)
    b = ref.trace(segs, threads=4)
    assert np.array_equal(a["hit"], b["hit"]) and np.array_equal(bits(a["t"]), bits(b["t"]))


def test_position_to_leaf_port_vs_reference(oracle_mod, maps):
    from tools import mapgen, seggen
    m, path = maps("rooms8")
    lo, hi = mapgen.seg_bounds(m)
    pts = np.ascontiguousarray(seggen.uniform(3, 0, 100_000, lo, hi)[:, :4])
    got = oracle_mod.Port(m).leaf(pts)
    assert got.min() >= 0 and got.max() < len(m.leaves)
    if oracle_mod.have_reference():
        assert np.array_equal(got, oracle_mod.Reference(path).leaf(pts))


def test_reference_struct_sizes(oracle_mod):
    """the generator's dtypes are the reference's on-disk structs (bsp.h:29-139)."""
    if not oracle_mod.have_reference():
        pytest.skip("reference object not built on this box")
    from tools import ibsp
    sz = oracle_mod.Reference.struct_sizes()
    assert sz["header"] == ibsp.HEADER_BYTES == 152
    assert (sz["texture"], sz["plane"], sz["node"], sz["leaf"], sz["brush"], sz["brushside"]) == (
        ibsp.TEXTURE_DT.itemsize, ibsp.PLANE_DT.itemsize, ibsp.NODE_DT.itemsize, ibsp.LEAF_DT.itemsize,
        ibsp.BRUSH_DT.itemsize, ibsp.BRUSHSIDE_DT.itemsize)
    assert sz["intersection"] == 24


def test_bytes_per_trace_formula(oracle_mod):
    ctr = dict(traces=10, nodes=100, leaves=20, leaf_brushes=10, solid=5, sides=30, crossings=0, hits=0, max_depth=0)
    b = oracle_mod.bytes_per_trace(ctr, b_io=48)
    assert b["b_map"] == pytest.approx(28 * 10 + 8 * 2 + 12 * 1 + 8 * 0.5 + 20 * 3)
    assert b["b_total"] == pytest.approx(48 + b["b_map"])
