"""GPU parity: the sm_100a kernels, driven through the C ABI (include/bspgpu.h), against the
oracle on the same seeded inputs.  Bar: bit-exact on every field ({hit,t,pos} are the
reference's outputs; plane/brush/startsolid are the extension fields defined by
oracle/restate.c).  Reference semantics cited: bsp.cc:120-271."""
import numpy as np
import pytest

from conftest import bits

pytestmark = pytest.mark.gpu

VARIANTS = {"smem": 1, "global": 3}


def _segments(m, n_uniform=200_000, n_long=200_000, n_short=100_000, n_edge=16384, seed=21):
    from tools import mapgen, seggen
    lo, hi = mapgen.seg_bounds(m)
    diag = float(np.sqrt(((hi - lo).astype(np.float64) ** 2).sum()))
    return np.vstack((
        seggen.edge_cases(m, n_edge, seed),
        seggen.uniform(seed, 0, n_uniform, lo, hi),
        seggen.long_rays(seed + 1, 0, n_long, lo, hi, 0.25 * diag, diag),
        seggen.long_rays(seed + 2, 0, n_short, lo, hi, 1.0, 32.0),
    ))


def _compare(res, pos, exp, exp_pos, what):
    got = bits(res.view(np.uint32).reshape(-1, 4)) if res.dtype != np.uint32 else res
    want = bits(exp.view(np.uint32).reshape(-1, 4))
    bad = (got != want).any(axis=1)
    if bad.any():
        i = int(np.nonzero(bad)[0][0])
        # north_star: any mismatch is logged with index and both hex values
        pytest.fail(f"{what}: {int(bad.sum())} mismatching results; first at {i}: got {res[i]} ({got[i]}) want {exp[i]} ({want[i]})")
    if pos is not None:
        badp = (bits(pos[:, :3]) != bits(exp_pos)).any(axis=1) | (bits(pos[:, 3]) != bits(exp["t"]))
        assert not badp.any(), f"{what}: {int(badp.sum())} mismatching pos; first at {int(np.nonzero(badp)[0][0])}"


def test_kat_vectors(native, maps):
    """SURVEY.md 8c KAT-0 / KAT-1, captured from the unmodified reference."""
    from tools import mapgen
    for name, vecs in (("kat0", mapgen.KAT0_VECTORS), ("kat1", mapgen.KAT1_VECTORS)):
        m, path = maps(name)
        g = native.BspGpu(path)
        segs = np.zeros((len(vecs), 8), np.float32)
        for i, v in enumerate(vecs):
            segs[i, 0:3] = v[0]; segs[i, 4:7] = v[1]
        pos = np.zeros((len(vecs), 4), np.float32)
        res = g.trace(segs, out=np.zeros(len(vecs), native.RESULT_DT), pos=pos)
        for i, v in enumerate(vecs):
            assert (res["flags"][i] & 1) == v[2], (name, i)
            assert int(bits(res["t"][i:i + 1])[0]) == v[3], (name, i, hex(int(bits(res["t"][i:i + 1])[0])))
            # Intersection::pos = start + t*(end-start) in fp32 (bsp.cc:265-267); the survey lists it rounded
            s32, e32 = segs[i, 0:3], segs[i, 4:7]
            want = (s32 + res["t"][i] * (e32 - s32)) if v[2] else np.zeros(3, np.float32)
            assert np.array_equal(bits(pos[i, :3]), bits(want.astype(np.float32))), (name, i)
            if v[4] is not None:
                np.testing.assert_allclose(pos[i, :3], np.asarray(v[4], np.float32), rtol=1e-6, atol=1e-6, err_msg=f"{name} {i}")
        g.close()


@pytest.mark.parametrize("variant", list(VARIANTS))
@pytest.mark.parametrize("mapname", ["kat1", "rooms8", "city10k", "tilted"])
def test_parity_vs_oracle(native, maps, oracle_mod, mapname, variant):
    m, path = maps(mapname)
    g = native.BspGpu(m)
    if variant == "smem" and g.info()["staged_bytes"] > 200 * 1024:
        with pytest.raises(native.BspGpuError):
            g.set_option(1, VARIANTS[variant]); g.trace(np.zeros((1, 8), np.float32))
        return
    g.set_option(1, VARIANTS[variant])
    segs = _segments(m)
    exp, exp_pos, ctr = oracle_mod.Port(m).trace(segs)
    pos = np.zeros((len(segs), 4), np.float32)
    res = g.trace(segs, out=np.zeros(len(segs), native.RESULT_DT), pos=pos)
    _compare(res, pos, exp, exp_pos, f"{mapname}/{variant}")
    assert g.hit_count() == int((exp["flags"] & 1).sum()) == ctr["hits"]
    if oracle_mod.have_reference():
        ref = oracle_mod.Reference(path).trace(segs)
        assert np.array_equal(ref["hit"], (res["flags"] & 1).astype(np.uint8))
        assert np.array_equal(bits(ref["t"]), bits(res["t"]))
        assert np.array_equal(bits(ref["pos"]), bits(pos[:, :3]))
    g.close()


@pytest.mark.parametrize("refill,max_steps,threads", [(1, 1, 128), (32, 1000, 256), (8, 4, 1024), (16, 64, 512)])
@pytest.mark.parametrize("side_steps,leaf_thr", [(1, 1), (8, 16), (1000, 32)])
def test_scheduling_knobs_do_not_change_results(native, maps, oracle_mod, refill, max_steps, threads, side_steps, leaf_thr):
    m, _ = maps("rooms8")
    segs = _segments(m, 100_000, 50_000, 50_000, 4096)
    exp, exp_pos, _ = oracle_mod.Port(m).trace(segs)
    g = native.BspGpu(m)
    g.set_option(6, refill); g.set_option(7, max_steps); g.set_option(2, threads); g.set_option(10, side_steps); g.set_option(11, leaf_thr)
    res = g.trace(segs)
    _compare(res, None, exp, exp_pos, f"knobs {refill}/{max_steps}/{threads}")
    g.close()


@pytest.mark.parametrize("n", [0, 1, 31, 32, 33, 255, 257, 100_003])
def test_ragged_sizes_and_packed_layout(native, maps, oracle_mod, n):
    m, _ = maps("rooms8")
    segs = _segments(m, 60_000, 30_000, 10_000, 4096)[:n]
    g = native.BspGpu(m)
    res = g.trace(segs, out=np.zeros(n, native.RESULT_DT))
    packed = np.ascontiguousarray(np.hstack((segs[:, 0:3], segs[:, 4:7])))
    res24 = g.trace(packed, out=np.zeros(n, native.RESULT_DT), packed24=True)
    if n:
        exp, exp_pos, _ = oracle_mod.Port(m).trace(segs)
        _compare(res, None, exp, exp_pos, f"n={n}")
        _compare(res24, None, exp, exp_pos, f"n={n} packed24")
    assert g.hit_count() == int((res24["flags"] & 1).sum())
    g.close()


def test_host_pipeline_chunks_match_device_path(native, maps, oracle_mod):
    """host pointers go through the 3-stream chunk pipeline; device tensors go straight to the kernel."""
    import torch
    m, _ = maps("rooms8")
    segs = _segments(m, 300_000, 100_000, 50_000, 4096)
    exp, exp_pos, _ = oracle_mod.Port(m).trace(segs)
    g = native.BspGpu(m)
    g.set_option(5, 50_000)  # chunk size: 10 chunks, exercises buffer reuse
    pos = np.zeros((len(segs), 4), np.float32)
    res = g.trace(segs, out=np.zeros(len(segs), native.RESULT_DT), pos=pos)
    _compare(res, pos, exp, exp_pos, "host pipeline")
    d_segs = torch.from_numpy(segs).cuda()
    d_out = torch.zeros((len(segs), 4), dtype=torch.int32, device="cuda")
    d_pos = torch.zeros((len(segs), 4), dtype=torch.float32, device="cuda")
    g.set_stream(torch.cuda.current_stream().cuda_stream)
    g.trace(d_segs, out=d_out, pos=d_pos)
    torch.cuda.synchronize()
    res_d = d_out.cpu().numpy().view(native.RESULT_DT).reshape(-1)
    _compare(res_d, d_pos.cpu().numpy(), exp, exp_pos, "device path")
    assert g.info()["last_launches"] == 2            # the trace kernel + the (empty) exotic-ray kernel behind it
    g.close()


def test_device_generator_matches_numpy(native, maps):
    import torch
    from tools import mapgen, seggen
    m, _ = maps("rooms8")
    lo, hi = mapgen.seg_bounds(m)
    g = native.BspGpu(m)
    n, first = 100_000, 12_345_678_901
    buf = torch.zeros((n, 8), dtype=torch.float32, device="cuda")
    g.generate(buf, first, n, 0, 99, lo, hi)
    assert np.array_equal(bits(buf.cpu().numpy()), bits(seggen.uniform(99, first, n, lo, hi)))
    g.generate(buf, first, n, 1, 98, lo, hi, 100.0, 5000.0)
    assert np.array_equal(bits(buf.cpu().numpy()), bits(seggen.long_rays(98, first, n, lo, hi, 100.0, 5000.0)))
    views = seggen.make_views(5, 3, lo, hi)
    g.generate(buf, 1000, n, 2, 0, lo, hi, views=views, width=320, height=200, tan_half=1.0, far=32768.0)
    assert np.array_equal(bits(buf.cpu().numpy()), bits(seggen.camera(views, 320, 200, 1000, n, 1.0, 32768.0)))
    g.close()


def test_position_to_leaf(native, maps, oracle_mod):
    """batched BSP::position_to_leaf (bsp.cc:273-286)."""
    from tools import mapgen, seggen
    for name in ("kat1", "rooms8", "city10k"):
        m, path = maps(name)
        lo, hi = mapgen.seg_bounds(m)
        pts = np.ascontiguousarray(seggen.uniform(3, 0, 200_000, lo, hi)[:, :4])
        g = native.BspGpu(m)
        got = g.leaf(pts)
        assert np.array_equal(got, oracle_mod.Port(m).leaf(pts))
        if oracle_mod.have_reference():
            assert np.array_equal(got, oracle_mod.Reference(path).leaf(pts))
        g.close()


def test_compact_results_and_per_call_stream(native, maps, oracle_mod):
    """BSPGPU_OUT_COMPACT8 (8-byte records: t, flags | (plane+1) << 2) on the host and the device path, and
    bspgpu_trace_stream: a trace enqueued on the stream that is still generating its segments sees all of them."""
    import torch
    m, _ = maps("rooms8")
    segs = _segments(m, 150_000, 50_000, 20_000, 4096)
    exp, _, _ = oracle_mod.Port(m).trace(segs)
    want_bits = (exp["flags"] | ((exp["plane"] + 1).astype(np.uint32) << 2)).astype(np.uint32)
    g = native.BspGpu(m)
    g.set_option(5, 40_000)
    r8 = g.trace(segs, compact=True)
    assert r8.dtype == native.RESULT8_DT and np.array_equal(bits(r8["t"]), bits(exp["t"])) and np.array_equal(r8["bits"], want_bits)
    small = g.trace(segs[:100], compact=True)
    assert np.array_equal(small.view(np.uint32), r8[:100].view(np.uint32))
    d_segs = torch.from_numpy(segs).cuda()
    d8 = torch.zeros((len(segs), 2), dtype=torch.int32, device="cuda")
    g.trace(d_segs, out=d8, compact=True)
    assert np.array_equal(d8.cpu().numpy().view(np.uint32), r8.view(np.uint32).reshape(-1, 2))
    # BSPGPU_OUT_COMPACT4: one float per segment, t on a hit and T_MISS (2.0) on a miss -- {hit, t}, the reference's own outputs;
    # host path (chunked, small, pageable), device path, both kernels, and the multi-chunk pinned path of the packed layout
    want_t4 = np.where((exp["flags"] & 1) != 0, exp["t"], np.float32(native.T_MISS)).astype(np.float32)
    for variant in (1, 3):
        g.set_option(1, variant)
        r4 = g.trace(segs, compact=4)
        assert r4.dtype == np.float32 and np.array_equal(bits(r4), bits(want_t4)), f"compact4 variant {variant}"
        assert np.array_equal(bits(g.trace(segs[:77], compact=4)), bits(want_t4[:77]))
        d4 = torch.full((len(segs),), -1.0, dtype=torch.float32, device="cuda")
        g.trace(d_segs, out=d4, compact=4)
        assert np.array_equal(bits(d4.cpu().numpy()), bits(want_t4))
        packed = np.ascontiguousarray(np.concatenate((segs[:, 0:3], segs[:, 4:7]), axis=1))
        assert np.array_equal(bits(g.trace(packed, out=np.zeros(len(segs), np.float32), packed24=True, compact=4)), bits(want_t4))
    g.set_option(1, 0)
    assert ((r4 == np.float32(native.T_MISS)) == ((exp["flags"] & 1) == 0)).all() and (want_t4[(exp["flags"] & 1) != 0] <= 1.0).all()
    # ordering: fill the segments on a side stream, trace on that stream without any synchronisation in between
    side = torch.cuda.Stream()
    big = torch.from_numpy(np.tile(segs, (8, 1))).cuda()
    torch.cuda.synchronize()
    for _ in range(3):
        with torch.cuda.stream(side):
            work = torch.zeros_like(big)
            for _ in range(4):
                work.copy_(big * 0.5); work.add_(work)          # several kernels deep; ends up equal to `big`
            out = torch.zeros((len(big), 4), dtype=torch.int32, device="cuda")
            g.trace(work, out=out, async_=True)                 # api: current stream = side
        side.synchronize()
        got = out.cpu().numpy().view(native.RESULT_DT).reshape(-1)
        _compare(got[: len(segs)], None, exp, None, "trace ordered after its producer")
        _compare(got[-len(segs):], None, exp, None, "trace ordered after its producer (tail)")
    g.close()


def test_small_batches_and_latency_path(native, maps, oracle_mod):
    """a batch of one (BSP::trace_seg) up to a few thousand segments takes the single-stream bounce-buffer path and
    launches only as many CTAs as there is work for"""
    m, _ = maps("rooms8")
    segs = _segments(m, 3000, 600, 600, 200)
    assert len(segs) >= 4097
    exp, exp_pos, _ = oracle_mod.Port(m).trace(segs)
    g = native.BspGpu(m)
    for tiny in (1, 0):
        g.set_option(18, tiny)                       # BSPGPU_OPT_TINY_CALLS: up to 16 host segments are ONE kernel over the pinned buffer
        for n in (1, 2, 16, 17, 33, 1000, 4096, 4097):
            pos = np.zeros((n, 4), np.float32)
            res = g.trace(segs[:n], out=np.zeros(n, native.RESULT_DT), pos=pos)
            _compare(res, pos, exp[:n], exp_pos[:n], f"small batch {n} tiny {tiny}")
            assert g.hit_count() == int((exp["flags"][:n] & 1).sum())
            if n <= 4096:
                assert g.info()["grid_ctas"] <= -(-n // 32)
            if tiny and n <= 16:
                assert g.info()["grid_ctas"] == 1 and g.info()["last_launches"] == 1
    g.set_option(18, 1)
    # the tiny path in every layout: packed segments, 8- and 4-byte results, positions only; extreme coordinates too
    for n in (1, 7, 16):
        packed = np.ascontiguousarray(np.concatenate((segs[:n, 0:3], segs[:n, 4:7]), axis=1))
        r8 = g.trace(packed, packed24=True, compact=True)
        assert np.array_equal(bits(r8["t"]), bits(exp["t"][:n])) and np.array_equal(r8["bits"], exp["flags"][:n] | ((exp["plane"][:n] + 1).astype(np.uint32) << 2))
        r4 = g.trace(packed, packed24=True, compact=4)
        assert np.array_equal(bits(r4), bits(np.where((exp["flags"][:n] & 1) != 0, exp["t"][:n], np.float32(native.T_MISS)).astype(np.float32)))
        pos = g.trace(segs[:n], pos=np.zeros((n, 4), np.float32))
        assert np.array_equal(bits(pos[:, :3]), bits(exp_pos[:n]))
    from helpers import extreme_segments
    ext = extreme_segments(256, 2048)
    with np.errstate(all="ignore"):
        eexp, _, _ = oracle_mod.Port(m).trace(ext)
    for k in range(0, len(ext), 16):
        _compare(g.trace(ext[k:k + 16], out=np.zeros(16, native.RESULT_DT)), None, eexp[k:k + 16], None, "tiny path, extreme coordinates")
    g.close()


def test_pageable_callers(native, maps, oracle_mod):
    """plain (pageable) numpy arrays, large enough for the pageable-caller paths: bounced through the pinned ring by host
    threads (default), page-locked for the call, or left to the driver; 16-byte results + positions, and the packed / compact layouts"""
    m, _ = maps("rooms8")
    from tools import mapgen, seggen
    lo, hi = mapgen.seg_bounds(m)
    segs = seggen.uniform(31, 0, 1_500_007, lo, hi)                # 48 MB in, 24 MB out; not a multiple of the chunk
    exp, exp_pos, _ = oracle_mod.Port(m).trace(segs[:300_000])
    g = native.BspGpu(m)
    g.set_option(5, 200_000)
    outs = []
    for pageable in (2, 1, 0):
        g.set_option(17, pageable)
        pos = np.zeros((len(segs), 4), np.float32)
        outs.append((g.trace(segs, out=np.zeros(len(segs), native.RESULT_DT), pos=pos), pos))
        _compare(outs[-1][0][:300_000], pos[:300_000], exp, exp_pos, f"pageable option {pageable}")
    for res, pos in outs[1:]:
        assert np.array_equal(outs[0][0].view(np.uint32), res.view(np.uint32)) and np.array_equal(bits(outs[0][1]), bits(pos))
    # the reference's own argument layout (packed vec3 pairs) in, compact records out, through the ring; fewer chunks than ring slots too
    g.set_option(17, 2)
    packed = np.ascontiguousarray(np.hstack((segs[:, 0:3], segs[:, 4:7])))
    for chunk in (200_000, 1_000_000, 2_000_000):
        g.set_option(5, chunk)
        r8 = g.trace(packed, out=np.zeros(len(segs), native.RESULT8_DT), packed24=True, compact=True)
        assert np.array_equal(bits(r8["t"]), bits(outs[0][0]["t"]))
        assert np.array_equal(r8["bits"], outs[0][0]["flags"] | ((outs[0][0]["plane"] + 1).astype(np.uint32) << 2))
    with pytest.raises(native.BspGpuError):
        g.set_option(17, 3)
    g.close()


def test_segment_binning(native, maps, oracle_mod):
    """BSPGPU_OPT_BIN_SEGMENTS: the counting-sort pre-pass changes the order lanes pick segments up in,
    never a result, and every result lands at its segment's own index"""
    import torch
    for name in ("city10k", "rooms8"):
        m, _ = maps(name)
        segs = _segments(m, 200_000, 200_000, 200_000, 8192)
        exp, exp_pos, _ = oracle_mod.Port(m).trace(segs)
        g = native.BspGpu(m)
        g.set_option(16, 1)
        d_segs = torch.from_numpy(segs).cuda()
        out = torch.zeros((len(segs), 4), dtype=torch.int32, device="cuda")
        pos = torch.zeros((len(segs), 4), dtype=torch.float32, device="cuda")
        for launch in (1 << 30, 100_000):
            g.set_option(12, launch)
            out.zero_(); pos.zero_()
            g.trace(d_segs, out=out, pos=pos)
            assert g.info()["last_prepass_ms"] > 0
            _compare(out.cpu().numpy().view(native.RESULT_DT).reshape(-1), pos.cpu().numpy(), exp, exp_pos, f"{name} binned, launch cap {launch}")
            assert g.hit_count() == int((exp["flags"] & 1).sum())
        g.close()


def test_clusters_and_pvs(native, maps, oracle_mod):
    """batched position_to_leaf -> BSP_Leaf::cluster -> PVS bit (reference bsp_renderer.cc:147-168, bsp.h:129-134),
    on the q3-style map that carries clusters and a real vis lump, loaded from its 17-slot file"""
    from tools import mapgen, seggen
    m, path = maps("q3style")
    lo, hi = mapgen.seg_bounds(m)
    pts = np.ascontiguousarray(seggen.uniform(9, 0, 100_000, lo, hi)[:, :3])
    g = native.BspGpu(path)
    nclusters = int(m.vis[:4].view(np.uint32)[0]); csize = int(m.vis[4:8].view(np.uint32)[0])
    assert g.info()["num_clusters"] == nclusters > 8 and g.info()["num_leaves"] == len(m.leaves)
    rows = m.vis[8:].reshape(nclusters, csize)
    leaf_exp = oracle_mod.Port(m).leaf(np.ascontiguousarray(np.hstack((pts, np.zeros((len(pts), 1), np.float32)))))
    cl_exp = m.leaves["cluster"][leaf_exp]
    assert (cl_exp >= 0).any() and (cl_exp < 0).any()
    for frm in (-1, 0, nclusters // 2, nclusters - 1):
        leaf, cluster, vis = g.leaf_clusters(pts, from_cluster=frm, want_visible=True)
        assert np.array_equal(leaf, leaf_exp) and np.array_equal(cluster, cl_exp)
        if frm < 0:
            want = np.ones(len(pts), np.uint8)
        else:
            c = np.maximum(cl_exp, 0)
            want = ((rows[frm, c >> 3] >> (c & 7)) & 1).astype(np.uint8) * (cl_exp >= 0)
        assert np.array_equal(vis, want), frm
    # the renderer's loop: which leaves are drawn from a viewpoint
    inside = pts[cl_exp >= 0][:20]
    for p in inside:
        vis, vc = g.visible_leaves(p)
        assert vc >= 0
        c = m.leaves["cluster"]
        want = np.where(c >= 0, (rows[vc, np.maximum(c, 0) >> 3] >> (np.maximum(c, 0) & 7)) & 1, 0).astype(np.uint8)
        assert np.array_equal(vis, want)
    vis, vc = g.visible_leaves(pts[cl_exp < 0][0])
    assert vc == -1 and vis.all()
    with pytest.raises(native.BspGpuError):
        g.leaf_clusters(pts[:4], from_cluster=nclusters, want_visible=True)
    g.close()


@pytest.mark.parametrize("mapname,kind,n", [("rooms8", "uniform", 64 * 1024 * 1024), ("city10k", "long", 128 * 1024 * 1024)])
def test_full_size_properties(native, maps, oracle_mod, mapname, kind, n):
    """BASELINE.json sizes (C2: 64 Mi uniform segments on rooms8; C3: one GPU's 128 Mi-segment shard of the
    1 Gi long segments on city10k), generated on the device.  Size-independent properties: every staging
    variant and both kernels give the same bits and the same hit count; tracing twice is idempotent; the
    hit counter equals the number of hit flags; a strided sample matches the oracle."""
    import torch
    from tools import mapgen, seggen
    m, _ = maps(mapname)
    lo, hi = mapgen.seg_bounds(m)
    diag = float(np.sqrt(((hi - lo).astype(np.float64) ** 2).sum()))
    g = native.BspGpu(m)
    g.set_stream(torch.cuda.current_stream().cuda_stream)
    segs = torch.empty((n, 8), dtype=torch.float32, device="cuda")
    first = 3 * n                                    # somewhere inside the global index space, like a rank's shard
    if kind == "uniform":
        g.generate(segs, first, n, 0, 2, lo, hi)
    else:
        g.generate(segs, first, n, 1, 3, lo, hi, 0.25 * diag, diag)
    configs = [(0, -1), (3, -1), (3, 0), (0, 4)]       # (variant, shared-memory stack slots)
    ref_out, ref_hits = None, None
    for variant, sched in configs:
        g.set_option(1, variant); g.set_option(15, sched)
        out = torch.empty((n, 4), dtype=torch.int32, device="cuda")
        g.trace(segs, out=out)
        hits = g.hit_count()
        if ref_out is None:
            ref_out, ref_hits = out, hits
            assert int((out[:, 3] & 1).sum().item()) == hits
            again = torch.empty_like(out)
            g.trace(segs, out=again)
            assert torch.equal(out, again)                                        # idempotent
            del again
        else:
            assert hits == ref_hits, (variant, sched)
            assert torch.equal(out, ref_out), (variant, sched)
            del out
    stride = 61                                      # >= 1 M rows of each configuration against the reference object
    idx = torch.arange(0, n, stride, device="cuda")
    sample = segs[idx].cpu().numpy()
    gi = first + np.arange(0, n, stride, dtype=np.uint64)
    # the sampled device segments are the generator's (seed, global index) floats
    want = np.vstack([seggen.uniform(2, int(i), 1, lo, hi) if kind == "uniform" else seggen.long_rays(3, int(i), 1, lo, hi, 0.25 * diag, diag)
                      for i in gi[:64]])
    assert np.array_equal(bits(sample[:64]), bits(want))
    assert len(sample) >= 1_000_000
    got = ref_out[idx].cpu().numpy().view(native.RESULT_DT).reshape(-1)
    if oracle_mod.have_reference():
        # the unmodified reference (work_queue-threaded): {hit, t} are its outputs
        ref = oracle_mod.Reference(maps(mapname)[1])
        r = ref.trace(sample, threads=max(1, len(__import__("os").sched_getaffinity(0))))
        bad = ((got["flags"] & 1).astype(np.uint8) != r["hit"]) | (bits(got["t"]) != bits(r["t"]))
        assert not bad.any(), f"{mapname} {n}: {int(bad.sum())} of {len(sample)} sampled rows differ from the reference; first {int(np.nonzero(bad)[0][0])}"
    exp, _, _ = oracle_mod.Port(m).trace(sample[:200_000])
    _compare(got[:200_000], None, exp, None, f"{mapname} {n} strided sample (extension fields vs port)")
    g.close()


def test_multi_launch_and_concurrent_callers(native, maps, oracle_mod):
    """n larger than one launch's segment cap is split into several launches; calls from several
    host threads on one handle are serialised by the handle's mutex."""
    import threading
    m, _ = maps("rooms8")
    segs = _segments(m, 150_000, 50_000, 20_000, 4096)
    exp, exp_pos, _ = oracle_mod.Port(m).trace(segs)
    g = native.BspGpu(m)
    g.set_option(12, 40_000)                       # ~6 launches per call
    res = g.trace(segs)
    _compare(res, None, exp, exp_pos, "multi-launch")
    assert g.info()["last_launches"] == 2 * -(-len(segs) // 40_000)      # every trace launch is followed by the (usually empty) exotic-ray kernel
    assert g.hit_count() == int((exp["flags"] & 1).sum())
    outs = [None] * 4

    def work(i):
        outs[i] = g.trace(segs)
    threads = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [t.start() for t in threads]; [t.join() for t in threads]
    for o in outs:
        _compare(o, None, exp, exp_pos, "concurrent callers")
    g.close()


def test_errors(native, maps):
    m, _ = maps("kat0")
    g = native.BspGpu(m)
    with pytest.raises(native.BspGpuError):
        g.set_option(999, 1)
    lib = native.load_library()
    assert lib.bspgpu_trace(g.h, None, 5, None, 0) < 0
    assert b"null" in lib.bspgpu_last_error()
    g.close()


@pytest.mark.parametrize("name", ["kat0", "kat1", "rooms8", "tilted"])
def test_golden_reference_outputs(native, name):
    """committed fixtures = outputs of the unmodified reference (tests/golden/make_golden.py)."""
    from helpers import load_golden
    m, segs, hit, t, pos = load_golden(name)
    g = native.BspGpu(m)
    p = np.zeros((len(segs), 4), np.float32)
    res = g.trace(segs, out=np.zeros(len(segs), native.RESULT_DT), pos=p)
    assert np.array_equal((res["flags"] & 1).astype(np.uint8), hit)
    assert np.array_equal(bits(res["t"]), bits(t))
    assert np.array_equal(bits(p[:, :3]), bits(pos))
    g.close()


@pytest.mark.parametrize("name", ["extreme_rooms8", "extreme_odd"])
def test_golden_reference_outputs_for_extreme_floats(native, name):
    """the unmodified reference's answers for infinite / NaN / huge / denormal coordinates (and a solid brush without sides)"""
    from helpers import check_against_extreme_golden, load_golden
    m, segs, hit, t, pos = load_golden(name)
    g = native.BspGpu(m)
    for variant in (1, 3):
        g.set_option(1, variant)
        p = np.zeros((len(segs), 4), np.float32)
        res = g.trace(segs, out=np.zeros(len(segs), native.RESULT_DT), pos=p)
        check_against_extreme_golden(res["flags"] & 1, res["t"], p[:, :3], hit, t, pos, f"{name} variant {variant}")
    g.close()


def test_experiment_options_are_rejected_in_the_product_build(native, maps):
    """the measured-slower round-1 kernels are not in libbspgpu.so (only in -DBSPGPU_EXPERIMENTS builds): asking for
    them fails loudly instead of silently running something else; so do values that could hang a kernel."""
    m, _ = maps("kat1")
    g = native.BspGpu(m)
    for opt, val in ((9, 1), (9, 3), (14, 1), (8, -5), (8, 7), (8, 1 << 40), (6, 0), (6, 33), (11, 0), (10, 0), (5, 10), (15, 9)):
        with pytest.raises(native.BspGpuError):
            g.set_option(opt, val)
    g.set_option(8, 0)                               # 0 = default
    g.set_option(1, 2)                               # TOPTREE is accepted as a value but needs an experiments build to run
    with pytest.raises(native.BspGpuError, match="EXPERIMENTS"):
        g.trace(np.zeros((4, 8), np.float32))
    g.set_option(1, 0)
    assert len(g.trace(np.ones((4, 8), np.float32))) == 4
    g.close()


def test_odd_map_on_gpu(native, oracle_mod):
    """brush without sides, leaf with only non-solid brushes, unreachable node (see test_hostsim)."""
    from test_hostsim import _odd_map
    m = _odd_map()
    rs = np.random.Generator(np.random.PCG64(5))
    segs = np.zeros((50_000, 8), np.float32)
    segs[:, 0:3] = rs.uniform(-80, 80, (50_000, 3)); segs[:, 4:7] = rs.uniform(-80, 80, (50_000, 3))
    segs[:1000, 0:3] = np.round(segs[:1000, 0:3] / 16) * 16
    exp, exp_pos, _ = oracle_mod.Port(m).trace(segs)
    g = native.BspGpu(m)
    for variant in (1, 3):
        g.set_option(1, variant)
        pos = np.zeros((len(segs), 4), np.float32)
        res = g.trace(segs, out=np.zeros(len(segs), native.RESULT_DT), pos=pos)
        _compare(res, pos, exp, exp_pos, f"odd map variant {variant}")
    g.close()


@pytest.mark.parametrize("variant", ["global"])
def test_parity_mega200k_short_moves(native, maps, oracle_mod, variant):
    """C5: ~200k-brush map (54 MB flattened image, L2-resident), short player-movement segments."""
    from tools import mapgen, seggen
    m, _ = maps("mega200k")
    lo, hi = mapgen.seg_bounds(m)
    segs = np.vstack((seggen.edge_cases(m, 8192, 9), seggen.long_rays(5, 0, 400_000, lo, hi, 1.0, 32.0),
                      seggen.uniform(6, 0, 50_000, lo, hi)))
    exp, exp_pos, _ = oracle_mod.Port(m).trace(segs)
    g = native.BspGpu(m)
    assert g.info()["map_bytes"] > 40 * 1024 * 1024 and g.info()["tree_depth"] <= 32
    g.set_option(1, VARIANTS[variant])
    pos = np.zeros((len(segs), 4), np.float32)
    res = g.trace(segs, out=np.zeros(len(segs), native.RESULT_DT), pos=pos)
    _compare(res, pos, exp, exp_pos, f"mega200k/{variant}")
    g.close()


def test_parity_camera_grid(native, maps, oracle_mod):
    """C4: coherent pinhole-grid segments (row-major) and the same segments shuffled."""
    from tools import mapgen, seggen
    m, _ = maps("city10k")
    lo, hi = mapgen.seg_bounds(m)
    views = seggen.make_views(4, 3, lo, hi)
    segs = seggen.camera(views, 480, 270, 0, 3 * 480 * 270)
    exp, exp_pos, _ = oracle_mod.Port(m).trace(segs)
    g = native.BspGpu(m)
    res = g.trace(segs)
    _compare(res, None, exp, exp_pos, "camera grid")
    perm = np.random.Generator(np.random.PCG64(1)).permutation(len(segs))
    res_s = g.trace(np.ascontiguousarray(segs[perm]))
    _compare(res_s, None, exp[perm], None, "camera grid shuffled")
    g.close()


def test_depth_frame_demo(native, maps, oracle_mod):
    """tools/render_depth.py: a camera frame through one batched call (caller integration, 8f N2)."""
    from tools import mapgen, render_depth, seggen
    m, path = maps("city10k")
    lo, hi = mapgen.seg_bounds(m)
    views = seggen.make_views(4, 2, lo, hi)
    g = native.BspGpu(path)
    w, h = 160, 90
    for device in (True, False):
        t, hit = render_depth.render(g, views, w, h, view=1, device=device)
        exp, _, _ = oracle_mod.Port(m).trace(seggen.camera(views, w, h, w * h, w * h))
        assert np.array_equal(bits(t.reshape(-1)), bits(exp["t"]))
        assert np.array_equal(hit.reshape(-1), (exp["flags"] & 1).astype(bool))
    g.close()


def test_extreme_float_inputs(native, maps, oracle_mod):
    """huge / infinite / denormal / NaN coordinates: the reference just lets IEEE arithmetic run (NaN falls
    out of every comparison), and so must the kernels.  Result records must still agree bit for bit; `pos`
    is compared value-wise with NaN == NaN because NaN payloads differ between SSE and the GPU."""
    from helpers import extreme_segments
    m, _ = maps("rooms8")
    n = 60_000
    segs = extreme_segments(n, 2048)
    with np.errstate(all="ignore"):
        exp, exp_pos, _ = oracle_mod.Port(m).trace(segs)
    g = native.BspGpu(m)
    for variant in (1, 3):
        g.set_option(1, variant)
        pos = np.zeros((n, 4), np.float32)
        res = g.trace(segs, out=np.zeros(n, native.RESULT_DT), pos=pos)
        _compare(res, None, exp, None, f"extreme inputs variant {variant}")
        assert np.array_equal(pos[:, :3], exp_pos, equal_nan=True)
        finite = np.isfinite(exp_pos).all(axis=1)
        assert np.array_equal(bits(pos[finite, :3]), bits(exp_pos[finite]))
    g.close()


def test_mixed_pointer_kinds_and_pos_only(native, maps, oracle_mod):
    """segments in HBM + results to host, segments on host + results in HBM, and pos-only output."""
    import ctypes as C
    import torch
    m, _ = maps("rooms8")
    segs = _segments(m, 120_000, 40_000, 20_000, 4096)
    n = len(segs)
    exp, exp_pos, _ = oracle_mod.Port(m).trace(segs)
    g = native.BspGpu(m)
    g.set_option(5, 30_000)
    lib = native.load_library()
    d_segs = torch.from_numpy(segs).cuda()
    # device segments -> host results
    res = np.zeros(n, native.RESULT_DT)
    assert lib.bspgpu_trace(g.h, d_segs.data_ptr(), n, res.ctypes.data, 1) == 0          # BSPGPU_SEGS_DEVICE
    _compare(res, None, exp, None, "device segs, host out")
    # host segments -> device results
    d_out = torch.zeros((n, 4), dtype=torch.int32, device="cuda")
    assert lib.bspgpu_trace(g.h, segs.ctypes.data, n, d_out.data_ptr(), 2) == 0           # BSPGPU_OUT_DEVICE
    torch.cuda.synchronize()
    _compare(d_out.cpu().numpy().view(native.RESULT_DT).reshape(-1), None, exp, None, "host segs, device out")
    # pos only
    pos = np.zeros((n, 4), np.float32)
    assert lib.bspgpu_trace_pos(g.h, segs.ctypes.data, n, None, pos.ctypes.data, 0) == 0
    assert np.array_equal(bits(pos[:, :3]), bits(exp_pos)) and np.array_equal(bits(pos[:, 3]), bits(exp["t"]))
    # ASYNC is refused for host pointers
    assert lib.bspgpu_trace(g.h, segs.ctypes.data, n, res.ctypes.data, 8) < 0
    g.close()


@pytest.mark.parametrize("name", ["rooms8", "city10k", "tilted", "mega200k"])
def test_start_grid(native, maps, oracle_mod, name):
    """BSPGPU_OPT_START_GRID: walks that begin below the root are bit-identical to walks from node 0."""
    from helpers import grid_adversarial_segments
    from tools import mapgen, seggen
    m, _ = maps(name)
    lo, hi = mapgen.seg_bounds(m)
    segs = np.vstack((_segments(m, 50_000, 30_000, 150_000, 4096), seggen.long_rays(77, 0, 200_000, lo, hi, 0.01, 8.0),
                      grid_adversarial_segments(m, 300_000)))
    exp, exp_pos, _ = oracle_mod.Port(m).trace(segs)
    g = native.BspGpu(m)
    assert g.info()["start_grid_cells"] > 100_000
    for variant in ((1, 3) if g.info()["staged_bytes"] < 200 * 1024 else (3,)):
        g.set_option(1, variant)
        for grid in (1, 0):
            g.set_option(13, grid)
            pos = np.zeros((len(segs), 4), np.float32)
            res = g.trace(segs, out=np.zeros(len(segs), native.RESULT_DT), pos=pos)
            _compare(res, pos, exp, exp_pos, f"{name} variant {variant} grid {grid}")
    g.close()


@pytest.mark.parametrize("name", ["kat1", "rooms8", "city10k", "tilted"])
def test_axis_node_records(native, maps, oracle_mod, name):
    """node planes with a bitwise +unit normal are stored as their axis and plain rays skip the dot products
    (map_layout.h: NodeRec, trace_core.h: ray_is_plain); rays with zero / non-finite components are traced by the
    lane that draws them with the general arithmetic.  Same bits as an image flattened with general planes only,
    as the oracle, in both staging variants, with and without the start grid."""
    from helpers import extreme_segments, grid_adversarial_segments
    from tools import mapgen, seggen
    m, _ = maps(name)
    lo, hi = mapgen.seg_bounds(m)
    segs = np.vstack((_segments(m, 50_000, 30_000, 150_000, 4096), seggen.long_rays(79, 0, 200_000, lo, hi, 0.01, 8.0),
                      extreme_segments(40_000, float(np.abs(np.concatenate((lo, hi))).max()))))
    # a sprinkling of exact zeros (both signs) in otherwise ordinary segments
    rs = np.random.Generator(np.random.PCG64(3))
    z = rs.integers(0, len(segs), 60_000)
    segs[z, rs.integers(0, 3, len(z))] = np.array([0.0, -0.0], np.float32)[rs.integers(0, 2, len(z))]
    g = native.BspGpu(m)
    if g.info()["start_grid_cells"]:
        segs = np.vstack((segs, grid_adversarial_segments(m, 200_000)))
    with np.errstate(all="ignore"):
        exp, exp_pos, _ = oracle_mod.Port(m).trace(segs)
    gg = native.BspGpu(m, general_planes_only=True)
    assert gg.info()["num_general_planes"] == gg.info()["num_nodes"]
    if name in ("rooms8", "city10k"):
        assert g.info()["num_general_planes"] == 0
    for h, what in ((g, "axis records"), (gg, "general planes only")):
        for variant in ((1, 3) if h.info()["staged_bytes"] < 200 * 1024 else (3,)):
            for grid, steps in ((1, 8), (0, 1)):
                h.set_option(1, variant); h.set_option(13, grid); h.set_option(7, steps)
                res = h.trace(segs)
                _compare(res, None, exp, None, f"{name} {what} variant {variant} grid {grid} steps {steps}")
    g.close(); gg.close()


@pytest.mark.parametrize("name", ["kat1", "rooms8", "city10k", "tilted", "mega200k"])
def test_smem_stack(native, maps, oracle_mod, name):
    """BSPGPU_OPT_SMEM_STACK: the first 2..8 traversal-stack slots of every thread in shared memory, deeper ones in local
    memory -- same bits as the all-local stack (2 slots: most pushes overflow; tilted is 42 levels deep)."""
    from tools import mapgen, seggen
    m, _ = maps(name)
    lo, hi = mapgen.seg_bounds(m)
    segs = np.vstack((_segments(m, 50_000, 30_000, 150_000, 4096), seggen.long_rays(81, 0, 300_000, lo, hi, 0.01, 8.0)))
    exp, exp_pos, _ = oracle_mod.Port(m).trace(segs)
    g = native.BspGpu(m)
    g.set_option(1, 3)
    for slots, steps in ((2, 8), (4, 1), (6, 8), (8, 3), (0, 8)):
        g.set_option(15, slots); g.set_option(7, steps)
        pos = np.zeros((len(segs), 4), np.float32)
        res = g.trace(segs, out=np.zeros(len(segs), native.RESULT_DT), pos=pos)
        assert g.info()["smem_bytes"] >= slots * 256 * 16 and g.info()["smem_stack_slots"] == slots
        _compare(res, pos, exp, exp_pos, f"{name} smem stack {slots} steps {steps}")
    if g.info()["staged_bytes"] < 140 * 1024:
        # the same behind a staged map (SMEM variant, one 1024-thread CTA per SM); explicit request only
        g.set_option(1, 1)
        for slots in (2, 6):
            g.set_option(15, slots)
            pos = np.zeros((len(segs), 4), np.float32)
            res = g.trace(segs, out=np.zeros(len(segs), native.RESULT_DT), pos=pos)
            assert g.info()["variant"] == 1 and g.info()["smem_bytes"] > g.info()["staged_bytes"]
            _compare(res, pos, exp, exp_pos, f"{name} staged map + smem stack {slots}")
    g.close()


@pytest.mark.parametrize("variant", list(VARIANTS))
@pytest.mark.parametrize("mapname", ["kat1", "rooms8", "city10k", "q3style"])
def test_queue_scheduler(native, maps, oracle_mod, mapname, variant):
    """BSPGPU_OPT_SCHEDULER = 5 (pool_kernels.cuh: rays wait in CTA-wide queues, not in lanes) changes where a ray waits,
    never what is computed: bit-exact against the oracle, for knob settings that force every queue path."""
    m, path = maps(mapname)
    g = native.BspGpu(m)
    if variant == "smem" and g.info()["staged_bytes"] > 120 * 1024:
        g.close()
        pytest.skip("no room for the ray pool beside the staged map")
    g.set_option(1, VARIANTS[variant]); g.set_option(9, 5)
    try:
        g.trace(np.ones((4, 8), np.float32))
    except native.BspGpuError as e:
        assert "EXPERIMENTS" in str(e)
        g.close()
        pytest.skip("the queue scheduler measured slower than the shipped kernel and is only in -DBSPGPU_EXPERIMENTS builds (profiles/README.md)")
    segs = _segments(m)
    exp, exp_pos, ctr = oracle_mod.Port(m).trace(segs)
    for refill, max_steps, side_steps, leaf_thr in [(8, 0, 3, 16), (1, 1, 1, 1), (32, 1000, 1000, 32), (8, 4, 2, 8)]:
        g.set_option(6, refill); g.set_option(7, max_steps); g.set_option(10, side_steps); g.set_option(11, leaf_thr)
        pos = np.zeros((len(segs), 4), np.float32)
        res = g.trace(segs, out=np.zeros(len(segs), native.RESULT_DT), pos=pos)
        _compare(res, pos, exp, exp_pos, f"queues {mapname}/{variant} knobs {refill}/{max_steps}/{side_steps}/{leaf_thr}")
        assert g.hit_count() == ctr["hits"]
    for n in (1, 33, 4097):
        res = g.trace(segs[:n].copy())
        _compare(res, None, exp[:n], exp_pos[:n], f"queues {mapname}/{variant} n={n}")
    g.close()
