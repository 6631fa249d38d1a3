#!/usr/bin/env python
"""bench.py -- BSP segment traces/second on N B200s (BASELINE.json's metric).

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus 8 --steps 10 --warmup 3
    python bench.py --impl reference ...       # the reference's own CPU implementation, same config

A step = one pass of the hot path (n x BSP::trace_seg, reference bsp.cc:255-271) over one batch
of synthetic segments.  Default workload = the configuration the metric's target is quoted on
(BASELINE.json configs[2]): the ~10k-brush `city10k` map, "long" segments (start uniform,
isotropic direction, length in [1/4,1] x AABB diagonal), 128 Mi segments per GPU per step, so
N=8 is the full 1B-segment configuration and per-GPU work is fixed (weak scaling).  Segments are
generated on each GPU from (seed, global index); the map is replicated; shards need no data-path
collective -- NCCL carries only the per-step hit-count all-reduce and a small result gather.

`value`   : whole-job traces/s with segments and results resident in HBM, CUDA-event timed,
            max over ranks.
`e2e`     : the same metric through the public C-ABI call with HOST (pinned) buffers: every
            step copies that step's segments H2D and its results D2H inside the timed region.
`roofline`: algorithmic bytes per trace (BASELINE.md section 5 formula; visit-count means from the
            counter-instrumented CPU restatement) x traces per launch / mean launch time.
`cpu_baseline`: the unmodified reference (oracle/_ref), work_queue-threaded on all host cores,
            on a bounded sample of the same workload (rank 0, N=1 only).  This leg is the ONLY place
            the arm touches oracle/: it also refreshes the visit counts and spot-checks a sample of the
            timed run's results; without it (N>1, --no-cpu) the committed counts in
            profiles/bytes_per_trace.json are used and no checker runs.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (map, kind, seed, (len_lo, len_hi) as fractions of the diagonal or absolute units)
    "city10k-long": dict(map="city10k", kind="long", seed=3, len_frac=(0.25, 1.0)),
    "rooms8-uniform": dict(map="rooms8", kind="uniform", seed=2),
    "mega200k-short": dict(map="mega200k", kind="long", seed=5, len_abs=(1.0, 32.0)),
    "city10k-camera": dict(map="city10k", kind="camera", seed=4, width=3840, height=2160, views=64),
    # the same segments in a seeded random order ("also report a shuffled order for the divergence contrast")
    "city10k-camera-shuffled": dict(map="city10k", kind="camera", seed=4, width=3840, height=2160, views=64, shuffle=True),
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def workload_setup(name):
    """map + generator description shared by both arms."""
    from tools import mapgen
    w = dict(WORKLOADS[name])
    m, path = mapgen.get_map(w["map"])
    lo, hi = mapgen.seg_bounds(m)
    diag = float(np.sqrt(((hi - lo).astype(np.float64) ** 2).sum()))
    if "len_frac" in w:
        w["len"] = (w["len_frac"][0] * diag, w["len_frac"][1] * diag)
    elif "len_abs" in w:
        w["len"] = w["len_abs"]
    w.update(m=m, path=path, lo=lo, hi=hi, diag=diag)
    if w["kind"] == "camera":
        from tools import seggen
        w["view_array"] = seggen.make_views(w["seed"], w["views"], lo, hi)
    return w


def host_segments(w, first, count):
    """numpy generator (bit-identical to the device generator) -- used by the CPU arms."""
    from tools import seggen
    if w["kind"] == "uniform":
        return seggen.uniform(w["seed"], first, count, w["lo"], w["hi"])
    if w["kind"] == "long":
        return seggen.long_rays(w["seed"], first, count, w["lo"], w["hi"], w["len"][0], w["len"][1])
    return seggen.camera(w["view_array"], w["width"], w["height"], first, count)


def device_segments(g, w, buf, first, count):
    if w["kind"] == "uniform":
        g.generate(buf, first, count, 0, w["seed"], w["lo"], w["hi"])
    elif w["kind"] == "long":
        g.generate(buf, first, count, 1, w["seed"], w["lo"], w["hi"], w["len"][0], w["len"][1])
    else:
        total = w["views"] * w["width"] * w["height"]
        g.generate(buf, first % total, count, 2, 0, w["lo"], w["hi"], views=w["view_array"], width=w["width"], height=w["height"])
        if w.get("shuffle"):
            import torch
            gen = torch.Generator(device=buf.device); gen.manual_seed(w["seed"] + first)
            buf.copy_(buf[torch.randperm(count, device=buf.device, generator=gen)])


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    FIELDS = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(gpu_index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def window(self, t0, t1):
        self.t0, self.t1 = t0, t1

    def stop(self):
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if not (getattr(self, "t0", 0) - 0.05 <= ts <= getattr(self, "t1", 1e30) + 0.15):
                continue
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); smax.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons), "samples": len(sm)}


def host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return os.cpu_count() or 1


def bind_to_gpu_numa_node(gpu_index):
    """best effort: run this rank (and first-touch its pinned buffers) on the CPUs of the NUMA node the GPU hangs off."""
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(gpu_index)],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        if bus.startswith("0000"):
            bus = bus[4:]                                  # 00000000:1b:00.0 -> 0000:1b:00.0
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def algorithmic_bytes(w, counters=None):
    """B = 48 + 28 n_node + 8 n_leaf + 12 n_lb + 8 n_solid + 20 n_side.
    Visit-count means come from the counter-instrumented CPU restatement: measured live by the cpu_baseline
    leg when it runs (N=1), otherwise the committed means of profiles/bytes_per_trace.json (same generator,
    same seed -- written from earlier cpu_baseline legs).  The product path never touches the oracle."""
    if counters is not None:
        n = max(1, counters["traces"])
        mean = {k: counters[k] / n for k in ("nodes", "leaves", "leaf_brushes", "solid", "sides")}
        b_map = 28 * mean["nodes"] + 8 * mean["leaves"] + 12 * mean["leaf_brushes"] + 8 * mean["solid"] + 20 * mean["sides"]
        return {"b_io": 48, "b_map": b_map, "b_total": 48 + b_map, "means": mean, "source": "cpu_baseline leg (oracle port counters on 200000 segments of this workload)"}
    rec = json.load(open(os.path.join(ROOT, "profiles", "bytes_per_trace.json")))[w["map"] + "/" + w["kind"]]
    return {"b_io": 48, "b_map": rec["b_total"] - 48, "b_total": rec["b_total"], "means": rec["means"], "source": "profiles/bytes_per_trace.json"}


def cpu_reference(w, seconds_budget=12.0, threads=None, parity=None):
    """THE cpu_baseline LEG -- the only place this arm touches oracle/: the unmodified reference on the host
    cores (single-threaded and work_queue-threaded), the port's visit counters for the roofline numerator, and
    a spot check of `parity` = (segments, gpu results) sampled from the timed run.  Runs after every timed region."""
    import oracle
    extras = {}
    port = oracle.Port(w["m"])
    _, _, ctr = port.trace(host_segments(w, 0, 200_000), want_pos=False)
    extras["counters"] = ctr
    if parity is not None:
        sseg, got = parity
        exp, _, _ = port.trace(sseg, want_pos=False)
        bad = int((got.view(np.uint32).reshape(-1, 4) != exp.view(np.uint32).reshape(-1, 4)).any(axis=1).sum())
        extras["parity"] = {"checked": int(len(sseg)), "mismatches": bad}
    threads = threads or host_threads()
    kind = "reference" if oracle.have_reference() else "port"
    if kind == "reference":
        ref = oracle.Reference(w["path"])
        probe = host_segments(w, 0, 200_000)
        ref.trace(probe, threads=1)
        rate1 = len(probe) / ref.seconds
        n1 = int(max(200_000, min(8_000_000, rate1 * seconds_budget * 0.3)))
        segs = host_segments(w, 0, n1)
        ref.trace(segs, threads=1)
        single = n1 / ref.seconds
        nmt = int(max(n1, min(64_000_000, single * threads * seconds_budget * 0.35)))
        if nmt != n1:
            segs = host_segments(w, 0, nmt)
        ref.trace(segs, threads=threads)          # warm-up: starts the reference's worker threads
        best = 0.0
        for _ in range(3):                        # noisy -> warm up + best of k
            ref.trace(segs, threads=threads)
            best = max(best, nmt / ref.seconds)
        return extras, {"value": best, "unit": "traces/s", "cores": threads, "kind": kind,
                "single_thread": single,
                "sample": f"{nmt} segments of this workload through BSP::trace_seg, reference work_queue with {threads - 1} workers + caller, "
                          f"best of 3; single-thread loop on {n1} segments = {single:.3e} traces/s"}
    segs = host_segments(w, 0, 2_000_000)
    t0 = time.time(); port.trace(segs, want_pos=False); dt = time.time() - t0
    return extras, {"value": len(segs) / dt, "unit": "traces/s", "cores": 1, "kind": "port", "single_thread": len(segs) / dt,
            "sample": f"{len(segs)} segments through oracle/restate.c, 1 thread (reference object not built on this box)"}


# --------------------------------------------------------------------------------------- arms

def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = workload_setup(a.workload)
    import oracle
    threads = host_threads()
    ref = oracle.Reference(w["path"]) if oracle.have_reference() else None
    kind = "reference" if ref else "port"
    # size a step so the whole run stays within a few minutes
    probe = host_segments(w, 0, 400_000)
    if ref:
        ref.trace(probe, threads=threads); ref.trace(probe, threads=threads)
        rate = len(probe) / ref.seconds
    else:
        port = oracle.Port(w["m"]); t0 = time.time(); port.trace(probe, want_pos=False); rate = len(probe) / (time.time() - t0)
        threads = 1
    per_step = int(max(400_000, min(32_000_000, rate * 4.0)))
    segs = host_segments(w, 0, per_step)
    times = []
    for i in range(a.warmup + a.steps):
        if ref:
            ref.trace(segs, threads=threads); dt = ref.seconds
        else:
            t0 = time.time(); port.trace(segs, want_pos=False); dt = time.time() - t0
        if i >= a.warmup:
            times.append(dt)
    total = sum(times)
    value = per_step * a.steps / total
    sample = f"{per_step} segments per step (bounded sample of {a.workload}), {'work_queue-threaded' if ref else 'single-thread port'}"
    line = {"impl": "reference", "metric": "bsp_segment_traces_per_sec", "value": value, "unit": "traces/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * total / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": a.workload, "map": w["map"], "segments_per_step": per_step, "host_threads": threads},
            "cpu_baseline": {"value": value, "unit": "traces/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "traces/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def level_peaks():
    """measured denominators of the level that holds the map (tools/microbench on one B200, profiles/r2_microbench.json):
    L2 read bandwidth with L1 bypassed (ld.global.cg over a 96 MB resident set) and the shared-memory bandwidth of the
    trace's own access pattern (every lane a different dense 16-byte record, one LDS.128)."""
    d = json.load(open(os.path.join(ROOT, "profiles", "r2_microbench.json")))
    return {"l2": (max(d[k] for k in d if k.startswith("l2_read_cg_")), "measured: tools/microbench ld.global.cg (L1 bypassed) over L2-resident sets of 16-96 MB, best size (profiles/r2_microbench.json)"),
            "smem": (d["smem_rec16_gbs"], "measured: tools/microbench scattered dense 16-byte records, one LDS.128 per lane -- the staged kernel's pattern (profiles/r2_microbench.json; conflict-free LDS.128 streams reach %.0f GB/s)" % d["smem_read_gbs"]),
            "gathers": d}


def roofline_block(w, bpt, per_gpu, kernel_ms, variant, workload):
    """(B - B_io) x traces/s against the measured peak of the level that holds the map (shared memory when
    the map is staged, else L2), and separately B_io x traces/s against the HBM copy peak."""
    rate = per_gpu / (kernel_ms * 1e-3)
    peaks = level_peaks()
    level = "smem" if variant == 1 else "l2"
    peak, src = peaks[level]
    achieved = bpt["b_map"] * rate / 1e9
    hbm_peak, hbm_src = measured_peak()
    traffic, ncu = None, None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp)).get(workload)
            if tj:
                traffic = tj["dram_bytes_per_segment"] * per_gpu
                # what bounds the kernel according to the committed ncu capture of this workload (not measured by this run):
                # the SM's L1 data pipe moves one wavefront per clock (a scattered 16-byte gather is one wavefront per distinct
                # line a warp touches, whatever level serves it) and the four schedulers issue one warp instruction per clock each
                ncu = {k: tj[k] for k in ("l1_data_pipe_pct_of_peak", "issue_slots_pct", "l2_throughput_pct", "lanes_per_instruction",
                                          "l1_data_pipe_wavefronts_per_segment", "warp_instructions_per_segment", "source") if k in tj}
                ncu["note"] = "from the committed ncu --set full capture of this workload's kernel (profiles/), per launch of 16 Mi segments"
        except Exception:
            pass
    m = bpt["means"]
    gathers = m["nodes"] + m["sides"] / 2.0 + m["solid"] / 4.0 + m["leaf_brushes"]
    gpeak = peaks["gathers"]["smem_rec16_per_s" if level == "smem" else "gather16_indep_2mb_per_s"]
    return {"bound": level, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "peak_source": src, "bytes_per_trace": bpt["b_total"], "bytes_per_trace_map": bpt["b_map"], "bytes_per_trace_io": bpt["b_io"],
            "visit_means": bpt["means"], "visit_means_source": bpt["source"], "kernel_ms": kernel_ms,
            "hbm_stream": {"achieved": bpt["b_io"] * rate / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": bpt["b_io"] * rate / 1e9 / hbm_peak, "peak_source": hbm_src},
            "ncu": ncu,
            "diagnostics": {"record_gathers_per_trace": gathers, "record_gathers_per_s": gathers * rate, "scattered_gather_rate_of_the_level_per_s": gpeak,
                            "note": "record gathers per second against the microbenchmark's fully scattered rate of the same level; lanes share top-of-tree records, so this is a diagnostic, not a ceiling"},
            "note": "achieved = algorithmic map bytes (the records the reference walk touches, BASELINE.md section 5) per second; traffic = DRAM bytes per launch from ncu (the 48 B/segment stream and nothing else)"}


def check_sample(w, seg_rows, res_rows):
    """CHECKER (oracle port, allowed here: bench.py's cpu_baseline / parity leg): gathered sample rows against oracle/restate.c"""
    import oracle
    exp, _, _ = oracle.Port(w["m"]).trace(np.ascontiguousarray(seg_rows), want_pos=False)
    bad = int((np.ascontiguousarray(res_rows).view(np.uint32).reshape(-1, 4) != exp.view(np.uint32).reshape(-1, 4)).any(axis=1).sum())
    return {"checked": int(len(seg_rows)), "mismatches": bad, "against": "oracle/restate.c (bit-exact on all four result words)"}


def run_workload(name, a, local, stream, rank, world, per_gpu, steps, want_e2e):
    """device-resident throughput of one workload on every rank (weak scaling: per_gpu segments per rank per step), hit-count
    all-reduce and sample gather through the library's NCCL communicator, parity of the gathered sample on rank 0."""
    import torch
    import torch.distributed as dist
    import bsp_b200
    from bsp_b200 import sharding

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if rank == 0:
        w = workload_setup(name)                      # rank 0 builds the map cache first, the others then just load it
    if world > 1:
        dist.barrier()
    if rank != 0:
        w = workload_setup(name)
    g = bsp_b200.BspGpu(w["path"], device=local)
    g.set_stream(stream.cuda_stream)
    for opt, val in ((2, a.threads), (6, a.refill), (7, a.max_steps), (1, a.variant), (4, a.top_bytes), (3, a.ctas), (8, a.block_segs), (5, a.chunk)):
        if val is not None:
            g.set_option(opt, val)
    libcomm = g if sharding.init_library_comm(g) else None

    first = rank * per_gpu
    res = {}
    with torch.cuda.stream(stream):
        segs = torch.empty((per_gpu, 8), dtype=torch.float32, device="cuda")
        out = torch.empty((per_gpu, 4), dtype=torch.int32, device="cuda")
        device_segments(g, w, segs, first, per_gpu)
        hits_dev = torch.zeros(1, dtype=torch.int64, device="cuda")

        def step():
            g.trace(segs, out=out, async_=True, stream=False)
            if libcomm is not None:
                sharding.reduce_hits(hits_dev, libcomm)       # NCCL inside libbspgpu.so: the hit-count all-reduce (8 bytes)
            else:
                g.hit_count_device(hits_dev)

        for _ in range(a.warmup):
            step()
        launches_per_step = g.info()["last_launches"]

        sampler = ClockSampler(local) if rank == 0 else None
        barrier()
        t_wall0 = time.time()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        ev[0].record(stream)
        for i in range(steps):
            kev[i][0].record(stream)
            g.trace(segs, out=out, async_=True, stream=False)
            kev[i][1].record(stream)
            if libcomm is not None:
                sharding.reduce_hits(hits_dev, libcomm)
            else:
                g.hit_count_device(hits_dev)
            ev[i + 1].record(stream)
        barrier()
        t_wall1 = time.time()
        if sampler:
            sampler.window(t_wall0, t_wall1)
        elapsed_ms = ev[0].elapsed_time(ev[-1])
        kernel_ms = [k0.elapsed_time(k1) for k0, k1 in kev]
        total_hits = int(hits_dev.item())

        # ---- correctness evidence at THIS world size: every rank's own flag count must add up to the all-reduced counter,
        #      and a strided sample of every rank's shard (segments + results, gathered over NCCL) is re-traced by the oracle port
        own = (out[:, 3] & 1).sum().to(torch.int64).reshape(1)
        if world > 1:
            dist.all_reduce(own)
        flag_hits = int(own.item())
        idx = sharding.sample_indices(per_gpu, a.parity_rows).to("cuda")
        g_res = sharding.gather_equal(out[idx].contiguous(), libcomm)
        g_seg = sharding.gather_equal(segs[idx].contiguous(), libcomm)
        stream.synchronize()

    elapsed_ms = sharding.max_over_ranks(elapsed_ms, "cuda")
    res["value"] = per_gpu * world * steps / (elapsed_ms * 1e-3)
    res["ms_per_step"] = elapsed_ms / steps
    res["hits"] = total_hits
    if flag_hits != total_hits:
        raise SystemExit(f"bench.py: {name}: all-reduced hit counter {total_hits} != sum of the ranks' hit flags {flag_hits}")
    if rank == 0:
        parity = None
        if not a.no_cpu or world > 1:
            parity = check_sample(w, g_seg.reshape(-1, 8).cpu().numpy(), g_res.reshape(-1, 4).cpu().numpy())
            parity["rows_per_rank"] = int(len(idx)); parity["ranks"] = world
            parity["gathered_with"] = "ncclSend/ncclRecv inside libbspgpu.so (bspgpu_dist_gather)" if world > 1 else "local"
            if parity["mismatches"]:
                raise SystemExit(f"bench.py: {name}: {parity['mismatches']} of {parity['checked']} sampled results differ from the oracle at world size {world}")
        res["parity_sample"] = parity
        res["clocks"] = sampler.stop() if sampler else None
        info = g.info()
        res["kernel"] = {"variant": info["variant"], "grid": info["grid_ctas"], "block": info["block_threads"], "smem": info["smem_bytes"],
                         "smem_stack_slots": info["smem_stack_slots"], "general_node_planes": info["num_general_planes"]}
        res["kernel_ms"] = float(np.mean(kernel_ms))
        res["launches_per_step"] = launches_per_step
        res["hit_count_reduction"] = "ncclAllReduce inside libbspgpu.so (bspgpu_dist_hit_count)" if world > 1 else "single GPU"
    res["w"] = w
    if want_e2e:
        res["e2e"] = e2e_leg(g, w, a, local, rank, world, segs, out, barrier)
    del segs, out
    g.close()
    torch.cuda.empty_cache()
    return res


def link_probe(n_bytes=256 << 20, reps=4, sync=None):
    """what this box's host link moves with pinned memory: H2D alone, D2H alone, both at once (per direction);
    sync = a barrier to call before every measurement (all ranks measure the same thing at the same time)"""
    import torch
    h_in = torch.empty(n_bytes, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n_bytes, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n_bytes, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n_bytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(h2d, d2h):
        torch.cuda.synchronize()
        if sync:
            sync()
        t0 = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        return reps * n_bytes / (time.perf_counter() - t0) / 1e9
    run(True, True)
    return {"h2d_alone_gbs": run(True, False), "d2h_alone_gbs": run(False, True), "both_each_gbs": run(True, True)}


def e2e_leg(g, w, a, local, rank, world, segs, out, barrier):
    """the same metric through the public C-ABI call with HOST buffers: every step copies that step's segments H2D and its
    results D2H inside the timed region.  Segments are handed over as the reference's own argument layout -- two packed
    glm::vec3 = 24 bytes (BSP::trace_seg, bsp.h:210; BSP::trace_segs in the C++ mirror).  Measured for both result layouts
    (16-byte TraceResult, 8-byte bspgpu_result8), with pinned and with pageable caller memory, plus small-batch latencies."""
    import ctypes as C
    import torch
    import bsp_b200
    from bsp_b200 import sharding
    e2e_n = min(len(segs), a.e2e_segments)
    lib = bsp_b200.load_library()
    bind_to_gpu_numa_node(local)
    h_in_ptr = lib.bspgpu_host_alloc(e2e_n * 24)
    h_out_ptr = lib.bspgpu_host_alloc(e2e_n * 16)
    if not h_in_ptr or not h_out_ptr:
        raise SystemExit("bench.py: pinned allocation failed")
    h_in = np.ctypeslib.as_array((C.c_float * (e2e_n * 6)).from_address(h_in_ptr)).reshape(e2e_n, 6)
    h_out = np.ctypeslib.as_array((C.c_uint32 * (e2e_n * 4)).from_address(h_out_ptr)).reshape(e2e_n, 4)
    seg_host = segs[:e2e_n].cpu().numpy()
    h_in[:, 0:3] = seg_host[:, 0:3]
    h_in[:, 3:6] = seg_host[:, 4:7]
    del seg_host
    h_res = h_out.view(bsp_b200.RESULT_DT).reshape(-1)
    h_res8 = h_out.reshape(-1)[: e2e_n * 2].view(bsp_b200.RESULT8_DT)
    e2e_steps = max(2, min(a.steps, 5))
    g.reset_stream()
    ref_rows = out[:e2e_n: max(1, e2e_n // 65536)].cpu().numpy().view(np.uint32)
    dev_hits_prefix = int((out[:e2e_n, 3] & 1).sum().item())

    def timed(fn, steps):
        for _ in range(2):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()                                                  # H2D + kernels + D2H, synchronous on return
        torch.cuda.synchronize()
        return e2e_n * world * steps / sharding.max_over_ranks(time.perf_counter() - t0, "cuda")

    layouts = {}
    layouts["result16"] = {"value": timed(lambda: g.trace(h_in, out=h_res, packed24=True), e2e_steps), "h2d_bytes_per_step": e2e_n * 24, "d2h_bytes_per_step": e2e_n * 16}
    if int((h_res["flags"] & 1).sum()) != dev_hits_prefix or not np.array_equal(h_out[:: max(1, e2e_n // 65536)], ref_rows):
        raise SystemExit("bench.py: host-path results differ from device-path results")
    layouts["compact8"] = {"value": timed(lambda: g.trace(h_in, out=h_res8, packed24=True, compact=True), e2e_steps), "h2d_bytes_per_step": e2e_n * 24, "d2h_bytes_per_step": e2e_n * 8}
    stride = max(1, e2e_n // 65536)
    if int((h_res8["bits"] & 1).sum()) != dev_hits_prefix or not np.array_equal(h_res8["t"][::stride].view(np.uint32), ref_rows[:, 0]) \
            or not np.array_equal(h_res8["bits"][::stride], ref_rows[:, 3] | ((ref_rows[:, 1] + 1) << 2)):
        raise SystemExit("bench.py: compact host-path results differ from device-path results")
    h_t = h_out.reshape(-1)[:e2e_n].view(np.float32)
    layouts["compact4"] = {"value": timed(lambda: g.trace(h_in, out=h_t, packed24=True, compact=4), e2e_steps), "h2d_bytes_per_step": e2e_n * 24, "d2h_bytes_per_step": e2e_n * 4}
    # {hit, t} back from the 4-byte form: hit <=> value != T_MISS, t = value on a hit and 1 on a miss (the reference's own outputs)
    hit4 = h_t != np.float32(bsp_b200.T_MISS)
    t4 = np.where(hit4, h_t, np.float32(1.0)).astype(np.float32)
    if int(hit4.sum()) != dev_hits_prefix or not np.array_equal(t4[::stride].view(np.uint32), ref_rows[:, 0]) or not np.array_equal(hit4[::stride], (ref_rows[:, 3] & 1) != 0):
        raise SystemExit("bench.py: 4-byte host-path results differ from device-path results")
    best = max(layouts, key=lambda k: layouts[k]["value"])
    e2e = {"value": layouts[best]["value"], "unit": "traces/s", "h2d_bytes_per_step": layouts[best]["h2d_bytes_per_step"],
           "d2h_bytes_per_step": layouts[best]["d2h_bytes_per_step"], "layout": best, "layouts": layouts,
           "layout_note": "24 B packed vec3 pairs in (the reference's trace_seg arguments); out: result16 = 16-byte TraceResult {t, plane, brush, flags}, "
                          "compact8 = 8-byte bspgpu_result8 {t, flags | (plane+1) << 2} (BSPGPU_OUT_COMPACT8: everything the reference defines, hit and t, plus the plane extension); "
                          "compact4 = one float, t on a hit and 2.0 on a miss (BSPGPU_OUT_COMPACT4: exactly the reference's outputs hit and t -- a hit has t <= 1 -- and pos is start + t * (end - start) from the caller's own segment); "
                          "pinned host memory, chunked 3-stream pipeline",
           "segments_per_gpu_per_step": e2e_n, "steps": e2e_steps}
    if rank == 0:
        # pageable caller memory (what an existing caller of the reference has): bounced through the library's pinned ring by host
        # threads (the default), page-locked for the call, or handed to the driver as it is
        pg_n = min(e2e_n, 16 * 1024 * 1024)
        p_in = np.array(h_in[:pg_n]); p_out = np.zeros(pg_n, bsp_b200.RESULT_DT); p_out8 = np.zeros(pg_n, bsp_b200.RESULT8_DT)
        pageable = {}
        for opt, key, o in ((2, "pinned_ring_and_host_threads", p_out), (2, "pinned_ring_and_host_threads_8B_out", p_out8),
                            (1, "page_locked_for_the_call", p_out), (0, "driver_staged", p_out)):
            g.set_option(17, opt)
            g.trace(p_in, out=o, packed24=True, compact=o is p_out8)
            t0 = time.perf_counter()
            for _ in range(3):
                g.trace(p_in, out=o, packed24=True, compact=o is p_out8)
            pageable[key] = pg_n * 3 / (time.perf_counter() - t0)
        g.set_option(17, 2)
        if not np.array_equal(p_out.view(np.uint32).reshape(-1, 4)[:: max(1, pg_n // 65536)], out[:pg_n: max(1, pg_n // 65536)].cpu().numpy().view(np.uint32)):
            raise SystemExit("bench.py: pageable host-path results differ from device-path results")
        e2e["e2e_pageable"] = {"value": pageable["pinned_ring_and_host_threads"], "unit": "traces/s", "segments": pg_n, "variants": pageable,
                               "note": "plain malloc'ed caller arrays, 24 B in / 16 B out (8 B where said); default = BSPGPU_OPT_PAGEABLE 2: host threads copy "
                                       "chunk c+1 into and chunk c-2 out of the library's pinned ring while the GPU works on the chunks in between"}
        # small batches: BSP::trace_seg is a batch of one
        lat = {}
        for n, tiny in ((1, 1), (1, 0), (16, 1), (16, 0), (1024, 1), (65536, 1)):
            g.set_option(18, tiny)                    # BSPGPU_OPT_TINY_CALLS: up to 16 host segments = one kernel over the pinned buffer
            s_in = np.array(h_in[:n]); s_out = np.zeros(n, bsp_b200.RESULT_DT)
            for _ in range(20):
                g.trace(s_in, out=s_out, packed24=True)
            reps = 200 if n <= 1024 else 50
            t0 = time.perf_counter()
            for _ in range(reps):
                g.trace(s_in, out=s_out, packed24=True)
            lat[str(n) if tiny else f"{n}_general_small_path"] = {"us_per_call": 1e6 * (time.perf_counter() - t0) / reps, "grid_ctas": g.info()["grid_ctas"]}
        g.set_option(18, 1)
        lat["note"] = "host arrays through bspgpu_trace from Python (ctypes call included); a batch of one is BSP::trace_seg, the reference's own call shape"
        e2e["latency"] = lat
        e2e["link_ceiling"] = link_probe()
        lc = e2e["link_ceiling"]
        e2e["link_ceiling"]["segments_per_s_24in_8out"] = min(lc["both_each_gbs"] * 1e9 / 24, lc["both_each_gbs"] * 1e9 / 8)
        e2e["link_ceiling"]["segments_per_s_24in_16out"] = min(lc["both_each_gbs"] * 1e9 / 24, lc["both_each_gbs"] * 1e9 / 16)
        e2e["link_ceiling"]["note"] = "pinned copies on this rank's GPU alone (256 MiB x 4); ceiling = the slower direction at the both-ways rate; other ranks idle while rank 0 probes"
    if world > 1:
        # the box's host side is shared: the same probe with every rank copying at once (sums over the ranks)
        import torch
        import torch.distributed as dist
        barrier(); one = link_probe(reps=3, sync=barrier)
        t = torch.tensor([one["h2d_alone_gbs"], one["d2h_alone_gbs"], one["both_each_gbs"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        if rank == 0:
            lay = e2e["layouts"][e2e["layout"]]
            e2e["link_ceiling_all_ranks"] = {
                "h2d_alone_gbs": t[0].item(), "d2h_alone_gbs": t[1].item(), "both_each_gbs": t[2].item(),
                "achieved_h2d_gbs": e2e["value"] * lay["h2d_bytes_per_step"] / e2e_n / 1e9, "achieved_d2h_gbs": e2e["value"] * lay["d2h_bytes_per_step"] / e2e_n / 1e9,
                "note": "all %d ranks copying at once from pinned memory (sum over the ranks, GB/s): input direction alone, output direction alone, "
                        "both at once (per direction); achieved = what the timed e2e steps moved in each direction" % world}
    lib.bspgpu_host_free(h_in_ptr)
    lib.bspgpu_host_free(h_out_ptr)
    return e2e


def run_ours(a):
    import torch
    import torch.distributed as dist
    import bsp_b200

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        log(f"warning: --gpus {a.gpus} but WORLD_SIZE={world}; using WORLD_SIZE")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the trace has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        bsp_b200.build()                                  # the native library, once; the other ranks load it after the barrier
    if world > 1:
        dist.barrier()
    stream = torch.cuda.Stream()

    main_res = run_workload(a.workload, a, local, stream, rank, world, a.segments_per_gpu, a.steps, want_e2e=True)
    w = main_res.pop("w")
    line = None
    if rank == 0:
        cpu_line, counters = None, None
        if world == 1 and not a.no_cpu:
            extras, cpu_line = cpu_reference(w)
            counters = extras.get("counters")
        bpt = algorithmic_bytes(w, counters)
        per_gpu = a.segments_per_gpu
        line = {
            "metric": "bsp_segment_traces_per_sec", "value": main_res["value"], "unit": "traces/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": main_res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": a.workload, "map": w["map"], "map_counts": w["m"].counts(), "segments_per_gpu_per_step": per_gpu,
                       "segment_kind": w["kind"], "seed": w["seed"], "parallelism": f"segments sharded x{world}, map replicated",
                       "l2": f"inputs larger than L2 ({per_gpu * 48 / 1e9:.1f} GB streamed per step per GPU)",
                       "kernel": main_res["kernel"], "collectives": main_res["hit_count_reduction"]},
            "hits": main_res["hits"], "parity_sample": main_res["parity_sample"],
            "roofline": roofline_block(w, bpt, per_gpu, main_res["kernel_ms"], main_res["kernel"]["variant"], a.workload),
            "e2e": main_res["e2e"],
            "gpu_launches": main_res["launches_per_step"] * a.steps,
            "clocks": main_res["clocks"],
            "provenance": "every number in this line was measured by this run (driver-run when the driver launched it)",
        }
        if cpu_line is not None:
            line["cpu_baseline"] = cpu_line

    # ---- the other configurations of BASELINE.json, device-resident, same timing rules, each with its own parity sample:
    #      C2 rooms8 (64 Mi uniform segments, map staged in shared memory), C4 camera grids on the city (row-major, and in a
    #      shuffled order for the divergence contrast), C5 mega200k short moves.
    #      At N > 1 only C5 is repeated (BASELINE.json quotes it on 8 GPUs); the others are single-GPU configurations.
    also = [x for x in a.also.split(",") if x and x != a.workload]
    if world > 1:
        also = [x for x in also if x == "mega200k-short"]
    also_out = {}
    for name in also:
        try:
            n2 = 64 * 1024 * 1024 if name == "rooms8-uniform" else a.segments_per_gpu
            r2 = run_workload(name, a, local, stream, rank, world, n2, max(3, min(a.steps, 10)), want_e2e=False)
            w2 = r2.pop("w")
            if rank == 0:
                bpt2 = algorithmic_bytes(w2)              # committed visit counts (profiles/bytes_per_trace.json)
                also_out[name] = {"value": r2["value"], "unit": "traces/s", "n_gpus": world, "ms_per_step": r2["ms_per_step"], "segments_per_gpu_per_step": n2,
                                  "hits": r2["hits"], "parity_sample": r2["parity_sample"], "kernel": r2["kernel"], "clocks": r2["clocks"],
                                  "roofline": roofline_block(w2, bpt2, n2, r2["kernel_ms"], r2["kernel"]["variant"], name)}
        except SystemExit:
            raise
        except Exception as e:  # pragma: no cover
            if rank == 0:
                also_out[name] = {"error": repr(e)}
    if rank == 0:
        if also_out:
            line["also"] = also_out
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="city10k-long", choices=list(WORKLOADS))
    ap.add_argument("--segments-per-gpu", type=int, default=128 * 1024 * 1024)
    ap.add_argument("--e2e-segments", type=int, default=32 * 1024 * 1024)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--also", default="rooms8-uniform,city10k-camera,city10k-camera-shuffled,mega200k-short",
                    help="other BASELINE.json configurations reported under 'also' (comma list, '' to skip); at N>1 only mega200k-short runs")
    ap.add_argument("--parity-rows", type=int, default=4096, help="rows of every rank's shard re-traced by the oracle port")
    ap.add_argument("--variant", type=int, default=None)
    ap.add_argument("--threads", type=int, default=None)
    ap.add_argument("--refill", type=int, default=None)
    ap.add_argument("--max-steps", type=int, default=None)
    ap.add_argument("--top-bytes", type=int, default=None)
    ap.add_argument("--ctas", type=int, default=None)
    ap.add_argument("--block-segs", type=int, default=None)
    ap.add_argument("--chunk", type=int, default=None, help="host path: segments per pipelined chunk")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
