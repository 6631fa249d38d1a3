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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # rank 0 builds the native library and the map cache first, the others then just load them
    if rank == 0:
        bsp_b200.build()
        w = workload_setup(a.workload)
    if world > 1:
        dist.barrier()
    if rank != 0:
        w = workload_setup(a.workload)

    g = bsp_b200.BspGpu(w["path"], device=local)
    stream = torch.cuda.Stream()
    g.set_stream(stream.cuda_stream)
    for opt, val in ((2, a.threads), (6, a.refill), (7, a.max_steps), (1, a.variant), (4, a.top_bytes), (3, a.ctas), (8, a.block_segs), (5, a.chunk)):
        if val is not None:
            g.set_option(opt, val)

    per_gpu = a.segments_per_gpu
    first = rank * per_gpu
    with torch.cuda.stream(stream):
        segs = torch.empty((per_gpu, 8), dtype=torch.float32, device="cuda")
        out = torch.empty((per_gpu, 4), dtype=torch.int32, device="cuda")
        device_segments(g, w, segs, first, per_gpu)
        hits_dev = torch.zeros(1, dtype=torch.int64, device="cuda")
        sample_idx = torch.arange(0, per_gpu, max(1, per_gpu // 4096), device="cuda")[:4096]

        def step():
            g.trace(segs, out=out, async_=True)
            g.hit_count_device(hits_dev)
            if world > 1:
                dist.all_reduce(hits_dev)                 # NCCL: the hit-count reduction (8 bytes)

        for _ in range(a.warmup):
            step()
        launches_per_step = g.info()["last_launches"]

        sampler = ClockSampler(local) if rank == 0 else None
        barrier()
        t_wall0 = time.time()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
        ev[0].record(stream)
        for i in range(a.steps):
            kev[i][0].record(stream)
            g.trace(segs, out=out, async_=True)
            kev[i][1].record(stream)
            g.hit_count_device(hits_dev)
            if world > 1:
                dist.all_reduce(hits_dev)
            ev[i + 1].record(stream)
        barrier()
        t_wall1 = time.time()
        if sampler:
            sampler.window(t_wall0, t_wall1)
        elapsed_ms = ev[0].elapsed_time(ev[-1])
        kernel_ms = [k0.elapsed_time(k1) for k0, k1 in kev]
        total_hits = int(hits_dev.item())

        # small result gather over NCCL (sampled; parity on the gathered rows is checked by rank 0 below)
        sample_res = out[sample_idx].contiguous()
        sample_seg = segs[sample_idx].contiguous()
        if world > 1:
            gathered = [torch.empty_like(sample_res) for _ in range(world)] if rank == 0 else None
            dist.gather(sample_res, gathered, dst=0)

    t = torch.tensor([elapsed_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    total_segments = per_gpu * world * a.steps
    value = total_segments / (elapsed_ms * 1e-3)

    # ---------------- e2e: host buffers through the same public call.  Segments are handed over as the
    # reference's own argument layout -- two packed glm::vec3 = 24 bytes (BSP::trace_seg, bsp.h:210;
    # BSP::trace_segs in the C++ mirror) -- results come back as 16-byte TraceResults.
    e2e_n = min(per_gpu, a.e2e_segments)
    lib = bsp_b200.load_library()
    bind_to_gpu_numa_node(local)
    h_in_ptr = lib.bspgpu_host_alloc(e2e_n * 24)
    h_out_ptr = lib.bspgpu_host_alloc(e2e_n * 16)
    if not h_in_ptr or not h_out_ptr:
        raise SystemExit("bench.py: pinned allocation failed")
    import ctypes as C
    h_in = np.ctypeslib.as_array((C.c_float * (e2e_n * 6)).from_address(h_in_ptr)).reshape(e2e_n, 6)
    h_out = np.ctypeslib.as_array((C.c_uint32 * (e2e_n * 4)).from_address(h_out_ptr)).reshape(e2e_n, 4)
    seg_host = segs[:e2e_n].cpu().numpy()
    h_in[:, 0:3] = seg_host[:, 0:3]
    h_in[:, 3:6] = seg_host[:, 4:7]
    del seg_host
    h_res = h_out.view(bsp_b200.RESULT_DT).reshape(-1)
    e2e_steps = max(2, min(a.steps, 5))
    g.reset_stream()
    for _ in range(2):
        g.trace(h_in, out=h_res, packed24=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        g.trace(h_in, out=h_res, packed24=True)            # H2D + kernels + D2H, synchronous on return
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = e2e_n * world * e2e_steps / float(te.item())
    e2e_hits = int((h_res["flags"] & 1).sum())
    dev_hits_prefix = int((out[:e2e_n, 3] & 1).sum().item())
    if e2e_hits != dev_hits_prefix:
        raise SystemExit(f"bench.py: host-path results differ from device-path results ({e2e_hits} vs {dev_hits_prefix} hits)")
    if not np.array_equal(h_out[:: max(1, e2e_n // 65536)], out[:e2e_n: max(1, e2e_n // 65536)].cpu().numpy().view(np.uint32)):
        raise SystemExit("bench.py: host-path results differ from device-path results")

    if rank == 0:
        clocks = sampler.stop() if sampler else None
        cpu_line, counters, parity = None, None, None
        if world == 1 and not a.no_cpu:
            extras, cpu_line = cpu_reference(w, parity=(sample_seg.cpu().numpy(), sample_res.cpu().numpy()))
            counters, parity = extras.get("counters"), extras.get("parity")
        bpt = algorithmic_bytes(w, counters)
        peak, peak_src = measured_peak()
        mean_kernel_ms = float(np.mean(kernel_ms))
        achieved = bpt["b_total"] * per_gpu / (mean_kernel_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp)).get(a.workload)
                if tj:
                    traffic = tj["dram_bytes_per_segment"] * per_gpu
            except Exception:
                pass
        info = g.info()
        line = {
            "metric": "bsp_segment_traces_per_sec", "value": value, "unit": "traces/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": elapsed_ms / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": a.workload, "map": w["map"], "map_counts": w["m"].counts(), "segments_per_gpu_per_step": per_gpu,
                       "segment_kind": w["kind"], "seed": w["seed"], "parallelism": f"segments sharded x{world}, map replicated",
                       "l2": f"inputs larger than L2 ({per_gpu * 48 / 1e9:.1f} GB streamed per step per GPU)",
                       "kernel": {"variant": info["variant"], "grid": info["grid_ctas"], "block": info["block_threads"], "smem": info["smem_bytes"]}},
            "hits": total_hits, "parity_sample": parity,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "bytes_per_trace": bpt["b_total"], "bytes_per_trace_map": bpt["b_map"],
                         "bytes_per_trace_io": bpt["b_io"], "visit_means": bpt["means"], "visit_means_source": bpt["source"], "kernel_ms": mean_kernel_ms,
                         "hbm_stream_gbs": bpt["b_io"] * per_gpu / (mean_kernel_ms * 1e-3) / 1e9,
                         "map_level": map_level_roofline(bpt, per_gpu / (mean_kernel_ms * 1e-3), info["variant"]),
                         "note": "achieved = algorithmic bytes (map records the reference walk touches + 48 B segment/result I/O) per second; "
                                 "map bytes are served from shared memory / L1 / L2, only the I/O stream reaches HBM"},
            "e2e": {"value": e2e_value, "unit": "traces/s", "h2d_bytes_per_step": e2e_n * 24, "d2h_bytes_per_step": e2e_n * 16,
                    "layout": "24 B packed vec3 pairs in (the reference's trace_seg arguments), 16 B TraceResult out, pinned host memory, chunked 3-stream pipeline",
                    "segments_per_gpu_per_step": e2e_n, "steps": e2e_steps},
            "gpu_launches": launches_per_step * a.steps,
            "clocks": clocks,
        }
        if cpu_line is not None:
            line["cpu_baseline"] = cpu_line
        if world == 1 and a.also and a.also != a.workload:
            # the other single-GPU configuration of BASELINE.json (configs[1]: rooms8, 64 Mi uniform segments, map staged
            # in shared memory), device-resident, same timing rules -- reported beside the headline, not instead of it
            try:
                del segs, out
                torch.cuda.empty_cache()
                line["also"] = {a.also: secondary_workload(a.also, a, local, stream)}
            except Exception as e:  # pragma: no cover
                line["also"] = {a.also: {"error": str(e)}}
        print(json.dumps(line), flush=True)

    lib.bspgpu_host_free(h_in_ptr)
    lib.bspgpu_host_free(h_out_ptr)
    g.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def map_level_roofline(bpt, traces_per_s, variant):
    """the level that actually holds the map: record gathers per second against the measured scattered-gather
    rate of that level (profiles/r1_microbench.json, tools/microbench) -- the binding resource of this kernel."""
    try:
        peaks = json.load(open(os.path.join(ROOT, "profiles", "r1_microbench.json")))
        m = bpt["means"]
        # node records + side-plane PAIRS (two sides per 32-byte gather; a brush's executed sides round up to a pair,
        # +1/4 gather per brush on average) + brush references
        gathers = m["nodes"] + m["sides"] / 2.0 + m["solid"] / 4.0 + m["leaf_brushes"]
        in_smem = variant == 1
        peak = peaks["peaks_used_by_bench"]["smem_chase_gathers_per_s" if in_smem else "global_gathers_per_s"]
        return {"level": "shared memory" if in_smem else "L1/L2 read-only path", "record_gathers_per_trace": gathers,
                "achieved_gathers_per_s": gathers * traces_per_s, "peak_gathers_per_s": peak, "frac": gathers * traces_per_s / peak,
                "map_bytes_gbs": bpt["b_map"] * traces_per_s / 1e9,
                "level_bandwidth_gbs": peaks["smem_read_gbs" if in_smem else "l2_read_gbs"],
                "note": "peak = fully scattered 32-byte gathers (every lane a different record); the trace's top-of-tree gathers are "
                        "shared between lanes, so frac can exceed 1"}
    except Exception as e:  # pragma: no cover
        return {"error": str(e)}


def secondary_workload(name, a, local, stream):
    import torch
    import bsp_b200
    w2 = workload_setup(name)
    n2 = 64 * 1024 * 1024
    g2 = bsp_b200.BspGpu(w2["path"], device=local)
    g2.set_stream(stream.cuda_stream)
    with torch.cuda.stream(stream):
        segs2 = torch.empty((n2, 8), dtype=torch.float32, device="cuda")
        out2 = torch.empty((n2, 4), dtype=torch.int32, device="cuda")
        device_segments(g2, w2, segs2, 0, n2)
        for _ in range(3):
            g2.trace(segs2, out=out2, async_=True)
        torch.cuda.synchronize()
        steps = max(3, min(a.steps, 10))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            g2.trace(segs2, out=out2, async_=True)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
    bpt = algorithmic_bytes(w2)          # committed visit counts
    peak, _ = measured_peak()
    info = g2.info()
    res = {"value": n2 / (ms * 1e-3), "unit": "traces/s", "ms_per_step": ms, "steps": steps, "segments_per_step": n2,
           "hits": g2.hit_count(), "bytes_per_trace": bpt["b_total"],
           "roofline_frac": bpt["b_total"] * n2 / (ms * 1e-3) / 1e9 / peak,
           "kernel": {"variant": info["variant"], "grid": info["grid_ctas"], "block": info["block_threads"], "smem": info["smem_bytes"]}}
    g2.close()
    return res


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
    ap.add_argument("--also", default="rooms8-uniform", help="second workload reported under 'also' at N=1 ('' to skip)")
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
