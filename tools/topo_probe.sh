#!/bin/bash
# one-off: what the 8-GPU box looks like from the host side (PCIe / NUMA), for reading the multi-GPU e2e numbers
nvidia-smi topo -m 2>&1 | head -30
echo "--- numa"; ls /sys/devices/system/node/ | head; for n in /sys/devices/system/node/node*; do echo $n $(cat $n/cpulist) $(grep MemTotal $n/meminfo); done
echo "--- gpu numa"; for b in $(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader | tr 'A-Z' 'a-z' | sed 's/^0000//'); do echo $b $(cat /sys/bus/pci/devices/$b/numa_node 2>/dev/null) $(cat /sys/bus/pci/devices/$b/current_link_speed 2>/dev/null) x$(cat /sys/bus/pci/devices/$b/current_link_width 2>/dev/null); done
nproc; lscpu | grep -i "model name\|socket\|numa" | head
python - <<'PY'
import json, torch, time
# the box's aggregate host<->device bandwidth with 1 / 2 / 4 / 8 GPUs copying at once from one process (pinned memory):
# H2D alone, D2H alone, and both directions together -- the ceiling of the host-buffer (e2e) path at every N
G = torch.cuda.device_count()
n = 1 << 28
hi = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(G)]
ho = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(G)]
di = [torch.empty(n, dtype=torch.uint8, device=f"cuda:{i}") for i in range(G)]
do = [torch.empty(n, dtype=torch.uint8, device=f"cuda:{i}") for i in range(G)]
s1 = [torch.cuda.Stream(device=i) for i in range(G)]
s2 = [torch.cuda.Stream(device=i) for i in range(G)]
def run(k, h2d, d2h, reps=6):
    for i in range(k): torch.cuda.synchronize(i)
    t0 = time.perf_counter()
    for _ in range(reps):
        for i in range(k):
            if h2d:
                with torch.cuda.stream(s1[i]): di[i].copy_(hi[i], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2[i]): ho[i].copy_(do[i], non_blocking=True)
    for i in range(k): torch.cuda.synchronize(i)
    return reps * k * n / (time.perf_counter() - t0) / 1e9
run(G, True, True, 1)
out = {}
for k in (1, 2, 4, 8):
    if k > G: break
    both = run(k, True, True)
    out[str(k)] = {"h2d_alone_gbs": round(run(k, True, False), 1), "d2h_alone_gbs": round(run(k, False, True), 1), "both_each_direction_gbs": round(both, 1),
                   "segments_per_s_24in_16out": round(min(both / 24, both / 16) * 1e9), "segments_per_s_24in_8out": round(both / 24 * 1e9)}
print("LINK_CEILING", json.dumps(out))
PY
