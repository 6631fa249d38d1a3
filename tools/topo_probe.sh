#!/bin/bash
# one-off: what the 8-GPU box looks like from the host side (PCIe / NUMA), for reading the multi-GPU e2e numbers
nvidia-smi topo -m 2>&1 | head -30
echo "--- numa"; ls /sys/devices/system/node/ | head; for n in /sys/devices/system/node/node*; do echo $n $(cat $n/cpulist) $(grep MemTotal $n/meminfo); done
echo "--- gpu numa"; for b in $(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader | tr 'A-Z' 'a-z' | sed 's/^0000//'); do echo $b $(cat /sys/bus/pci/devices/$b/numa_node 2>/dev/null) $(cat /sys/bus/pci/devices/$b/current_link_speed 2>/dev/null) x$(cat /sys/bus/pci/devices/$b/current_link_width 2>/dev/null); done
nproc; lscpu | grep -i "model name\|socket\|numa" | head
python - <<'PY'
import torch, time
# all 8 GPUs copying H2D at once from one process: aggregate host->device bandwidth of the box
n = 1 << 28
hs = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(8)]
ds = [torch.empty(n, dtype=torch.uint8, device=f"cuda:{i}") for i in range(8)]
ss = [torch.cuda.Stream(device=i) for i in range(8)]
def run(k, reps=8):
    for i in range(k): torch.cuda.synchronize(i)
    t0 = time.perf_counter()
    for _ in range(reps):
        for i in range(k):
            with torch.cuda.stream(ss[i]): ds[i].copy_(hs[i], non_blocking=True)
    for i in range(k): torch.cuda.synchronize(i)
    return reps * k * n / (time.perf_counter() - t0) / 1e9
run(8, 1)
for k in (1, 2, 4, 8): print("H2D aggregate GB/s with", k, "GPUs:", round(run(k), 1))
PY
