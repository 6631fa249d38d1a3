"""Summarise an .ncu-rep (read here, without a GPU): headline counters + where the warp
instructions go, by SIMT utilisation bucket and by code region."""
import collections
import csv
import subprocess
import sys


def main(rep, listing=None, thresh=0.3):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, vals = rows[0], rows[2]
    keys = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
            "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"]
    print("kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
    for k in keys:
        if k in hdr:
            print(f"  {k:75s} {vals[hdr.index(k)]} {rows[1][hdr.index(k)]}")
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warp") and h.endswith("_per_issue_active.ratio") or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
            try:
                if float(vals[i]) >= 0.15:
                    print(f"  {h:75s} {vals[i]}")
            except ValueError:
                pass
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr = rows[1]; data = rows[2:]
    ia, isrc, iinst, ithr, isamp = (hdr.index(x) for x in ("Address", "Source", "Instructions Executed", "Thread Instructions Executed", "# Samples"))
    tot = sum(int(r[iinst]) for r in data); tott = sum(int(r[ithr]) for r in data); tots = max(1, sum(int(r[isamp]) for r in data))
    print(f"  warp-instructions {tot}  thread-instructions {tott}  avg active threads {tott / tot:.2f}")
    base = int(data[0][ia], 16)
    buckets = collections.OrderedDict((k, 0.0) for k in ("<4", "4-8", "8-16", "16-28", ">=28"))
    lines = []
    for r in data:
        n = int(r[iinst])
        if not n:
            continue
        thr = int(r[ithr]) / n
        pct = 100.0 * n / tot
        b = "<4" if thr < 4 else "4-8" if thr < 8 else "8-16" if thr < 16 else "16-28" if thr < 28 else ">=28"
        buckets[b] += pct
        lines.append(f"{int(r[ia], 16) - base:5x} {pct:5.2f}% thr {thr:5.1f} smp {100 * int(r[isamp]) / tots:5.2f}%  {r[isrc].strip()[:70]}")
    print("  share of warp instructions by active threads:", {k: round(v, 1) for k, v in buckets.items()})
    if listing:
        open(listing, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
