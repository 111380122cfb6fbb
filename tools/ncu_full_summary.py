#!/usr/bin/env python
"""One markdown row per kernel of an `ncu --set full` report: time, DRAM traffic and % of peak, L2 hit, IPC, tensor / XU pipe,
occupancy, registers, top warp-stall reasons.   ncu_full_summary.py rep.ncu-rep [name ...]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
names = sys.argv[2:]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
SCALE = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3}


def scaled(r, key):
    """value of column `key` in MB (bytes) or us (time), whatever unit ncu chose for the column"""
    i = hdr.index(key)
    try:
        return float(r[i].replace(",", "")) * SCALE.get(units[i], 1.0)
    except Exception:
        return float("nan")



def col(r, key, exact=True):
    for i, h in enumerate(hdr):
        if (h == key) if exact else (key in h):
            return r[i]
    return ""


def f(x, d=1):
    try:
        return f"{float(x.replace(',', '')):.{d}f}"
    except Exception:
        return "-"


stall_cols = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
print("| launch | kernel | grid | time us | DRAM rd MB | DRAM wr MB | DRAM % of peak | L2 hit % | IPC | tensor pipe % | XU pipe % | warps/sched | regs | top stalls (warps per issue) |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for i, r in enumerate(data):
    k = col(r, "Kernel Name").replace("void ", "").replace("cf::", "").split("(")[0]
    stalls = sorted(((float(r[hdr.index(h)] or 0), h.split("stalled_")[1].split("_per")[0]) for h in stall_cols), reverse=True)[:3]
    st = ", ".join(f"{n} {v:.1f}" for v, n in stalls)
    nm = names[i] if i < len(names) else ""
    print(f"| {nm} | `{k}` | {col(r, 'Grid Size')} | {scaled(r, 'gpu__time_duration.sum'):.1f} | {scaled(r, 'dram__bytes_read.sum'):.1f} | "
          f"{scaled(r, 'dram__bytes_write.sum'):.1f} | {f(col(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'))} | "
          f"{f(col(r, 'lts__t_sector_hit_rate.pct'))} | {f(col(r, 'sm__inst_executed.avg.per_cycle_elapsed'), 2)} | "
          f"{f(col(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'))} | "
          f"{f(col(r, 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'))} | "
          f"{f(col(r, 'smsp__warps_active.avg.per_cycle_active'))} | {col(r, 'launch__registers_per_thread')} | {st} |")
