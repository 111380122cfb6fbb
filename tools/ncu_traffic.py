#!/usr/bin/env python
"""Per-kernel time and DRAM traffic from `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`:
prints a markdown table and (optionally) writes the per-class DRAM bytes of one step as JSON for bench.py's `traffic`."""
import csv, json, re, sys
from collections import OrderedDict

SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6}

def cls(n):
    if "k_mbf" in n: return "fused_blocks"
    if "k_pw" in n: return "pointwise_gemm"
    if "k_dw" in n: return "depthwise"
    if "k_stem" in n: return "stem"
    if "k_heads" in n: return "heads"
    return "decode_topk"

def main():
    rows = list(csv.DictReader([l for l in open(sys.argv[1]) if not l.startswith("==")]))
    k = OrderedDict()
    for r in rows:
        v = float(r["Metric Value"].replace(",", "")) * SCALE.get(r["Metric Unit"], 1)
        k.setdefault((int(r["ID"]), r["Kernel Name"], r["Grid Size"]), {})[r["Metric Name"]] = v
    # a capture window longer than one step: keep the first COMPLETE step (first stem launch .. the top-k launch behind it)
    ids = list(k.keys())
    first = next((j for j, kk in enumerate(ids) if "k_stem" in kk[1]), 0)
    last = next((j for j in range(first, len(ids)) if "k_topk" in ids[j][1]), len(ids) - 1)
    k = OrderedDict((kk, k[kk]) for kk in ids[first:last + 1])
    agg = OrderedDict()
    print("| # | kernel | grid | time us | DRAM read MB | DRAM write MB | DRAM GB/s |")
    print("|---|---|---|---|---|---|---|")
    for (i, n, g), m in k.items():
        short = re.sub(r"\(.*", "", n.replace("void ", "").replace("cf::", ""))
        t, rd, wr = m["gpu__time_duration.sum"], m["dram__bytes_read.sum"], m["dram__bytes_write.sum"]
        print(f"| {i} | `{short}` | {g} | {t / 1e3:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {(rd + wr) / t:.0f} |")
        a = agg.setdefault(cls(n), {"ns": 0, "rd": 0, "wr": 0, "n": 0})
        a["ns"] += t; a["rd"] += rd; a["wr"] += wr; a["n"] += 1
    tot = sum(a["ns"] for a in agg.values())
    print("\n| class | launches | time us | share | DRAM MB |\n|---|---|---|---|---|")
    for c, a in agg.items():
        print(f"| {c} | {a['n']} | {a['ns'] / 1e3:.0f} | {100 * a['ns'] / tot:.1f}% | {(a['rd'] + a['wr']) / 1e6:.0f} |")
    print(f"| total | {sum(a['n'] for a in agg.values())} | {tot / 1e3:.0f} | 100% | {sum(a['rd'] + a['wr'] for a in agg.values()) / 1e6:.0f} |")
    if len(sys.argv) > 2:
        json.dump({"source": sys.argv[1], "dram_bytes_per_step": {c: a["rd"] + a["wr"] for c, a in agg.items()},
                   "launches": {c: a["n"] for c, a in agg.items()}}, open(sys.argv[2], "w"), indent=1)

if __name__ == "__main__":
    main()
