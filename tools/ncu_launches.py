#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-launch table and
per-kernel-family shares (cold-cache, serialised times: compare SHARES, not absolutes)."""
import csv
import re
import sys
from collections import OrderedDict


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    return list(csv.DictReader(lines))


def family(name):
    m = re.match(r"(?:void )?(?:cf::)?(\w+)", name)
    return m.group(1) if m else name


def main():
    rows = load(sys.argv[1])
    tot = sum(float(r["Metric Value"].replace(",", "")) for r in rows)
    unit = rows[0]["Metric Unit"]
    print(f"| # | kernel | grid | block | {unit} | share |")
    print("|---|---|---|---|---|---|")
    fam = OrderedDict()
    for r in rows:
        v = float(r["Metric Value"].replace(",", ""))
        name = r["Kernel Name"]
        short = re.sub(r"\(.*", "", name.replace("void ", "").replace("cf::", ""))
        print(f"| {r['ID']} | `{short}` | {r['Grid Size']} | {r['Block Size']} | {v:.0f} | {100 * v / tot:.1f}% |")
        f = fam.setdefault(family(name), [0, 0.0])
        f[0] += 1
        f[1] += v
    print()
    print(f"| kernel family | launches | {unit} | share |")
    print("|---|---|---|---|")
    for k, (n, v) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {v:.0f} | {100 * v / tot:.1f}% |")
    print(f"| total | {len(rows)} | {tot:.0f} | 100% |")


if __name__ == "__main__":
    main()
