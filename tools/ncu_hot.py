#!/usr/bin/env python
"""Hottest SASS instructions (warp-stall samples) of one kernel of an .ncu-rep: `ncu_hot.py rep.ncu-rep <launch index> [top]`."""
import csv
import subprocess
import sys

rep, idx = sys.argv[1], int(sys.argv[2])
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(idx), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
kname = rows[0][1] if rows and len(rows[0]) > 1 else "?"
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[h], [r for r in rows[h + 1:] if len(r) > 5 and r[0].startswith("0x")]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
tot = sum(int(r[isamp] or 0) for r in data)
print("kernel", kname[:90], "| samples", tot, "| instructions", len(data))
top = sorted(((int(r[isamp] or 0), i) for i, r in enumerate(data)), reverse=True)[:top_n]
for s, i in sorted(top, key=lambda t: t[1]):
    r = data[i]
    prev = data[i - 1][isrc].strip() if i else ""
    print(f"{i:5d} {s:6d} {100 * s / max(tot, 1):5.1f}%  ex={r[iex]:>8s} | {prev[:56]:56s} || {r[isrc].strip()[:64]}")
