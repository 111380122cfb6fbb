#!/usr/bin/env python
"""Warp-stall samples of a warp-specialised kernel by code region: ncu_roles.py rep.ncu-rep [launch] -- prints the hottest
instructions with their stall reasons, and samples / executed instructions between the markers found in the SASS
(UTCHMMA, UTMALDG, LDTM, STTM ...) so that the roles can be told apart."""
import csv
import subprocess
import sys

rep = sys.argv[1]
idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(idx), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[h], [r for r in rows[h + 1:] if len(r) > 5 and r[0].startswith("0x")]
isrc, iex = hdr.index("Source"), hdr.index("Instructions Executed")
isamp = hdr.index("# Samples") if "# Samples" in hdr else hdr.index("Warp Stall Sampling (All Samples)")
stall = [(c, hdr.index(c)) for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(int(r[isamp] or 0) for r in data)
print("samples", tot, "static instructions", len(data))
marks = [i for i, r in enumerate(data) if any(k in r[isrc] for k in ("UTCHMMA", "UTMALDG", "LDTM", "STTM", "UBLKCP", "EXIT"))]
print("markers:", [(i, data[i][isrc].split()[0] if not data[i][isrc].strip().startswith("@") else data[i][isrc].split()[1]) for i in marks][:80])
top = sorted(((int(r[isamp] or 0), i) for i, r in enumerate(data)), reverse=True)[:top_n]
for s, i in sorted(top, key=lambda t: t[1]):
    r = data[i]
    why = {c[6:]: r[k] for c, k in stall if r[k] not in ("0", "")}
    print(f"{i:5d} {s:6d} {100 * s / max(tot, 1):5.1f}% ex={r[iex]:>9s} | {r[isrc].strip()[:60]:60s} {why}")
if len(sys.argv) > 4:  # region boundaries given: a,b,c,...
    b = [int(x) for x in sys.argv[4].split(",")]
    for a, z in zip(b[:-1], b[1:]):
        e = sum(int(data[i][iex] or 0) for i in range(a, z))
        s = sum(int(data[i][isamp] or 0) for i in range(a, z))
        print(f"region [{a},{z}): inst {e / 1e6:8.2f}M samples {s} ({100 * s / tot:.1f}%)")
