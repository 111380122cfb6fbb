#!/usr/bin/env python
"""Pipeline trace of the fused MBConv kernel (k_mbf): clock64 of every hand-off of CTA 0 for a window of jobs, printed as per-event
deltas (cycles) so that the critical path of the warp-specialised pipeline can be read off.
    python tools/mbf_trace.py --mask 0x2 --j0 200 --nj 40 [--batch 32]"""
import argparse
import ctypes as C
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--mask", default="0x2")
ap.add_argument("--mbd", default="0")
ap.add_argument("--j0", type=int, default=200)
ap.add_argument("--nj", type=int, default=40)
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--size", type=int, default=640)
a = ap.parse_args()
os.environ["CF_MBF"] = a.mask
os.environ["CF_MBD"] = a.mbd
os.environ["CF_MBF_TRACE"] = f"{a.j0},{a.nj}"
pkg = importlib.import_module("lightweight-face-detection-centernet_b200")
eng = pkg.Engine(os.path.join(ROOT, "tests", "golden", "weights_e100.npz"), max_batch=a.batch, max_h=a.size, max_w=a.size, device=0)
x = torch.from_numpy(np.random.RandomState(0).randint(0, 256, size=(a.batch, a.size, a.size, 3), dtype=np.uint8)).cuda()
eng.forward(x)
eng.forward(x)  # the second (warm) forward overwrites the trace
torch.cuda.synchronize()
buf = (C.c_uint64 * (a.nj * 32))()
eng.lib.cf_debug_mbf_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
rc = eng.lib.cf_debug_mbf_trace(eng.h, buf, a.nj)
assert rc == 0, rc
t = np.frombuffer(buf, dtype=np.uint64).reshape(a.nj, 32).astype(np.int64)
names = {0: "P.xempty", 1: "P.tma", 2: "S.xfull", 3: "S.aempty", 4: "S.arrive", 5: "X.afull", 6: "X.eempty", 7: "X.commit", 8: "T.efull", 9: "T.ldtm",
         19: "T.swish", 10: "T.bar1", 11: "T.bar2", 12: "T.dw", 13: "T.dfree", 14: "T.dfull", 15: "J.dfull", 16: "J.commit", 17: "E.start", 18: "E.end", 24: "X.top", 25: "X.issue", 26: "J.pempty", 27: "T.top", 28: "T.end"}
order = [0, 1, 2, 3, 4, 24, 5, 6, 7, 8, 9, 19, 10, 11, 12, 13, 14, 15, 16, 17, 18]
t0 = t[t > 0].min()
print("job  " + " ".join(f"{names[e]:>9s}" for e in order))
for j in range(a.nj):
    print(f"{a.j0 + j:4d} " + " ".join(f"{(t[j, e] - t0) if t[j, e] else -1:9d}" for e in order))
# steady-state rate and per-role step durations
def col(e):
    v = t[:, e]
    return v[v > 0]
for e in (1, 4, 7, 14):
    v = col(e)
    if len(v) > 2:
        print(f"{names[e]:9s}: mean interval {np.diff(v).mean():8.1f} cycles over {len(v)} jobs")
pairs = [(27, 8), (14, 28), (6, 25), (25, 7), (15, 26), (26, 16), (7, 24), (17, 18), (16, 18), (14, 15), (15, 16), (0, 1), (2, 3), (3, 4), (5, 6), (6, 7), (8, 9), (9, 19), (19, 10), (10, 11), (11, 12), (12, 13), (13, 14), (7, 8), (4, 5), (1, 2)]
for a_, b_ in pairs:
    m = (t[:, a_] > 0) & (t[:, b_] > 0)
    if m.any():
        d = t[m, b_] - t[m, a_]
        print(f"{names[a_]:>9s} -> {names[b_]:9s}: mean {d.mean():8.1f}  min {d.min():6d}  max {d.max():6d}")
# whole-run view: completion rate over windows of 64 jobs
v = t[:, 14]
v = v[v > 0]
if len(v) > 128:
    print("T.dfull span", int(v.max() - v.min()), "cycles for", len(v), "jobs;", "first event at", int(t[t > 0].min() - t0))
    for k in range(0, len(v) - 64, 64):
        print(f"  jobs {a.j0 + k:4d}..{a.j0 + k + 63:4d}: {(v[k + 64] - v[k]) / 64:7.1f} cycles/job")
