#!/usr/bin/env python
"""Development check of the whole-block fused kernel (k_mbf): the default engine with CF_MBF=<mask> against the layer-wise
schedule (CF_MBF=0) on the same input -- every block tap and the heads -- then per-launch times of the fused plan.
    python tools/mbf_check.py --mask 0x2 [--batch 4] [--size 640] [--time]"""
import argparse
import importlib
import os
import sys

import cv2
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("lightweight-face-detection-centernet_b200")

ap = argparse.ArgumentParser()
ap.add_argument("--mask", default="0x2")
ap.add_argument("--mbd", default="0", help="blocks whose depth-wise + projection run as one kernel (CF_MBD)")
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--size", type=int, default=640)
ap.add_argument("--time", action="store_true")
ap.add_argument("--tbatch", type=int, default=32)
a = ap.parse_args()
W = os.path.join(ROOT, "tests", "golden", "weights_e100.npz")
z = np.load(os.path.join(ROOT, "tests", "golden", "images_jpeg.npz"))
names = ["27", "8", "1", "17", "2"]
imgs = np.stack([cv2.resize(cv2.imdecode(z["img_" + names[i % 5]], cv2.IMREAD_COLOR), (a.size, a.size)) for i in range(a.batch)])
x = torch.from_numpy(imgs).cuda()


def run(mask, mbd="0"):
    os.environ["CF_MBF"] = mask
    os.environ["CF_MBD"] = mbd
    eng = pkg.Engine(W, max_batch=a.batch, max_h=a.size, max_w=a.size, device=0)
    eng.forward(x)
    torch.cuda.synchronize()
    out = {f"block{i}": eng.tap(f"block{i}").clone() for i in range(12)}
    out.update({k: v.clone() for k, v in eng.heads().items()})
    dets, inds = eng.decode_topk(100)
    out["inds"] = inds.clone()
    torch.cuda.synchronize()
    return out


ref = run("0")
got = run(a.mask, a.mbd)
bad = False
for k in ref:
    if k == "inds":
        same = bool((ref[k] == got[k]).all())
        print(f"{k:8s} identical={same}")
        continue
    r, g = ref[k].double(), got[k].double()
    d = (r - g).abs().max().item()
    m = r.abs().max().item()
    nan = int(torch.isnan(got[k]).sum().item())
    print(f"{k:8s} max|ref|={m:10.4g} max|diff|={d:10.4g} rel={d / max(m, 1e-30):9.3g} nan={nan}")
    if nan or d / max(m, 1e-30) > 1e-4:
        bad = True
print("MBF CHECK", "FAILED" if bad else "ok", "mask", a.mask, "mbd", a.mbd)
if a.time:
    os.environ["CF_MBF"] = a.mask
    os.environ["CF_MBD"] = a.mbd
    eng = pkg.Engine(W, max_batch=a.tbatch, max_h=a.size, max_w=a.size, device=0)
    xt = torch.from_numpy(np.random.RandomState(0).randint(0, 256, size=(a.tbatch, a.size, a.size, 3), dtype=np.uint8)).cuda()
    eng.forward(xt)
    eng.decode_topk(100)
    torch.cuda.synchronize()
    ms, cls = eng.time_steps(10)
    print("per-launch us:", [round(t * 1e3, 1) for t in ms])
    print("fused launches us:", [round(t * 1e3, 1) for t, c in zip(ms, cls) if c == 6], "total us", round(sum(ms) * 1e3, 1))
