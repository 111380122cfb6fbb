#!/usr/bin/env python
"""Run N forwards (+ path-C decode) of a batch -- the workload for `ncu -k regex:... -s ... -c ...` captures."""
import argparse
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("lightweight-face-detection-centernet_b200")
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--size", type=int, default=640)
ap.add_argument("--n", type=int, default=2)
ap.add_argument("--pw", type=int, default=1)
a = ap.parse_args()
eng = pkg.Engine(os.path.join(ROOT, "tests", "golden", "weights_e100.npz"), max_batch=a.batch, max_h=a.size, max_w=a.size, device=0, pw_engine=a.pw)
x = torch.from_numpy(np.random.RandomState(0).randint(0, 256, size=(a.batch, a.size, a.size, 3), dtype=np.uint8)).cuda()
for _ in range(a.n):
    eng.forward(x)
    eng.decode_topk(100)
torch.cuda.synchronize()
print("done", eng.launches)
