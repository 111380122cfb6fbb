#!/usr/bin/env python
"""Device time of one resident-input forward + path-C decode at small batches (configs[0]: the demo.py drop-in case), with and
without the CUDA-graph replay of the forward chain (CF_GRAPH=0 in the environment = eager launches)."""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("lightweight-face-detection-centernet_b200")
w = os.path.join(ROOT, "tests", "golden", "weights_e100.npz")
for B in (1, 2, 8):
    eng = pkg.Engine(w, max_batch=B, max_h=640, max_w=640, device=0)
    x = torch.from_numpy(np.random.RandomState(0).randint(0, 256, size=(B, 640, 640, 3), dtype=np.uint8)).cuda()
    for _ in range(5):
        eng.forward(x)
        eng.decode_topk(100)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50):
        eng.forward(x)
        eng.decode_topk(100)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 50
    print(f"CF_GRAPH={os.environ.get('CF_GRAPH', '1')} batch {B} @ 640x640: {ms:.4f} ms per forward + decode ({B / ms * 1e3:.0f} img/s)", flush=True)
    eng.close()
