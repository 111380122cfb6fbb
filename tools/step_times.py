#!/usr/bin/env python
"""Per-launch CUDA-event times of one forward + path-C decode (cf_time_steps), batch 32 @ 640x640 by default."""
import argparse
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("lightweight-face-detection-centernet_b200")

def names(pw, size):
    """Launch names of the engine's plan: bench.py's table for the engine's fused-block masks + the decode kernels."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    return [n for n, _ in bench.launch_table(size, size, pkg._lib.fused_blocks(pw), pkg._lib.dwp_blocks(pw))]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--size", type=int, default=640)
    ap.add_argument("--pw", type=int, default=1)
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    w = os.path.join(ROOT, "tests", "golden", "weights_e100.npz")
    eng = pkg.Engine(w, max_batch=a.batch, max_h=a.size, max_w=a.size, device=0, pw_engine=a.pw)
    x = torch.from_numpy(np.random.RandomState(0).randint(0, 256, size=(a.batch, a.size, a.size, 3), dtype=np.uint8)).cuda()
    if os.environ.get("CF_SMOOTH_INPUT"):  # probe: an image-like input (neighbouring pixels hold near-equal bytes), against the uniform-random default
        g = np.add.outer(np.arange(a.size), np.arange(a.size)) // 5 % 256
        x = torch.from_numpy(np.broadcast_to(g[None, :, :, None], (a.batch, a.size, a.size, 3)).astype(np.uint8).copy()).cuda()
    eng.forward(x)
    eng.decode_topk(100)
    torch.cuda.synchronize()
    ms, cls = eng.time_steps(a.iters)
    nm = names(a.pw, a.size)
    if len(nm) != len(ms):
        nm = [f"step {i}" for i in range(len(ms))]
    cname = {1: "pw", 2: "dw", 3: "stem", 4: "heads", 5: "decode", 6: "fused"}
    print(f"| # | launch | class | us |\n|---|---|---|---|")
    for i, (n, t, c) in enumerate(zip(nm, ms, cls)):
        print(f"| {i} | {n} | {cname.get(c, c)} | {t * 1e3:.1f} |")
    tot = {}
    for t, c in zip(ms, cls):
        tot[c] = tot.get(c, 0.0) + t
    print("total us:", round(sum(ms) * 1e3, 1), {cname.get(c, c): round(v * 1e3, 1) for c, v in tot.items()})


if __name__ == "__main__":
    main()
