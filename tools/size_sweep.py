#!/usr/bin/env python
"""Throughput-vs-roofline curve over input size and batch (BASELINE.json configs[3] and configs[4] per-GPU shards).

For every (batch, H, W) point: u8 batches resident in HBM (up to 4 rotating input batches; from batch 32 on the activations alone are 3-49 GB per step),
3 warm-up steps, K steps of network + sigmoid/clamp + path-C top-100 decode timed with CUDA events on the launching
stream, then the layer-wise algorithmic bytes of that size (cf_work_model) / time against the measured
HBM peak, and the per-class split (cf_time_class).  The config-4 point additionally times the eval_widerface flow:
device letter-box of 480x640 frames (dataset/dataset.py:130-134) -> network -> path-B decode (eval_widerface.py:92-152).

Prints one JSON line per point and a markdown table at the end; never run under a profiler for the numbers.
"""
import argparse
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("lightweight-face-detection-centernet_b200")
L = pkg._lib
CLASS_NAMES = {1: "pw", 2: "dw", 3: "stem", 4: "heads", 5: "decode"}

# (label, batch, H, W): configs[1]; configs[3] (VGA letter-boxed onto the 640x640 canvas, 128 per GPU);
# configs[4] 320-max-side shards of 128 (centerface.py:69 pads H, W to multiples of 32) + batch scaling at 320x320
# (largest allocation last, so that the earlier points are on record whatever happens to it)
POINTS = [("C2 640x640", 32, 640, 640),
          ("C5 320x320", 128, 320, 320), ("C5 320x256", 128, 320, 256), ("C5 256x320", 128, 256, 320),
          ("C5 320x192", 128, 320, 192), ("320x320", 32, 320, 320), ("320x320", 8, 320, 320), ("320x320", 1, 320, 320),
          ("640x640", 8, 640, 640), ("640x640", 1, 640, 640),
          ("480x640 (no canvas)", 128, 480, 640), ("C4 640x640 canvas", 128, 640, 640)]


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "MEASURED_PEAKS.json"
    return 6650.0, "fallback (B200_PROFILING.md)"


def timed(fn, steps, warmup=3):
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(warmup + i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def f5_batch(n, h=640, w=640):
    """SURVEY.md 8d parity inputs: the five bundled JPEGs stretched to (w, h), cycled with RandomState(1234) horizontal
    flips (p = .5) and integer rolls in [0, 32) -- real head maps, so the decoders see realistic candidate counts."""
    import cv2
    import numpy as np
    z = np.load(os.path.join(ROOT, "tests", "golden", "images_jpeg.npz"))
    base = [cv2.resize(cv2.imdecode(z["img_" + k], cv2.IMREAD_COLOR), (w, h)) for k in ("1", "17", "2", "27", "8")]
    rng = np.random.RandomState(1234)
    imgs = []
    for i in range(n):
        im = base[i % 5]
        if rng.rand() < 0.5:
            im = im[:, ::-1]
        imgs.append(np.roll(im, (rng.randint(0, 32), rng.randint(0, 32)), axis=(0, 1)))
    return torch.from_numpy(np.ascontiguousarray(np.stack(imgs))).cuda()


def f5_decode_point(eng, steps, B=32):
    """Network + each decode path on the F5-derived batch: decode cost depends on the candidate count."""
    x = f5_batch(B)
    out = {"point": "F5-derived 640x640", "batch": B}

    def net(i):
        eng.forward(x)

    def path_c(i):
        eng.forward(x)
        eng.decode_topk(100)

    def path(variant, thr, lm):
        def fn(i):
            eng.forward(x)
            h = eng.heads()
            return pkg.decode_threshold(h["hm_sig"], h["wh"], h["reg"], h["lm"] if lm else None, variant, thr, 0.3, (640, 640), cap=1024)
        return fn

    t_net = timed(net, steps)
    out["network_ms"] = round(t_net, 4)
    for name, fn in (("pathC_top100", path_c), ("pathA_thr0.3_lm_nms", path(pkg.CF_DECODE_A, 0.3, True)),
                     ("pathB_thr0.35_nms", path(pkg.CF_DECODE_B, 0.35, False))):
        t = timed(fn, steps)
        out[name] = {"ms_per_step": round(t, 4), "images_per_s": round(B / (t * 1e-3), 1), "decode_ms": round(t - t_net, 4)}
    _, _, counts = path(pkg.CF_DECODE_A, 0.3, True)(0)
    c = counts.cpu().numpy()
    out["pathA_dets_per_image"] = {"min": int(c.min()), "mean": float(c.mean()), "max": int(c.max())}
    return out


def one_point(eng, a, gen, peak, label, B, H, W):
    # rotate over up to 4 input batches (the engine keeps the launch plans of its last 5 input buffers, as bench.py relies on);
    # at small batch the rotation is below the 126 MB L2 -- those points are latency-bound, not bandwidth-bound
    n_rot = max(2, min(4, -(-160_000_000 // (B * H * W * 3))))
    xs = [torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device="cuda", generator=gen) for _ in range(n_rot)]

    def step(i):
        eng.forward(xs[i % n_rot])
        eng.decode_topk(100)

    ms = timed(step, a.steps)
    by, fl = L.work_model(H, W, L.CF_IN_U8_HWC, 0, a.pw)
    cls = {}
    for c, name in CLASS_NAMES.items():
        t, n = eng.time_class(c, iters=3)
        if n:
            cb, _ = L.work_model(H, W, L.CF_IN_U8_HWC, c, a.pw)
            cls[name] = {"ms": round(t, 4), "frac_hbm": round(cb * B / (t * 1e-3) / 1e9 / peak, 3)}
    row = {"point": label, "batch": B, "h": H, "w": W, "ms_per_step": round(ms, 4), "images_per_s": round(B / (ms * 1e-3), 1),
           "alg_MB_per_image": round(by / 1e6, 2), "GBps": round(by * B / (ms * 1e-3) / 1e9, 1),
           "frac_hbm": round(by * B / (ms * 1e-3) / 1e9 / peak, 4), "TFLOPs": round(fl * B / (ms * 1e-3) / 1e12, 2),
           "act_GB_per_step": round(by * B / 1e9, 3), "input_rotation": n_rot, "classes": cls}
    if label.startswith("C4"):
        # the loader's flow: 480x640 frames -> device letter-box -> network -> path B (threshold 0.35, NMS 0.3)
        frames = [torch.randint(0, 256, (B, 480, 640, 3), dtype=torch.uint8, device="cuda", generator=gen) for _ in range(2)]

        def step_b(i):
            eng.forward(pkg.letterbox_u8(frames[i % 2], 640, 640))
            h = eng.heads()
            pkg.decode_threshold(h["hm_sig"], h["wh"], h["reg"], None, pkg.CF_DECODE_B, 0.35, 0.3, (640, 640), cap=1024)

        ms_b = timed(step_b, a.steps)
        row["letterbox_pathB_ms_per_step"] = round(ms_b, 4)
        row["letterbox_pathB_images_per_s"] = round(B / (ms_b * 1e-3), 1)
    return row


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--pw", type=int, default=L.CF_PW_TCGEN05)
    ap.add_argument("--max-batch", type=int, default=128)
    ap.add_argument("--out", default=None, help="markdown table path")
    a = ap.parse_args()
    peak, peak_src = hbm_peak()
    w = os.path.join(ROOT, "tests", "golden", "weights_e100.npz")
    eng = pkg.Engine(w, max_batch=a.max_batch, max_h=640, max_w=640, device=0, pw_engine=a.pw)
    gen = torch.Generator(device="cuda").manual_seed(0)  # seeded uniform u8 noise (SURVEY.md 8d throughput inputs)
    rows = []
    f5 = None
    try:
        f5 = f5_decode_point(eng, a.steps)
        print(json.dumps(f5), flush=True)
    except Exception as ex:  # keep the size sweep going
        print(json.dumps({"point": "F5-derived 640x640", "error": str(ex)}), flush=True)
    for label, B, H, W in POINTS:
        if B > a.max_batch:
            continue
        try:
            rows.append(one_point(eng, a, gen, peak, label, B, H, W))
        except Exception as ex:
            print(json.dumps({"point": label, "batch": B, "h": H, "w": W, "error": str(ex)}), flush=True)
            break
        print(json.dumps(rows[-1]), flush=True)
        torch.cuda.empty_cache()
    eng.close()
    md = [f"HBM peak {peak} GB/s ({peak_src}); engine {a.pw}; {a.steps} timed steps per point after 3 warm-up steps.", "",
          "| point | batch | HxW | ms/step | images/s | alg. MB/image | GB/s | frac of HBM peak | TFLOP/s | pw / dw / stem / heads / decode ms |",
          "|---|---|---|---|---|---|---|---|---|---|"]
    for r in rows:
        c = r["classes"]
        md.append(f"| {r['point']} | {r['batch']} | {r['h']}x{r['w']} | {r['ms_per_step']:.3f} | {r['images_per_s']:.0f} | "
                  f"{r['alg_MB_per_image']} | {r['GBps']:.0f} | {r['frac_hbm']:.3f} | {r['TFLOPs']} | "
                  + " / ".join(f"{c[k]['ms']:.3f}" if k in c else "-" for k in ("pw", "dw", "stem", "heads", "decode")) + " |")
    for r in rows:
        if "letterbox_pathB_ms_per_step" in r:
            md += ["", f"Config-4 loader flow at batch {r['batch']} (device letter-box of 480x640 frames -> network -> path-B decode): "
                       f"{r['letterbox_pathB_ms_per_step']:.3f} ms/step = {r['letterbox_pathB_images_per_s']:.0f} images/s."]
    if f5 and "error" not in f5:
        md += ["", f"F5-derived batch of {f5['batch']} at 640x640 (five JPEGs cycled with flips / rolls; {f5['pathA_dets_per_image']['mean']:.0f} path-A "
                   f"detections per image on average, max {f5['pathA_dets_per_image']['max']}): network alone {f5['network_ms']:.3f} ms; "
                   + "; ".join(f"{k} {f5[k]['ms_per_step']:.3f} ms = {f5[k]['images_per_s']:.0f} images/s (decode {f5[k]['decode_ms']:.3f} ms)"
                               for k in ("pathC_top100", "pathA_thr0.3_lm_nms", "pathB_thr0.35_nms")) + "."]
    text = "\n".join(md)
    print(text)
    if a.out:
        with open(a.out, "w") as f:
            f.write(text + "\n")


if __name__ == "__main__":
    main()
