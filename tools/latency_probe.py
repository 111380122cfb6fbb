#!/usr/bin/env python
"""Single-image latency of the drop-in CenterFace.__call__ (host image in, boxes out) and of batch-1..32 detect calls."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("lightweight-face-detection-centernet_b200")
import cv2
z = np.load(os.path.join(ROOT, "tests", "golden", "images_jpeg.npz"))
w = os.path.join(ROOT, "tests", "golden", "weights_e100.npz")
pkg.CenterFace.print_times = False
for n in ("8", "27", "17"):
    img = cv2.imdecode(z["img_" + n], cv2.IMREAD_COLOR)
    cf = pkg.CenterFace(img.shape[0], img.shape[1], weights=w)
    for gpu_resize in (True, False):
        cf.gpu_resize = gpu_resize
        for _ in range(5): cf(img)
        t0 = time.perf_counter()
        for _ in range(50): d, l = cf(img)
        dt = (time.perf_counter() - t0) / 50
        print(f"CenterFace.__call__ {n}.jpg {img.shape[:2]} gpu_resize={gpu_resize}: {dt*1e3:.3f} ms/image, {len(d)} dets", flush=True)
    cf.net.close()
e = pkg.Engine(w, max_batch=32)
for B in (1, 2, 4, 8, 16, 32):
    u8 = np.random.RandomState(0).randint(0, 256, (B, 640, 640, 3), dtype=np.uint8)
    for _ in range(3): e.detect_topk_host(u8)
    t0 = time.perf_counter()
    for _ in range(20): e.detect_topk_host(u8)
    dt = (time.perf_counter() - t0) / 20
    print(f"detect_topk_host batch {B}: {dt*1e3:.3f} ms/call, {B/dt:.0f} img/s", flush=True)
