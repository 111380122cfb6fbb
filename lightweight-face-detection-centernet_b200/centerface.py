"""Drop-in for the reference's ``centerface.py``: same class, constructor, ``__call__`` signature,
return types and public helpers, with the network and the decode running in the sm_100a library.

    from centerface import CenterFace          # the reference (centerface.py:11)
    centerface = CenterFace(h, w, landmarks=True)
    dets, lms = centerface(img, threshold=0.35)   # demo.py:34-38, :76-80

Differences from the reference are limited to its crashes: ``landmarks=False`` works here (the
reference never builds ``self.net`` in that case) and returns ``dets`` only.
"""
from __future__ import annotations

import datetime
import os

import numpy as np

from . import _lib as L
from .engine import Engine, ctdet_decode, decode_threshold  # noqa: F401  (re-exported)

DEFAULT_WEIGHTS = "weight/model_epoch_100.pt"  # cwd-relative like centerface.py:23


def _resolve_weights(weights):
    if weights is not None:
        return weights
    env = os.environ.get("CENTERFACE_B200_WEIGHTS")
    if env:
        return env
    if os.path.exists(DEFAULT_WEIGHTS):
        return DEFAULT_WEIGHTS
    raise FileNotFoundError(
        f"{DEFAULT_WEIGHTS} not found relative to {os.getcwd()} (centerface.py:23 loads it cwd-relative); "
        "pass weights=... or set CENTERFACE_B200_WEIGHTS")


class CenterFace(object):
    # centerface.py:12-15
    mean = np.array([0.408, 0.447, 0.470], dtype=np.float32).reshape(1, 1, 3)
    std = np.array([0.289, 0.274, 0.278], dtype=np.float32).reshape(1, 1, 3)
    print_times = True  # the reference prints "cpu times = ..." on every call (centerface.py:49)
    gpu_resize = True   # cv2.resize on the device (bit-exact with OpenCV's INTER_LINEAR); False = cv2 on the host

    def __init__(self, height, width, landmarks=True, weights=None, device=0, pw_engine=None, engine=None):
        self.landmarks = landmarks
        self.cuda = True
        self.img_h_new, self.img_w_new, self.scale_h, self.scale_w = self.transform(height, width)
        if engine is None:
            if pw_engine is None:
                pw_engine = int(os.environ.get("CENTERFACE_B200_PW", L.CF_PW_TCGEN05))
            engine = Engine(_resolve_weights(weights), max_batch=1, max_h=self.img_h_new, max_w=self.img_w_new,
                            device=device, pw_engine=pw_engine)
        self.net = engine

    def __call__(self, img, threshold=0.2):
        begin = datetime.datetime.now()
        # centerface.py:30-51 + :55-58 in one library call: resize, normalise, net, sigmoid/clamp, decode A (0.3 hard-coded at
        # :77 -- the `threshold` argument is ignored by the reference), NMS 0.3, float32 floor-division by the scales.
        if self.gpu_resize:
            dets, lms = self.net.detect_image_host(np.ascontiguousarray(img, dtype=np.uint8), self.img_h_new, self.img_w_new,
                                                   L.CF_DECODE_A, 0.3, 0.3, np.float32(self.scale_w), np.float32(self.scale_h),
                                                   landmarks=self.landmarks)
        else:
            import cv2
            img = cv2.resize(img, (self.img_w_new, self.img_h_new))  # centerface.py:30
            img = np.ascontiguousarray(img, dtype=np.uint8)[None]
            (dets, lms), = self.net.detect_threshold_host(
                img, L.CF_DECODE_A, 0.3, 0.3, np.float32(self.scale_w), np.float32(self.scale_h), landmarks=self.landmarks)
        end = datetime.datetime.now()
        if self.print_times:
            print("cpu times = ", end - begin)
        if len(dets) == 0:  # centerface.py:60-62
            dets = np.empty(shape=[0, 5], dtype=np.float32)
            lms = np.empty(shape=[0, 10], dtype=np.float32)
        if self.landmarks:
            return dets, lms
        return dets

    def transform(self, h, w):
        """centerface.py:68-71"""
        img_h_new, img_w_new = int(np.ceil(h / 32) * 32), int(np.ceil(w / 32) * 32)
        scale_h, scale_w = img_h_new / h, img_w_new / w
        return img_h_new, img_w_new, scale_h, scale_w

    def decode(self, heatmap, scale, offset, landmark, size, threshold=0.1):
        """centerface.py:73-109 on numpy head maps ([1,1,h,w], [1,2,h,w], [1,2,h,w], [1,10,h,w]),
        executed by the path-A kernel; returns (boxes [n,5], lms [n,10]) like the reference."""
        import torch
        dev = f"cuda:{self.net.device}"
        t = lambda a, c: torch.as_tensor(np.ascontiguousarray(a, np.float32).reshape(1, c, *np.shape(a)[-2:]), device=dev)  # noqa: E731
        cap = L.MAX_CAP
        d, l, n = decode_threshold(t(heatmap, 1), t(scale, 2), t(offset, 2), t(landmark, 10) if self.landmarks else None,
                                   L.CF_DECODE_A, 0.3, 0.3, size, cap=cap)
        n = int(n.item())
        if n < 0:
            raise L.CenterFaceError(f"{-n} candidates exceed the decode cap {cap}")
        if n == 0:
            return ([], []) if self.landmarks else []
        if self.landmarks:
            return d[0, :n].cpu().numpy(), l[0, :n].cpu().numpy()
        return d[0, :n].cpu().numpy()


    def nms(self, boxes, scores, nms_thresh):
        """centerface.py:111-151: greedy IoU NMS -> the list of kept indices (into ``boxes``), in keep order.  Runs on the engine's
        GPU in the reference's float32 arithmetic ("+1" areas, ``ovr >= nms_thresh`` suppresses); equal scores are visited in
        (index descending) order, i.e. ``np.argsort(scores, kind="stable")[::-1]``."""
        import ctypes as C
        boxes = np.ascontiguousarray(np.asarray(boxes, dtype=np.float32)[:, :4])
        scores = np.ascontiguousarray(np.asarray(scores, dtype=np.float32).reshape(-1))
        n = boxes.shape[0]
        if scores.shape[0] != n:
            raise ValueError(f"nms: {n} boxes but {scores.shape[0]} scores")
        keep = np.empty(max(n, 1), dtype=np.int32)
        cnt = C.c_int32(0)
        L.check(L.load().cf_nms_host(self.net.device, C.c_void_p(boxes.ctypes.data), C.c_void_p(scores.ctypes.data), n, float(nms_thresh),
                                     C.c_void_p(keep.ctypes.data), C.byref(cnt)), "cf_nms_host")
        return list(keep[:cnt.value].astype(np.int64))


class CenterFaceNet(object):
    """Model-level drop-in for ``efficientnet_b0()`` as eval_widerface.get_detections uses it
    (eval_widerface.py:76-90): ``model(x)[0]`` is a dict of cuda tensors 'hm','wh','lm','reg'."""

    def __init__(self, weights=None, max_batch=32, max_h=640, max_w=640, device=0, pw_engine=L.CF_PW_TCGEN05):
        self.engine = Engine(_resolve_weights(weights), max_batch, max_h, max_w, device, pw_engine)

    def eval(self):
        return self

    def cuda(self, device=None):
        return self

    def __call__(self, x):
        import torch
        x = x.to(f"cuda:{self.engine.device}", torch.float32).contiguous()
        self.engine.forward(x)
        out = self.engine.heads()
        out.pop("hm_sig")
        return [{k: v.clone() for k, v in out.items()}]


def get_detections(data_batch, model, cuda=True, threshold=0.35):
    """eval_widerface.get_detections (eval_widerface.py:76-90) with path B on the GPU:
    returns a list (one per image) of [n,5] float32 arrays."""
    import torch
    x = data_batch["input"]
    eng = model.engine
    x = x.to(f"cuda:{eng.device}", torch.float32).contiguous()
    eng.forward(x)
    h = eng.heads()
    cap = 1024
    while True:
        dets, _, counts = decode_threshold(h["hm_sig"], h["wh"], h["reg"], None, L.CF_DECODE_B, threshold, 0.3, (640, 640), cap=cap)
        counts = counts.cpu().numpy()
        if (counts >= 0).all() or cap == L.MAX_CAP:
            break
        cap = L.MAX_CAP
    if (counts < 0).any():
        raise L.CenterFaceError(f"{int(-counts.min())} candidates exceed the decode cap {cap}")
    dets = dets.cpu().numpy()
    return [dets[i, :counts[i]].copy() for i in range(len(counts))]
