"""Python handle over the C ABI (one ``cf_engine`` per GPU).  PyTorch is used only for device
memory, streams and zero-copy views of the engine's buffers."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .weights import load_state_dict, pack_weights


class _DevView:
    """Expose a raw device pointer through __cuda_array_interface__ so torch can wrap it."""

    def __init__(self, ptr, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}


def _host_ptr(a):
    """numpy array or (pinned) torch CPU tensor -> (address, keepalive)."""
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data, a
    assert a.device.type == "cpu" and a.is_contiguous()
    return a.data_ptr(), a


class Engine:
    """Owns a ``cf_engine``: weights on the device + activation buffers for ``max_batch`` images of
    ``max_h x max_w``.  Replaces efficientnet_b0() + load_state_dict + .cuda()
    (centerface.py:19-24)."""

    def __init__(self, weights, max_batch=1, max_h=640, max_w=640, device=0, pw_engine=L.CF_PW_TCGEN05):
        self.lib = L.load()
        if isinstance(weights, (str, bytes)) and not isinstance(weights, bytes):
            weights = load_state_dict(weights)
        blob = weights if isinstance(weights, bytes) else pack_weights(weights)
        if len(blob) != self.lib.cf_weights_blob_bytes():
            raise L.CenterFaceError(f"weight blob is {len(blob)} bytes, library expects {self.lib.cf_weights_blob_bytes()}")
        self.device = int(device)
        self.max_batch, self.max_h, self.max_w = int(max_batch), int(max_h), int(max_w)
        self.pw_engine = int(pw_engine)
        h = C.c_void_p()
        buf = C.create_string_buffer(blob, len(blob))
        L.check(self.lib.cf_create(C.cast(buf, C.c_void_p), len(blob), self.device, self.max_batch, self.max_h,
                                   self.max_w, self.pw_engine, C.byref(h)), "cf_create")
        self.h = h
        self.shape = None  # (B,H,W) of the last forward
        self._keep = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.cf_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- device-side API (torch tensors on this engine's GPU) -------------------------------
    def _stream(self):
        import torch
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def forward(self, x):
        """EfficientNet.forward (model/centernet.py:263-280).  x: cuda float32 [B,3,H,W] (normalised,
        what the reference feeds its net) or cuda uint8 [B,H,W,3] raw BGR (normalisation fused)."""
        import torch
        assert x.is_cuda and x.device.index == self.device and x.is_contiguous()
        if x.dtype == torch.uint8:
            B, H, W, c = x.shape
            fmt = L.CF_IN_U8_HWC
        else:
            assert x.dtype == torch.float32
            B, c, H, W = x.shape
            fmt = L.CF_IN_F32_NCHW
        assert c == 3
        L.check(self.lib.cf_forward(self.h, C.c_void_p(x.data_ptr()), fmt, B, H, W, self._stream()), "cf_forward")
        self._keep = x
        self.shape = (B, H, W)

    def heads(self):
        """Zero-copy views of the reference's output dict (+ 'hm_sig'), model/centernet.py:277-280."""
        import torch
        p = [C.c_void_p() for _ in range(5)]
        L.check(self.lib.cf_heads(self.h, *[C.byref(q) for q in p]), "cf_heads")
        B, H, W = self.shape
        h4, w4 = H // 4, W // 4
        dev = f"cuda:{self.device}"
        out = {}
        for name, q, c in zip(("hm", "wh", "lm", "reg", "hm_sig"), p, (1, 2, 10, 2, 1)):
            out[name] = torch.as_tensor(_DevView(q.value, (B, c, h4, w4)), device=dev)
        return out

    def tap(self, name):
        """NHWC fp32 view [B,h,w,c] of an intermediate activation (parity taps)."""
        import torch
        p, h, w, c = C.c_void_p(), C.c_int32(), C.c_int32(), C.c_int32()
        L.check(self.lib.cf_tap(self.h, name.encode(), C.byref(p), C.byref(h), C.byref(w), C.byref(c)), "cf_tap")
        return torch.as_tensor(_DevView(p.value, (self.shape[0], h.value, w.value, c.value)), device=f"cuda:{self.device}")

    def decode_topk(self, K=100):
        """ctdet_decode (centerface_ext.py:52-82) on the heads of the last forward -> dets [B,K,6], inds [B,K]."""
        import torch
        B = self.shape[0]
        dev = f"cuda:{self.device}"
        dets = torch.empty((B, K, 6), dtype=torch.float32, device=dev)
        inds = torch.empty((B, K), dtype=torch.int32, device=dev)
        L.check(self.lib.cf_decode_topk(self.h, K, C.c_void_p(dets.data_ptr()), C.c_void_p(inds.data_ptr()),
                                        self._stream()), "cf_decode_topk")
        return dets, inds

    def time_class(self, which, iters=5):
        """Mean device ms of one replay of a kernel class on the last forward's activations."""
        ms, n = C.c_float(), C.c_int32()
        L.check(self.lib.cf_time_class(self.h, which, iters, self._stream(), C.byref(ms), C.byref(n)), "cf_time_class")
        return ms.value, n.value

    def time_steps(self, iters=5):
        """Per-launch device ms of the current plan (network + path-C decode, in launch order) -> (ms[], cls[])."""
        cap = 128
        ms, cls, n = (C.c_float * cap)(), (C.c_int32 * cap)(), C.c_int32()
        L.check(self.lib.cf_time_steps(self.h, iters, self._stream(), ms, cls, cap, C.byref(n)), "cf_time_steps")
        return list(ms[:n.value]), list(cls[:n.value])

    @property
    def launches(self):
        return int(self.lib.cf_launch_count(self.h))

    # ---- host-side API (numpy / pinned buffers; the calls the e2e figure times) --------------
    def detect_topk_host(self, images, K=100, out_dets=None, out_inds=None):
        """u8 BGR [B,H,W,3] host batch at network size -> dets [B,K,6], inds [B,K] (host)."""
        B, H, W, c = images.shape
        assert c == 3
        if out_dets is None:
            out_dets = np.empty((B, K, 6), np.float32)
        if out_inds is None:
            out_inds = np.empty((B, K), np.int32)
        ip, k0 = _host_ptr(images)
        dp, k1 = _host_ptr(out_dets)
        np_, k2 = _host_ptr(out_inds)
        L.check(self.lib.cf_detect_topk_host(self.h, C.c_void_p(ip), B, H, W, K, C.c_void_p(dp), C.c_void_p(np_)),
                "cf_detect_topk_host")
        self._drained()
        self.shape = (B, H, W)
        return out_dets, out_inds

    def submit_topk_host(self, images, K, out_dets, out_inds=None):
        """Pipelined form: enqueue one host batch (pinned buffers; they must outlive the matching wait_host)."""
        B, H, W, c = images.shape
        assert c == 3
        ip, k0 = _host_ptr(images)
        dp, k1 = _host_ptr(out_dets)
        np_, k2 = _host_ptr(out_inds) if out_inds is not None else (None, None)
        L.check(self.lib.cf_submit_topk_host(self.h, C.c_void_p(ip), B, H, W, K, C.c_void_p(dp),
                                             C.c_void_p(np_) if np_ else None), "cf_submit_topk_host")
        self.shape = (B, H, W)
        self._inflight = getattr(self, "_inflight", [])
        self._inflight.append((k0, k1, k2))

    def comm_init(self, group=None):
        """Join the NCCL communicator of the library (one rank per process / GPU): rank 0 creates the id, torch.distributed (any
        backend) carries the 128 bytes to the others -- the only use of torch.distributed on the data path is this bootstrap."""
        import torch
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        idb = (C.c_ubyte * 128)()
        if rank == 0:
            L.check(self.lib.cf_comm_unique_id(idb), "cf_comm_unique_id")
        t = torch.tensor(list(bytes(idb)), dtype=torch.uint8)
        if dist.get_backend(group) == "nccl":
            t = t.to(f"cuda:{self.device}")
        dist.broadcast(t, src=0, group=group)
        raw = bytes(t.cpu().tolist())
        L.check(self.lib.cf_comm_init(self.h, world, rank, C.create_string_buffer(raw, 128)), "cf_comm_init")
        self.comm_world, self.comm_rank = world, rank

    def submit_topk_gather_host(self, images, K, out_dets_all, out_inds=None):
        """submit_topk_host + the exchange step: out_dets_all [world * B, K, 6] (pinned) receives every rank's boxes."""
        B, H, W, c = images.shape
        assert c == 3 and out_dets_all.shape[0] == self.comm_world * B
        ip, k0 = _host_ptr(images)
        dp, k1 = _host_ptr(out_dets_all)
        np_, k2 = _host_ptr(out_inds) if out_inds is not None else (None, None)
        L.check(self.lib.cf_submit_topk_gather_host(self.h, C.c_void_p(ip), B, H, W, K, C.c_void_p(dp),
                                                    C.c_void_p(np_) if np_ else None), "cf_submit_topk_gather_host")
        self.shape = (B, H, W)
        self._inflight = getattr(self, "_inflight", [])
        self._inflight.append((k0, k1, k2))

    def wait_host(self):
        """Block until the oldest submit_topk_host has delivered its outputs."""
        L.check(self.lib.cf_wait_host(self.h), "cf_wait_host")
        if getattr(self, "_inflight", None):
            self._inflight.pop(0)

    def _drained(self):
        """The detect_* calls wait for every pending submission in C: the keep-alive list of the pipelined API is stale afterwards."""
        self._inflight = []

    def detect_image_host(self, image, net_h, net_w, variant, threshold, nms_threshold=0.3, scale_w=0.0, scale_h=0.0, cap=1024,
                          landmarks=True):
        """One host u8 BGR image [h,w,3] at its own size -> (dets [n,5], lms [n,10] | None): resize on the device (bit-exact
        cv2.resize), network, decode, NMS, //scale -- the body of CenterFace.__call__ (centerface.py:29-62)."""
        h, w, c = image.shape
        assert c == 3 and image.dtype == np.uint8
        image = np.ascontiguousarray(image)
        want_lms = landmarks and variant == L.CF_DECODE_A
        while True:
            dets = np.empty((cap, 5), np.float32)
            lms = np.empty((cap, 10), np.float32) if want_lms else None
            count = np.empty((1,), np.int32)
            L.check(self.lib.cf_detect_image_host(
                self.h, C.c_void_p(image.ctypes.data), h, w, net_h, net_w, variant, threshold, nms_threshold, scale_w, scale_h, cap,
                C.c_void_p(dets.ctypes.data), C.c_void_p(lms.ctypes.data) if want_lms else None, C.c_void_p(count.ctypes.data)),
                "cf_detect_image_host")
            self._drained()
            self.shape = (1, net_h, net_w)
            n = int(count[0])
            if n >= 0:
                return dets[:n].copy(), (lms[:n].copy() if want_lms else None)
            if -n > L.MAX_CAP or cap == L.MAX_CAP:
                raise L.CenterFaceError(f"{-n} pixels above the threshold; the decode kernel caps at {L.MAX_CAP}")
            cap = L.MAX_CAP

    def detect_threshold_host(self, images, variant, threshold, nms_threshold=0.3, scale_w=0.0, scale_h=0.0,
                              cap=1024, landmarks=True):
        """u8 BGR [B,H,W,3] host batch -> list of (dets [n,5], lms [n,10] | None) per image.
        Variant A = CenterFace.decode+nms (+ //scale), variant B = eval_widerface.decode+nms."""
        B, H, W, c = images.shape
        assert c == 3
        want_lms = landmarks and variant == L.CF_DECODE_A
        while True:
            dets = np.empty((B, cap, 5), np.float32)
            lms = np.empty((B, cap, 10), np.float32) if want_lms else None
            counts = np.empty((B,), np.int32)
            ip, _k = _host_ptr(images)
            L.check(self.lib.cf_detect_threshold_host(
                self.h, C.c_void_p(ip), B, H, W, variant, threshold, nms_threshold, scale_w, scale_h, cap,
                C.c_void_p(dets.ctypes.data), C.c_void_p(lms.ctypes.data) if want_lms else None,
                C.c_void_p(counts.ctypes.data)), "cf_detect_threshold_host")
            self._drained()
            self.shape = (B, H, W)
            if (counts >= 0).all():
                break
            need = int(-counts.min())
            if need > L.MAX_CAP:
                raise L.CenterFaceError(f"{need} pixels above the threshold in one image; the decode kernel caps at {L.MAX_CAP}")
            cap = L.MAX_CAP
        return [(dets[i, :counts[i]].copy(), lms[i, :counts[i]].copy() if want_lms else None) for i in range(B)]


def resize_u8(images, dh, dw):
    """cv2.resize(img, (dw, dh)) (INTER_LINEAR) for a cuda uint8 batch [B,h,w,3] -> [B,dh,dw,3], bit-exact with OpenCV."""
    import torch
    lib = L.load()
    assert images.is_cuda and images.dtype == torch.uint8 and images.is_contiguous() and images.shape[3] == 3
    B, sh, sw, _ = images.shape
    tab = np.empty((3 * dw + 4 * dh,), np.int32)
    area2 = C.c_int32()
    L.check(lib.cf_resize_tables(sh, sw, dh, dw, C.c_void_p(tab.ctypes.data), tab.size, C.byref(area2)), "cf_resize_tables")
    t_dev = torch.from_numpy(tab).to(images.device)
    out = torch.empty((B, dh, dw, 3), dtype=torch.uint8, device=images.device)
    with torch.cuda.device(images.device):
        L.check(lib.cf_resize_u8(C.c_void_p(images.data_ptr()), B, sh, sw, C.c_void_p(out.data_ptr()), dh, dw, C.c_void_p(t_dev.data_ptr()),
                                 area2.value, C.c_void_p(torch.cuda.current_stream(images.device).cuda_stream)), "cf_resize_u8")
    return out


def warp_affine_u8(images, M, dw, dh):
    """cv2.warpAffine(img, M, (dw, dh), flags=cv2.INTER_LINEAR) (constant-0 border) for a cuda uint8 batch [B,h,w,3] ->
    [B,dh,dw,3], bit-exact with OpenCV; M is the forward 2x3 matrix (the reference's trans_input, dataset/dataset.py:130-134)."""
    import torch
    lib = L.load()
    assert images.is_cuda and images.dtype == torch.uint8 and images.is_contiguous() and images.shape[3] == 3
    B, sh, sw, _ = images.shape
    Mh = np.ascontiguousarray(np.asarray(M, dtype=np.float64).reshape(6))
    tab = np.empty((2 * dw + 2 * dh,), np.int32)
    L.check(lib.cf_warp_affine_tables(C.c_void_p(Mh.ctypes.data), dh, dw, C.c_void_p(tab.ctypes.data), tab.size), "cf_warp_affine_tables")
    t_dev = torch.from_numpy(tab).to(images.device)
    out = torch.empty((B, dh, dw, 3), dtype=torch.uint8, device=images.device)
    with torch.cuda.device(images.device):
        L.check(lib.cf_warp_affine_u8(C.c_void_p(images.data_ptr()), B, sh, sw, C.c_void_p(out.data_ptr()), dh, dw,
                                      C.c_void_p(t_dev.data_ptr()), C.c_void_p(torch.cuda.current_stream(images.device).cuda_stream)),
                "cf_warp_affine_u8")
    return out


def letterbox_matrix(h, w, out_w, out_h):
    """trans_input of the reference's loader for the validation split: get_affine_transform((w/2, h/2), max(h, w), 0,
    [out_w, out_h]) (dataset/dataset.py:113-131, utils/image.py:27-61), built with the same float32 points and the same
    cv2.getAffineTransform call."""
    import cv2
    c = np.array([w / 2., h / 2.], dtype=np.float32)
    s = max(h, w) * 1.0
    src = np.zeros((3, 2), dtype=np.float32)
    dst = np.zeros((3, 2), dtype=np.float32)
    src[0, :] = c
    src[1, :] = c + np.array([0, s * -0.5], dtype=np.float64)
    dst[0, :] = [out_w * 0.5, out_h * 0.5]
    dst[1, :] = np.array([out_w * 0.5, out_h * 0.5], np.float32) + np.array([0, out_w * -0.5], np.float32)
    third = lambda a, b: b + np.array([-(a - b)[1], (a - b)[0]], dtype=np.float32)  # noqa: E731  get_3rd_point
    src[2, :] = third(src[0, :], src[1, :])
    dst[2, :] = third(dst[0, :], dst[1, :])
    return cv2.getAffineTransform(np.float32(src), np.float32(dst))


def letterbox_u8(images, out_w=640, out_h=640):
    """The loader's letter-box of a cuda uint8 batch [B,h,w,3] of equally sized images to (out_h, out_w), on the device."""
    return warp_affine_u8(images, letterbox_matrix(images.shape[1], images.shape[2], out_w, out_h), out_w, out_h)


def ctdet_decode(heat, wh, reg=None, cat_spec_wh=False, K=100, return_inds=False):
    """Drop-in for centerface_ext.ctdet_decode (centerface_ext.py:52): cuda fp32 tensors heat [B,C,h,w] (post-sigmoid; the
    face model has C = 1), wh [B,2,h,w] (or [B,2C,h,w] with cat_spec_wh), reg [B,2,h,w] or None -> detections [B,K,6]."""
    import torch
    lib = L.load()
    assert heat.is_cuda and heat.dtype == torch.float32 and heat.dim() == 4
    B, cat, h, w = heat.shape
    assert tuple(wh.shape) == (B, 2 * cat if cat_spec_wh else 2, h, w), f"wh {tuple(wh.shape)} does not match heat {tuple(heat.shape)}"
    assert reg is None or tuple(reg.shape) == (B, 2, h, w)
    heat, wh = heat.contiguous(), wh.contiguous()
    reg = reg.contiguous() if reg is not None else None
    dets = torch.empty((B, K, 6), dtype=torch.float32, device=heat.device)
    inds = torch.empty((B, K), dtype=torch.int32, device=heat.device)
    scratch = torch.empty((B, cat, h, w), dtype=torch.float32, device=heat.device)
    with torch.cuda.device(heat.device):
        L.check(lib.cf_ctdet_decode_classes(C.c_void_p(heat.data_ptr()), C.c_void_p(wh.data_ptr()),
                                            C.c_void_p(reg.data_ptr()) if reg is not None else None, B, cat, h, w, K,
                                            1 if cat_spec_wh else 0, C.c_void_p(dets.data_ptr()), C.c_void_p(inds.data_ptr()),
                                            C.c_void_p(scratch.data_ptr()),
                                            C.c_void_p(torch.cuda.current_stream(heat.device).cuda_stream)), "cf_ctdet_decode_classes")
    return (dets, inds) if return_inds else dets


def decode_threshold(hm_sig, wh, reg, lm, variant, threshold, nms_threshold=0.3, size=(640, 640), scale_w=0.0,
                     scale_h=0.0, cap=1024):
    """Device-side paths A/B on cuda tensors hm_sig [B,1,h,w], wh/reg [B,2,h,w], lm [B,10,h,w]|None
    -> (dets [B,cap,5], lms [B,cap,10]|None, counts [B]) cuda tensors."""
    import torch
    lib = L.load()
    B, _, h, w = hm_sig.shape
    dev = hm_sig.device
    dets = torch.empty((B, cap, 5), dtype=torch.float32, device=dev)
    lms = torch.empty((B, cap, 10), dtype=torch.float32, device=dev) if (lm is not None and variant == L.CF_DECODE_A) else None
    counts = torch.empty((B,), dtype=torch.int32, device=dev)
    for t in (hm_sig, wh, reg, lm):
        assert t is None or (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32)
    p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None  # noqa: E731
    with torch.cuda.device(dev):
        L.check(lib.cf_decode_threshold(p(hm_sig), p(wh), p(reg), p(lm), B, h, w, variant, threshold, nms_threshold,
                                        size[0], size[1], scale_w, scale_h, cap, p(dets), p(lms), p(counts),
                                        C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "cf_decode_threshold")
    return dets, lms, counts


def inverse_affine(center, scale, output_size):
    """get_affine_transform(center, scale, 0, output_size, inv=1), utils/image.py:27-61, for rot = 0 and no shift:
    the same three float32 point pairs handed to the same cv2.getAffineTransform (host-side geometry, fp64 2x3)."""
    import cv2
    if not isinstance(scale, (np.ndarray, list)):
        scale = np.array([scale, scale], dtype=np.float32)
    center = np.asarray(center, dtype=np.float32)
    src_w, dst_w, dst_h = scale[0], output_size[0], output_size[1]
    src_dir = [0.0, src_w * -0.5]                      # get_dir([0, src_w * -0.5], 0)
    dst_dir = np.array([0, dst_w * -0.5], np.float32)
    src = np.zeros((3, 2), dtype=np.float32)
    dst = np.zeros((3, 2), dtype=np.float32)
    src[0, :] = center
    src[1, :] = center + src_dir
    dst[0, :] = [dst_w * 0.5, dst_h * 0.5]
    dst[1, :] = np.array([dst_w * 0.5, dst_h * 0.5], np.float32) + dst_dir

    def third(a, b):                                   # get_3rd_point, utils/image.py:70-72
        direct = a - b
        return b + np.array([-direct[1], direct[0]], dtype=np.float32)

    src[2:, :] = third(src[0, :], src[1, :])
    dst[2:, :] = third(dst[0, :], dst[1, :])
    return cv2.getAffineTransform(np.float32(dst), np.float32(src))


def ctdet_post_process(dets, c, s, h, w, num_classes=1):
    """Drop-in for utils.post_process.ctdet_post_process (utils/post_process.py:83-100), single face class: dets is a
    cuda fp32 tensor [B,K,6] (the output of ctdet_decode); returns the reference's 1-based class dict per image."""
    import torch
    assert num_classes == 1, "the face detector has one class"
    lib = L.load()
    assert dets.is_cuda and dets.dtype == torch.float32 and dets.is_contiguous() and dets.shape[2] == 6
    B, K, _ = dets.shape
    trans = np.stack([inverse_affine(c[i], s[i], (w, h)) for i in range(B)]).reshape(B, 6).astype(np.float64)
    t_dev = torch.from_numpy(trans).to(dets.device)
    out = torch.empty((B, K, 5), dtype=torch.float32, device=dets.device)
    with torch.cuda.device(dets.device):
        L.check(lib.cf_ctdet_post_process(C.c_void_p(dets.data_ptr()), C.c_void_p(t_dev.data_ptr()), B, K, C.c_void_p(out.data_ptr()),
                                          C.c_void_p(torch.cuda.current_stream(dets.device).cuda_stream)), "cf_ctdet_post_process")
    out = out.cpu().numpy()
    cls = dets[:, :, 5].cpu().numpy()
    return [{1: out[i][cls[i] == 0].tolist()} for i in range(B)]
