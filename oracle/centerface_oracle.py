"""CPU oracle for the CenterFace inference hot path (TEST INFRASTRUCTURE ONLY).

This file is a CPU restatement of the reference's algorithm for the one path this
repo accelerates (image batch -> backbone + FPN + heads -> heat-map decode -> boxes).
It is the *checker*: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The
product (``lightweight-face-detection-centernet_b200``) never does.

Where the arithmetic lives: the reference is pure Python on top of PyTorch ATen
(``conv2d``, ``conv_transpose2d``, ``batch_norm``, ``sigmoid``, ``max_pool2d``,
``topk``, ``gather``; no version pinned by the reference, README says 1.0.1).  The
oracle therefore restates the *graph* with the same ATen calls on CPU fp32 and the
decode paths with the same numpy scalar arithmetic, function by function.

Pinning: the reference ships no tests or golden vectors, so
the oracle is pinned against outputs of the reference itself, imported from
/root/reference in the build container by ``oracle/gen_golden.py``; that script
asserts bit-equality oracle == reference on every fixture and commits the
vectors under ``tests/golden/``.

All ``file:line`` citations are into the reference repository.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

# centerface.py:12-15  (BGR order, applied after /255)
MEAN = np.array([0.408, 0.447, 0.470], dtype=np.float32).reshape(1, 1, 3)
STD = np.array([0.289, 0.274, 0.278], dtype=np.float32).reshape(1, 1, 3)

# model/centernet.py:211-221   (t, c, n, s, k)
MB_SETTINGS = [
    (1, 16, 1, 1, 3),
    (6, 24, 2, 2, 3),
    (6, 32, 2, 2, 5),
    (6, 64, 2, 2, 3),
    (6, 96, 2, 1, 5),
    (6, 160, 2, 2, 5),
    (6, 320, 1, 1, 3),
]
HEADS = (("hm", 1), ("wh", 2), ("lm", 10), ("reg", 2))  # model/centernet.py:240-245


# --------------------------------------------------------------------------------------
# weights
# --------------------------------------------------------------------------------------
def load_weights(path):
    """Load a state dict from the reference ``.pt`` (centerface.py:23) or from the
    ``.npz`` re-pack written by oracle/gen_golden.py.  Returns {name: fp32 tensor}."""
    path = str(path)
    if path.endswith(".npz"):
        z = np.load(path)
        return {k: torch.from_numpy(z[k]) for k in z.files}
    sd = torch.load(path, map_location="cpu", weights_only=True)
    return {k: v for k, v in sd.items()}


def random_weights(seed=0):
    """Random-init state dict of the reference architecture (for synthetic throughput
    runs where no checkpoint is at hand).  Scales are chosen so activations stay
    finite through the BN-free backbone; BN stats are benign."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(name, co, ci, k, gain=1.0):
        fan = ci * k * k
        sd[name] = torch.randn(co, ci, k, k, generator=g) * (gain / fan ** 0.5)

    def bn(prefix, c):
        sd[prefix + ".weight"] = 1.0 + 0.1 * torch.randn(c, generator=g)
        sd[prefix + ".bias"] = 0.1 * torch.randn(c, generator=g)
        sd[prefix + ".running_mean"] = 0.1 * torch.randn(c, generator=g)
        sd[prefix + ".running_var"] = 1.0 + 0.1 * torch.rand(c, generator=g)
        sd[prefix + ".num_batches_tracked"] = torch.tensor(0)

    conv("first_conv.0.1.weight", 32, 3, 3, 1.5)
    cin = 32
    for idx, (t, c, n, s, k) in enumerate(MB_SETTINGS):
        for i in range(n):
            hid = cin * t
            p = f"layer{idx}.{i}.conv."
            j = 0
            if t != 1:
                conv(p + "0.1.weight", hid, cin, 1, 1.5)
                j = 1
            sd[p + f"{j}.1.weight"] = torch.randn(hid, 1, k, k, generator=g) * (1.5 / k)
            conv(p + f"{j + 1}.weight", c, hid, 1, 1.0)
            cin = c
    conv("conv_last.0.weight", 24, 320, 1)
    bn("conv_last.1", 24)
    for name, ch in (("up1", 96), ("up2", 32), ("up3", 24)):
        sd[name + ".up.weight"] = 0.5 + 0.25 * torch.rand(24, 1, 2, 2, generator=g)
        bn(name + ".bn_up", 24)
        conv(name + ".conv.0.weight", 24, ch, 1)
        bn(name + ".conv.1", 24)
    for head, oc in HEADS:
        conv(head + ".0.weight", 24, 24, 3)
        sd[head + ".0.bias"] = 0.05 * torch.randn(24, generator=g)
        conv(head + ".1.weight", oc, 24, 1)
        sd[head + ".1.bias"] = 0.05 * torch.randn(oc, generator=g)
    sd["hm.1.bias"] = torch.full((1,), -1.79)  # model/centernet.py:258
    return sd


# --------------------------------------------------------------------------------------
# network  (model/centernet.py)
# --------------------------------------------------------------------------------------
def _same_pad(k, s):
    """ConvReLU._get_padding, model/centernet.py:68-70 -> [left, right, top, bottom]."""
    p = max(k - s, 0)
    return [p // 2, p - p // 2, p // 2, p - p // 2]


def _swish(x):
    """Swish.forward, model/centernet.py:39-40."""
    return x * torch.sigmoid(x)


def _conv_swish(x, w, k, s, groups=1):
    """ConvReLU (a mis-named ZeroPad2d -> Conv2d(bias=False) -> Swish), centernet.py:58-66."""
    x = F.pad(x, _same_pad(k, s))
    x = F.conv2d(x, w, None, s, 0, 1, groups)
    return _swish(x)


def _mbconv(sd, prefix, x, cin, cout, t, k, s):
    """MBConvBlock.forward with se=False, eval mode, model/centernet.py:89-140."""
    hid = cin * t
    y = x
    j = 0
    if cin != hid:  # :109
        y = _conv_swish(y, sd[prefix + "conv.0.1.weight"], 1, 1)
        j = 1
    y = _conv_swish(y, sd[prefix + f"conv.{j}.1.weight"], k, s, groups=hid)  # :112
    y = F.conv2d(y, sd[prefix + f"conv.{j + 1}.weight"])  # :118 linear projection
    if cin == cout and s == 1:  # :101, :135-137 (drop-connect is identity in eval)
        return x + y
    return y


def _bn(sd, prefix, x, eps):
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"],
                        sd[prefix + ".weight"], sd[prefix + ".bias"], False, 0.1, eps)


def _idaup(sd, name, x, skip):
    """IDAUp.forward, model/centernet.py:186-204 (BN eps 1e-3)."""
    up = F.conv_transpose2d(x, sd[name + ".up.weight"], None, 2, 0, 0, 24)
    a = F.relu(_bn(sd, name + ".bn_up", up, 1e-3))
    b = F.relu(_bn(sd, name + ".conv.1", F.conv2d(skip, sd[name + ".conv.0.weight"]), 1e-3))
    return a + b


def forward(sd, x, return_taps=False):
    """EfficientNet.forward, model/centernet.py:263-280.
    x: fp32 [B,3,H,W] (H,W multiples of 32) -> {'hm','wh','lm','reg'} raw head maps at H/4."""
    with torch.no_grad():
        taps = {}
        x = _conv_swish(x, sd["first_conv.0.1.weight"], 3, 2)  # :224
        taps["stem"] = x
        cin = 32
        feats = []
        for idx, (t, c, n, s, k) in enumerate(MB_SETTINGS):
            for i in range(n):
                x = _mbconv(sd, f"layer{idx}.{i}.", x, cin, c, t, k, s if i == 0 else 1)
                cin = c
            feats.append(x)
            taps[f"layer{idx}"] = x
        x1, x2, x4 = feats[1], feats[2], feats[4]
        x = feats[6]
        x = _swish(_bn(sd, "conv_last.1", F.conv2d(x, sd["conv_last.0.weight"]), 1e-5))  # :178-184
        taps["conv_last"] = x
        x = _idaup(sd, "up1", x, x4)
        x = _idaup(sd, "up2", x, x2)
        x = _idaup(sd, "up3", x, x1)
        taps["fpn"] = x
        out = {}
        for head, _ in HEADS:  # :249-256: conv3x3(+bias) -> conv1x1(+bias), no activation between
            y = F.conv2d(x, sd[head + ".0.weight"], sd[head + ".0.bias"], 1, 1)
            out[head] = F.conv2d(y, sd[head + ".1.weight"], sd[head + ".1.bias"])
    if return_taps:
        return out, taps
    return out


def sigmoid_clamp(hm):
    """centerface.py:43 / eval_widerface.py:85."""
    return torch.clamp(torch.sigmoid(hm), min=1e-4, max=1 - 1e-4)


# --------------------------------------------------------------------------------------
# pre-processing  (centerface.py:30-37, :68-71)
# --------------------------------------------------------------------------------------
def transform(h, w):
    """CenterFace.transform, centerface.py:68-71."""
    img_h_new, img_w_new = int(np.ceil(h / 32) * 32), int(np.ceil(w / 32) * 32)
    scale_h, scale_w = img_h_new / h, img_w_new / w
    return img_h_new, img_w_new, scale_h, scale_w


def normalize_u8(img_u8_hwc):
    """centerface.py:32-34 on an already-resized BGR u8 image -> fp32 CHW."""
    img = img_u8_hwc.astype(np.float32) / 255.0
    img = (img - MEAN) / STD
    return np.ascontiguousarray(img.transpose(2, 0, 1))


def resize_linear_u8(src, dh, dw):
    """cv2.resize(src, (dw, dh)) with the default INTER_LINEAR for 8-bit images, restated.

    The arithmetic lives in a third-party dependency the reference calls (centerface.py:30): OpenCV, not
    pinned by the reference (no requirements file); this image has opencv-python 4.13.0.  Algorithm
    (modules/imgproc/src/resize.cpp, resizeGeneric_ with HResizeLinear / VResizeLinear for uchar):
      * scale = 1 / (dst / src) in double;  f = float((d + 0.5) * scale - 0.5);  s = floor(f);  f -= s
      * x: s < 0 -> (s, f) = (0, 0);  s >= w-1 -> (w-1, 0).   y: f is kept, the two rows are clipped into the image
      * coefficients are rounded to 11-bit fixed point: cvRound(w * 2048) (round half to even)
      * horizontal pass in int32, vertical pass  (((b0 * (H0 >> 4)) >> 16) + ((b1 * (H1 >> 4)) >> 16) + 2) >> 2
      * an exact 2x decimation in both axes is silently switched to INTER_AREA: (a + b + c + d + 2) >> 2
    tests/test_oracle_golden.py pins this against cv2 itself on the bundled JPEGs (bit-exact)."""
    sh, sw, _ = src.shape
    if sh == 2 * dh and sw == 2 * dw:
        s = src.astype(np.int32)
        return ((s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2).astype(np.uint8)

    def frac(dn, sn):
        scale = 1.0 / (dn / sn)
        f = ((np.arange(dn) + 0.5) * scale - 0.5).astype(np.float32)
        s0 = np.floor(f).astype(np.int32)
        return s0, (f - s0).astype(np.float32)

    def fix(f):
        return np.rint((np.float32(1.0) - f) * np.float32(2048)).astype(np.int32), np.rint(f * np.float32(2048)).astype(np.int32)

    sx, fx = frac(dw, sw)
    lo, hi = sx < 0, sx >= sw - 1
    fx[lo | hi] = 0
    sx[lo] = 0
    sx[hi] = sw - 1
    ax0, ax1 = fix(fx)
    sx1 = np.minimum(sx + 1, sw - 1)
    sy0, fy = frac(dh, sh)
    by0, by1 = fix(fy)
    sy, sy1 = np.clip(sy0, 0, sh - 1), np.clip(sy0 + 1, 0, sh - 1)
    S = src.astype(np.int32)
    h0 = S[sy][:, sx] * ax0[None, :, None] + S[sy][:, sx1] * ax1[None, :, None]
    h1 = S[sy1][:, sx] * ax0[None, :, None] + S[sy1][:, sx1] * ax1[None, :, None]
    out = (((by0[:, None, None] * (h0 >> 4)) >> 16) + ((by1[:, None, None] * (h1 >> 4)) >> 16) + 2) >> 2
    return out.astype(np.uint8)


def warp_affine_tables(M, dw, dh):
    """The integer source-coordinate tables of cv2.warpAffine(src, M, (dw,dh), flags=INTER_LINEAR) (OpenCV 4.x
    imgwarp.cpp: the forward 2x3 matrix is inverted in fp64, then every destination pixel gets a source position in
    1/32-pixel fixed point: AB_BITS = 10, INTER_BITS = 5, round_delta = 16):
        X(x,y) = (X0[y] + adelta[x]) >> 5,   Y(x,y) = (Y0[y] + bdelta[x]) >> 5
    -> (adelta[dw], bdelta[dw], X0[dh], Y0[dh]) int32.  This is the letter-box step of the reference's loader,
    dataset/dataset.py:130-134."""
    M = np.asarray(M, dtype=np.float64).reshape(2, 3)
    D = M[0, 0] * M[1, 1] - M[0, 1] * M[1, 0]
    D = 1.0 / D if D != 0 else 0.0
    m0, m4 = M[1, 1] * D, M[0, 0] * D
    m1, m3 = M[0, 1] * (-D), M[1, 0] * (-D)
    b1 = -m0 * M[0, 2] - m1 * M[1, 2]
    b2 = -m3 * M[0, 2] - m4 * M[1, 2]
    x = np.arange(dw, dtype=np.float64)
    y = np.arange(dh, dtype=np.float64)
    rnd = lambda v: np.rint(v).astype(np.int64).clip(-2 ** 31, 2 ** 31 - 1).astype(np.int32)  # noqa: E731  saturate_cast<int> = cvRound
    adelta, bdelta = rnd(m0 * x * 1024.0), rnd(m3 * x * 1024.0)
    X0 = rnd((m1 * y + b1) * 1024.0) + 16
    Y0 = rnd((m4 * y + b2) * 1024.0) + 16
    return adelta, bdelta, X0, Y0


def warp_affine_linear_u8(src, M, dw, dh):
    """cv2.warpAffine(src, M, (dw, dh), flags=cv2.INTER_LINEAR) for 8UC3, BORDER_CONSTANT 0, bit for bit: the tables
    above, then remap's fixed-point bilinear kernel -- 15-bit weights 32*(32-fx | fx)*(32-fy | fy) (the (0,0) entry is
    {32767,0,0,1}: saturate_cast<short>(32768) plus the table's sum fix-up), out-of-image samples = 0,
    (sum + 2^14) >> 15."""
    src = np.ascontiguousarray(src, dtype=np.uint8)
    H, W = src.shape[:2]
    adelta, bdelta, X0, Y0 = warp_affine_tables(M, dw, dh)
    X = (X0[:, None].astype(np.int64) + adelta[None, :]) >> 5
    Y = (Y0[:, None].astype(np.int64) + bdelta[None, :]) >> 5
    sx = np.clip(X >> 5, -32768, 32767)
    sy = np.clip(Y >> 5, -32768, 32767)
    fx, fy = X & 31, Y & 31
    w00 = 32 * (32 - fx) * (32 - fy)
    w01 = 32 * fx * (32 - fy)
    w10 = 32 * (32 - fx) * fy
    w11 = 32 * fx * fy
    zero = (fx == 0) & (fy == 0)
    w00 = np.where(zero, 32767, w00)
    w11 = np.where(zero, 1, w11)
    pad = np.zeros((H + 2, W + 2, 3), dtype=np.int64)  # constant border 0; positions further out are all-zero anyway
    pad[1:-1, 1:-1] = src

    def at(yy, xx):
        ok = (yy >= -1) & (yy <= H) & (xx >= -1) & (xx <= W)
        v = pad[np.clip(yy + 1, 0, H + 1), np.clip(xx + 1, 0, W + 1)]
        return np.where(ok[..., None], v, 0)

    acc = (at(sy, sx) * w00[..., None] + at(sy, sx + 1) * w01[..., None] + at(sy + 1, sx) * w10[..., None] +
           at(sy + 1, sx + 1) * w11[..., None])
    return np.clip((acc + (1 << 14)) >> 15, 0, 255).astype(np.uint8)


def letterbox_matrix(h, w, out_w, out_h):
    """trans_input of dataset/dataset.py:113-131 for the 'val' split: get_affine_transform(c, s, 0, [out_w, out_h]) with
    c = (w/2, h/2) (float32), s = max(h, w): a uniform scale out_w/s about the centre (utils/image.py:27-61).  Returned
    as the forward 2x3 fp64 matrix cv2.getAffineTransform gives for those three point pairs."""
    import cv2
    c = np.array([w / 2., h / 2.], dtype=np.float32)
    s = max(h, w) * 1.0
    src = np.zeros((3, 2), dtype=np.float32)
    dst = np.zeros((3, 2), dtype=np.float32)
    src_dir = np.array([0, s * -0.5], dtype=np.float64)   # get_dir([0, -s/2], 0)
    dst_dir = np.array([0, out_w * -0.5], np.float32)
    src[0, :] = c
    src[1, :] = c + src_dir
    dst[0, :] = [out_w * 0.5, out_h * 0.5]
    dst[1, :] = np.array([out_w * 0.5, out_h * 0.5], np.float32) + dst_dir
    third = lambda a, b: b + np.array([-(a - b)[1], (a - b)[0]], dtype=np.float32)  # noqa: E731  get_3rd_point
    src[2, :] = third(src[0, :], src[1, :])
    dst[2, :] = third(dst[0, :], dst[1, :])
    return cv2.getAffineTransform(np.float32(src), np.float32(dst))


def preprocess(img_u8_hwc, h_new, w_new):
    """centerface.py:30-37: cv2 bilinear stretch to (w_new,h_new) then normalise -> [1,3,H,W]."""
    import cv2
    img = cv2.resize(img_u8_hwc, (w_new, h_new))
    return torch.from_numpy(normalize_u8(img)).unsqueeze(0)


# --------------------------------------------------------------------------------------
# decode, path A (centerface.py:73-151) and path B (eval_widerface.py:92-152)
# --------------------------------------------------------------------------------------
def nms(boxes, scores, nms_thresh):
    """Greedy IoU NMS with '+1' areas; centerface.py:111-151 == eval_widerface.py:112-152.
    Ordering: the reference uses ``np.argsort(scores)[::-1]`` whose order among *equal*
    scores is unspecified (unstable sort); the oracle fixes it as a stable ascending sort
    reversed, i.e. (score desc, original index desc).  All arithmetic is float32."""
    x1, y1, x2, y2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    areas = (x2 - x1 + 1) * (y2 - y1 + 1)
    order = np.argsort(scores, kind="stable")[::-1]
    n = boxes.shape[0]
    suppressed = np.zeros((n,), dtype=bool)
    keep = []
    for _i in range(n):
        i = order[_i]
        if suppressed[i]:
            continue
        keep.append(i)
        ix1, iy1, ix2, iy2, iarea = x1[i], y1[i], x2[i], y2[i], areas[i]
        for _j in range(_i + 1, n):
            j = order[_j]
            if suppressed[j]:
                continue
            xx1 = max(ix1, x1[j])
            yy1 = max(iy1, y1[j])
            xx2 = min(ix2, x2[j])
            yy2 = min(iy2, y2[j])
            w = max(0, xx2 - xx1 + 1)
            h = max(0, yy2 - yy1 + 1)
            inter = w * h
            ovr = inter / (iarea + areas[j] - inter)
            if ovr >= nms_thresh:
                suppressed[j] = True
    return keep


def decode_a(heatmap, scale, offset, landmark, size, threshold=0.1, landmarks=True):
    """CenterFace.decode, centerface.py:73-109.  heatmap [1,1,h,w] (post sigmoid+clamp),
    scale/offset [1,2,h,w], landmark [1,10,h,w].  NOTE: the ``threshold`` argument is
    ignored by the reference (0.3 hard-coded, :77) and offsets are read but unused."""
    heatmap = np.squeeze(heatmap)
    scale0, scale1 = scale[0, 0, :, :], scale[0, 1, :, :]
    c0, c1 = np.where(heatmap > 0.3)
    boxes, lms = [], []
    if len(c0) > 0:
        for i in range(len(c0)):
            s0, s1 = scale0[c0[i], c1[i]] * 4, scale1[c0[i], c1[i]] * 4
            s = heatmap[c0[i], c1[i]]
            x1, y1 = max(0, (c1[i] + 0.5) * 4 - s0 / 2), max(0, (c0[i] + 0.5) * 4 - s1 / 2)
            x1, y1 = min(x1, size[1]), min(y1, size[0])
            boxes.append([x1, y1, min(x1 + s0, size[1]), min(y1 + s1, size[0]), s])
            if landmarks:
                lm = []
                for j in range(5):
                    lm.append((landmark[0, j * 2, c0[i], c1[i]] + c1[i] + 0.5) * 4)
                    lm.append((landmark[0, j * 2 + 1, c0[i], c1[i]] + c0[i] + 0.5) * 4)
                lms.append(lm)
        boxes = np.asarray(boxes, dtype=np.float32)
        keep = nms(boxes[:, :4], boxes[:, 4], 0.3)
        boxes = boxes[keep, :]
        if landmarks:
            lms = np.asarray(lms, dtype=np.float32)
            lms = lms[keep, :]
    if landmarks:
        return boxes, lms
    return boxes


def decode_b(heatmap, scale, offset, size=(640, 640), threshold=0.35):
    """eval_widerface.decode, eval_widerface.py:92-110 (per image: heatmap [1,h,w],
    scale/offset [2,h,w]).  Offsets ARE used here, with the channels swapped (:102)."""
    heatmap = np.squeeze(heatmap)
    scale0, scale1 = scale[0, :, :], scale[1, :, :]
    offset0, offset1 = offset[0, :, :], offset[1, :, :]
    c0, c1 = np.where(heatmap > threshold)
    boxes = []
    if len(c0) > 0:
        for i in range(len(c0)):
            s0, s1 = scale0[c0[i], c1[i]] * 4, scale1[c0[i], c1[i]] * 4
            o0, o1 = offset0[c0[i], c1[i]], offset1[c0[i], c1[i]]
            s = heatmap[c0[i], c1[i]]
            x1, y1 = max(0, (c1[i] + o1 + 0.5) * 4 - s0 / 2), max(0, (c0[i] + o0 + 0.5) * 4 - s1 / 2)
            x1, y1 = min(x1, size[1]), min(y1, size[0])
            boxes.append([x1, y1, min(x1 + s0, size[1]), min(y1 + s1, size[0]), s])
        boxes = np.asarray(boxes, dtype=np.float32)
        keep = nms(boxes[:, :4], boxes[:, 4], 0.3)
        boxes = boxes[keep, :]
    return boxes


def rescale(dets, lms, scale_w, scale_h):
    """centerface.py:55-62: float32 floor-division back to source-image pixels."""
    if len(dets) > 0:
        dets = np.array(dets, dtype=np.float32, copy=True)
        dets[:, 0:4:2], dets[:, 1:4:2] = dets[:, 0:4:2] // scale_w, dets[:, 1:4:2] // scale_h
        if lms is not None:
            lms = np.array(lms, dtype=np.float32, copy=True)
            lms[:, 0:10:2], lms[:, 1:10:2] = lms[:, 0:10:2] // scale_w, lms[:, 1:10:2] // scale_h
    else:
        dets = np.empty(shape=[0, 5], dtype=np.float32)
        if lms is not None:
            lms = np.empty(shape=[0, 10], dtype=np.float32)
    return dets, lms


def detect(sd, img_u8_hwc, landmarks=True):
    """CenterFace.__call__, centerface.py:29-66, on CPU."""
    h, w = img_u8_hwc.shape[:2]
    hn, wn, sh, sw = transform(h, w)
    x = preprocess(img_u8_hwc, hn, wn)
    out = forward(sd, x)
    hm = sigmoid_clamp(out["hm"]).numpy()
    dets, lms = decode_a(hm, out["wh"].numpy(), out["reg"].numpy(), out["lm"].numpy(), (hn, wn))
    return rescale(dets, lms, sw, sh)


# --------------------------------------------------------------------------------------
# decode, path C  (centerface_ext.py:11-82)
# --------------------------------------------------------------------------------------
def peak_nms(heat):
    """_nms, centerface_ext.py:44-50: keep a pixel iff it equals its 3x3 max (-inf padding)."""
    hmax = F.max_pool2d(heat, (3, 3), stride=1, padding=1)
    keep = (hmax == heat).float()
    return heat * keep


def topk(scores, K):
    """_topk, centerface_ext.py:11-27.  torch.topk's order among equal scores is unspecified; the oracle fixes it as
    (score desc, flat index asc) with a stable sort, in both stages: per class over the H*W pixels (:14), then over the
    C*K class-major candidates (:20) -- so ties go to the lower class, then the lower pixel.
    Returns scores [B,K], inds [B,K] int64 (pixel index inside the class plane), classes [B,K] int32, ys, xs float32."""
    b, c, h, w = scores.shape
    s, idx = torch.sort(scores.reshape(b, c, -1), dim=2, descending=True, stable=True)
    s, idx = s[:, :, :K], idx[:, :, :K]     # :14 (already < h*w: the % of :16 is the identity)
    ys = (idx / w).int().float()            # :18 true division then truncation
    xs = (idx % w).int().float()            # :19
    s2, pick = torch.sort(s.reshape(b, -1), dim=1, descending=True, stable=True)
    s2, pick = s2[:, :K], pick[:, :K]       # :20
    clses = (pick / K).int()                # :21
    take = lambda t: t.reshape(b, -1).gather(1, pick)  # noqa: E731  :22-25
    return s2, take(idx), clses, take(ys), take(xs)


def ctdet_decode(heat, wh, reg=None, K=100, cat_spec_wh=False):
    """ctdet_decode, centerface_ext.py:52-82.  heat [B,C,h,w] post-sigmoid (the face model: C = 1), wh [B,2,h,w]
    (or [B,2C,h,w] with cat_spec_wh), reg [B,2,h,w] -> detections [B,K,6], inds [B,K]."""
    b, cat, h, w = heat.shape
    heat = peak_nms(heat)
    scores, inds, clses, ys, xs = topk(heat, K)

    def gather(feat):  # _transpose_and_gather_feat :37-42
        feat = feat.permute(0, 2, 3, 1).contiguous().view(b, -1, feat.size(1))
        return feat.gather(1, inds.unsqueeze(2).expand(b, K, feat.size(2)))

    if reg is not None:
        r = gather(reg)
        xs = xs.view(b, K, 1) + r[:, :, 0:1]
        ys = ys.view(b, K, 1) + r[:, :, 1:2]
    else:
        xs = xs.view(b, K, 1) + 0.5
        ys = ys.view(b, K, 1) + 0.5
    g = gather(wh)
    if cat_spec_wh:  # :72-75
        g = g.view(b, K, cat, 2).gather(2, clses.view(b, K, 1, 1).expand(b, K, 1, 2).long()).view(b, K, 2)
    bboxes = torch.cat([xs - g[..., 0:1] / 2, ys - g[..., 1:2] / 2,
                        xs + g[..., 0:1] / 2, ys + g[..., 1:2] / 2], dim=2)
    dets = torch.cat([bboxes, scores.view(b, K, 1), clses.view(b, K, 1).float()], dim=2)
    return dets, inds


# --------------------------------------------------------------------------------------
# path C follow-up: inverse affine to source coordinates
# (utils/post_process.py:83-100 + utils/image.py:19-66), fp64 like the reference
# --------------------------------------------------------------------------------------
def inverse_affine(center, scale, out_w, out_h):
    """get_affine_transform(center, scale, 0, (out_w,out_h), inv=1), utils/image.py:27-61, for
    rot=0 and shift=0.  The three point pairs are (c, c+(0,-s/2), third) <- (o, o+(0,-w/2), third);
    for rot=0 the map is a uniform scale s/out_w plus translation, solved in closed form."""
    s = float(scale)
    a = s / float(out_w)
    tx = float(center[0]) - a * (out_w * 0.5)
    ty = float(center[1]) - a * (out_h * 0.5) + (a * out_w * 0.5 - s * 0.5)
    return np.array([[a, 0.0, tx], [0.0, a, ty]], dtype=np.float64)


def ctdet_post_process(dets, c, s, h, w):
    """ctdet_post_process for num_classes=1: maps [B,K,6] output-map boxes to source pixels."""
    dets = np.array(dets, dtype=np.float64, copy=True)
    out = []
    for i in range(dets.shape[0]):
        t = inverse_affine(c[i], s[i], w, h)
        for lo in (0, 2):
            p = dets[i, :, lo:lo + 2].astype(np.float32)
            q = np.concatenate([p, np.ones((p.shape[0], 1), np.float32)], 1)  # affine_transform uses fp32 pt
            dets[i, :, lo:lo + 2] = q.astype(np.float64) @ t.T
        out.append(np.concatenate([dets[i, :, :4].astype(np.float32),
                                   dets[i, :, 4:5].astype(np.float32)], axis=1))
    return out


# --------------------------------------------------------------------------------------
# box metrics used by the parity tests
# --------------------------------------------------------------------------------------
def box_iou(a, b):
    """Plain IoU of matched rows a[i] vs b[i] (x1,y1,x2,y2)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    iw = np.maximum(np.minimum(a[:, 2], b[:, 2]) - np.maximum(a[:, 0], b[:, 0]), 0)
    ih = np.maximum(np.minimum(a[:, 3], b[:, 3]) - np.maximum(a[:, 1], b[:, 1]), 0)
    inter = iw * ih
    ua = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1]) + (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1]) - inter
    return np.where(ua > 0, inter / np.maximum(ua, 1e-30), 1.0)
