"""Weight packer: reference ``state_dict`` -> the kernel-ready fp32 blob ``cf_create`` takes.

Replaces the ``load_state_dict`` step of ``CenterFace.__init__`` (centerface.py:19-24).  All
re-layouts and the exact algebraic folds happen here, once, in float64, rounded to fp32 last:

* the 7 eval-mode BatchNorms (model/centernet.py:182,193,197) are folded into the preceding
  1x1 conv / depth-wise transposed conv;
* each head ``conv1x1(conv3x3(x)+b0)+b1`` has no non-linearity in between
  (model/centernet.py:249-256), so the four heads collapse into one 3x3 conv 24->15
  (``W' = W1.W0``, ``b' = W1.b0 + b1``; zero padding only touches x, so this is exact);
* point-wise weights become ``[K=Cin][N=Cout]`` row-major, depth-wise ``[k*k][C]``, the stem
  ``[(ky*3+kx)*3+ci][co]`` so that the channel axis of the NHWC activations is contiguous;
* a 3x256 table maps a raw BGR byte to the reference's normalised fp32 value with the
  reference's own float32 operations (centerface.py:32-34), making the fused u8 path bit-exact.

The blob layout (header | entry table | 128-byte aligned fp32 payload) is the contract checked
by ``cf_weights_blob_bytes`` / ``cf_create`` in csrc/engine.cu.
"""
from __future__ import annotations

import struct

import numpy as np

MAGIC = b"CFB200W1"
ENTRY_ALIGN_FLOATS = 32
# (cin, cout, t, k, s) of the 12 MBConv blocks, model/centernet.py:211-234
BLOCKS = [
    (32, 16, 1, 3, 1),
    (16, 24, 6, 3, 2), (24, 24, 6, 3, 1),
    (24, 32, 6, 5, 2), (32, 32, 6, 5, 1),
    (32, 64, 6, 3, 2), (64, 64, 6, 3, 1),
    (64, 96, 6, 5, 1), (96, 96, 6, 5, 1),
    (96, 160, 6, 5, 2), (160, 160, 6, 5, 1),
    (160, 320, 6, 3, 1),
]
BLOCK_NAMES = ["layer0.0", "layer1.0", "layer1.1", "layer2.0", "layer2.1", "layer3.0", "layer3.1",
               "layer4.0", "layer4.1", "layer5.0", "layer5.1", "layer6.0"]
HEADS = (("hm", 1), ("wh", 2), ("lm", 10), ("reg", 2))  # output channel order of the collapsed conv
MEAN = np.array([0.408, 0.447, 0.470], dtype=np.float32)  # centerface.py:12-15 (BGR)
STD = np.array([0.289, 0.274, 0.278], dtype=np.float32)


def _np(sd, key):
    v = sd[key]
    if hasattr(v, "detach"):
        v = v.detach().cpu().numpy()
    return np.asarray(v, dtype=np.float64)


def _bn_fold(sd, prefix, eps):
    """scale, shift of an eval-mode BatchNorm: y = x*scale + shift."""
    scale = _np(sd, prefix + ".weight") / np.sqrt(_np(sd, prefix + ".running_var") + eps)
    shift = _np(sd, prefix + ".bias") - _np(sd, prefix + ".running_mean") * scale
    return scale, shift


def normalise_lut():
    """[3][256] fp32: (v/255 - mean[c]) / std[c] with the reference's float32 ops."""
    v = np.arange(256, dtype=np.uint8).astype(np.float32) / 255.0
    return np.stack([(v - MEAN[c]) / STD[c] for c in range(3)]).astype(np.float32)


def entries(sd):
    """Ordered (name, fp32 array) list; names/counts mirror expected_entries() in engine.cu."""
    out = []
    w = _np(sd, "first_conv.0.1.weight")  # [32,3,3,3] (co,ci,ky,kx)
    out.append(("stem.w", w.transpose(2, 3, 1, 0).reshape(27, 32)))
    out.append(("lut", normalise_lut().reshape(-1)))
    for i, ((cin, cout, t, k, s), name) in enumerate(zip(BLOCKS, BLOCK_NAMES)):
        hid = cin * t
        p = name + ".conv."
        j = 0
        if t != 1:
            out.append((f"b{i}.exp", _np(sd, p + "0.1.weight").reshape(hid, cin).T))
            j = 1
        out.append((f"b{i}.dw", _np(sd, p + f"{j}.1.weight").reshape(hid, k * k).T))
        out.append((f"b{i}.proj", _np(sd, p + f"{j + 1}.weight").reshape(cout, hid).T))
    sc, sh = _bn_fold(sd, "conv_last.1", 1e-5)  # conv_1x1_bn, model/centernet.py:178-184
    out.append(("clast.w", (_np(sd, "conv_last.0.weight").reshape(24, 320) * sc[:, None]).T))
    out.append(("clast.b", sh))
    for j, c in ((1, 96), (2, 32), (3, 24)):  # IDAUp, model/centernet.py:186-204 (BN eps 1e-3)
        p = f"up{j}"
        sc, sh = _bn_fold(sd, p + ".conv.1", 1e-3)
        out.append((p + ".w", (_np(sd, p + ".conv.0.weight").reshape(24, c) * sc[:, None]).T))
        out.append((p + ".b", sh))
        su, tu = _bn_fold(sd, p + ".bn_up", 1e-3)
        out.append((p + ".su", _np(sd, p + ".up.weight").reshape(24, 4) * su[:, None]))
        out.append((p + ".tu", tu))
    wc = np.zeros((16, 24, 3, 3), np.float64)
    bc = np.zeros((16,), np.float64)
    o = 0
    for head, oc in HEADS:  # model/centernet.py:240-256
        w0, b0 = _np(sd, head + ".0.weight"), _np(sd, head + ".0.bias")
        w1, b1 = _np(sd, head + ".1.weight").reshape(oc, 24), _np(sd, head + ".1.bias")
        wc[o:o + oc] = np.einsum("om,mikl->oikl", w1, w0)
        bc[o:o + oc] = w1 @ b0 + b1
        o += oc
    out.append(("heads.w", wc.transpose(2, 3, 1, 0).reshape(216, 16)))
    out.append(("heads.b", bc))
    return [(n, np.ascontiguousarray(a, dtype=np.float64).astype(np.float32).reshape(-1)) for n, a in out]


def pack_entries(ents):
    def rup(a, b):
        return (a + b - 1) // b * b

    table_end = rup(32 + 56 * len(ents), 128)
    total = sum(rup(a.size, ENTRY_ALIGN_FLOATS) for _, a in ents)
    payload = np.zeros((total,), np.float32)
    table = b""
    off = 0
    for name, a in ents:
        nb = name.encode()
        assert len(nb) < 40
        payload[off:off + a.size] = a
        table += struct.pack("<40sQQ", nb, off, a.size)
        off += rup(a.size, ENTRY_ALIGN_FLOATS)
    header = struct.pack("<8sIIQQ", MAGIC, 1, len(ents), table_end, total)
    blob = header + table
    blob += b"\0" * (table_end - len(blob))
    return blob + payload.tobytes()


def pack_weights(sd):
    """state_dict (torch tensors or numpy arrays, reference key names) -> bytes."""
    return pack_entries(entries(sd))


def load_state_dict(path):
    """Reference checkpoint (.pt, centerface.py:23) or its .npz re-pack -> {name: ndarray}."""
    path = str(path)
    if path.endswith(".npz"):
        z = np.load(path)
        return {k: z[k] for k in z.files}
    import torch
    sd = torch.load(path, map_location="cpu", weights_only=True)
    return {k: v.numpy() for k, v in sd.items()}
