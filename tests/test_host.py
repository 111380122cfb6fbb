"""Host-side logic that needs no GPU: the weight packer's folds, the blob contract, and that the
C-ABI library loads and exports every symbol declared in include/centerface_b200.h."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import ROOT


def test_library_exports_every_declared_symbol(pkg):
    hdr = open(os.path.join(ROOT, "include", "centerface_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(cf_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 15
    lib = pkg._lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(pkg._lib.SIGNATURES), declared ^ set(pkg._lib.SIGNATURES)
    assert lib.cf_abi_version() == 1


def test_blob_contract(pkg, weights_path):
    sd = pkg.load_state_dict(weights_path)
    blob = pkg.pack_weights(sd)
    assert len(blob) == pkg._lib.load().cf_weights_blob_bytes()
    assert blob[:8] == b"CFB200W1"


def test_create_rejects_bad_arguments_without_gpu(pkg, weights_path):
    lib = pkg._lib.load()
    h = C.c_void_p()
    assert lib.cf_create(None, 0, 0, 1, 640, 640, 0, C.byref(h)) == -1
    assert b"NULL" in lib.cf_last_error()
    blob = pkg.pack_weights(pkg.load_state_dict(weights_path))
    buf = C.create_string_buffer(blob, len(blob))
    assert lib.cf_create(C.cast(buf, C.c_void_p), len(blob), 0, 1, 650, 640, 0, C.byref(h)) == -1  # not /32
    assert lib.cf_create(C.cast(buf, C.c_void_p), len(blob) - 4, 0, 1, 640, 640, 0, C.byref(h)) == -3
    bad = bytearray(blob)
    bad[0:8] = b"XXXXXXXX"
    buf2 = C.create_string_buffer(bytes(bad), len(bad))
    assert lib.cf_create(C.cast(buf2, C.c_void_p), len(bad), 0, 1, 640, 640, 0, C.byref(h)) == -3
    if not torch.cuda.is_available():  # the product must fail loudly without a device: no CPU fallback
        assert lib.cf_create(C.cast(buf, C.c_void_p), len(blob), 0, 1, 640, 640, 0, C.byref(h)) == -4
        with pytest.raises(pkg.CenterFaceError):
            pkg.Engine(weights_path)


def test_work_model_matches_survey(pkg):
    LW = pkg.CF_PW_TCGEN05_LAYERWISE
    by, fl = pkg._lib.work_model(640, 640, pkg.CF_IN_F32_NCHW, 0, LW)
    assert abs(fl - 4.695e9) / 4.695e9 < 1e-3          # 4.695 GFLOP / image
    _, fl480 = pkg._lib.work_model(480, 640, pkg.CF_IN_F32_NCHW, 0, LW)
    assert abs(fl480 - 3.521e9) / 3.521e9 < 1e-3
    parts = [pkg._lib.work_model(640, 640, pkg.CF_IN_F32_NCHW, c, LW) for c in (1, 2, 3, 4)]
    assert abs(sum(p[0] for p in parts) - by) < 1 and abs(sum(p[1] for p in parts) - fl) < 1
    # a fused MBConv block removes both hidden tensors' traffic, not the flops (default engine: layer1.0 at least)
    assert 1 in pkg._lib.fused_blocks(pkg.CF_PW_TCGEN05) and pkg._lib.fused_blocks(LW) == []
    byf, flf = pkg._lib.work_model(640, 640, pkg.CF_IN_F32_NCHW, 0, pkg.CF_PW_TCGEN05)
    assert abs(flf - fl) < 1 and byf < 0.80 * by
    partsf = [pkg._lib.work_model(640, 640, pkg.CF_IN_F32_NCHW, c, pkg.CF_PW_TCGEN05) for c in (1, 2, 3, 4, 6)]
    assert abs(sum(p[0] for p in partsf) - byf) < 1


def _entries(pkg, sd_np):
    from importlib import import_module
    w = import_module(pkg.__name__ + ".weights")
    return dict(w.entries(sd_np)), w


def test_head_collapse_and_bn_folds_match_oracle(pkg, oracle, sd, weights_path):
    """The packer's exact folds reproduce the reference graph on random inputs (fp32 tolerance)."""
    ents, w = _entries(pkg, pkg.load_state_dict(weights_path))
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 24, 12, 10, generator=g)
    # heads: conv3x3+b -> conv1x1+b for four heads == one 3x3 conv 24->15
    wc = torch.from_numpy(ents["heads.w"]).reshape(3, 3, 24, 16).permute(3, 2, 0, 1).contiguous()
    y = F.conv2d(x, wc, torch.from_numpy(ents["heads.b"]), 1, 1)
    o = 0
    for head, oc in oracle.HEADS:
        ref = F.conv2d(F.conv2d(x, sd[head + ".0.weight"], sd[head + ".0.bias"], 1, 1), sd[head + ".1.weight"], sd[head + ".1.bias"])
        assert (y[:, o:o + oc] - ref).abs().max() < 2e-5 * max(1.0, ref.abs().max().item()), head
        o += oc
    # IDAUp (model/centernet.py:186-204)
    for j, c in ((1, 96), (2, 32), (3, 24)):
        low = torch.randn(2, 24, 5, 6, generator=g) * 3
        skip = torch.randn(2, c, 10, 12, generator=g) * 3
        ref = oracle._idaup(sd, f"up{j}", low, skip)
        wl = torch.from_numpy(ents[f"up{j}.w"]).reshape(c, 24).t().reshape(24, c, 1, 1)
        b = F.relu(F.conv2d(skip, wl, torch.from_numpy(ents[f"up{j}.b"])))
        su = torch.from_numpy(ents[f"up{j}.su"]).reshape(24, 1, 2, 2)
        a = F.relu(F.conv_transpose2d(low, su, torch.from_numpy(ents[f"up{j}.tu"]), 2, 0, 0, 24))
        assert (a + b - ref).abs().max() < 1e-4 * max(1.0, ref.abs().max().item()), j
    # conv_last (conv + BN(1e-5) + Swish) on an input at the magnitude the real network produces
    x = torch.randn(1, 320, 4, 4, generator=g) * 1e12
    ref = oracle._swish(oracle._bn(sd, "conv_last.1", F.conv2d(x, sd["conv_last.0.weight"]), 1e-5))
    wl = torch.from_numpy(ents["clast.w"]).reshape(320, 24).t().reshape(24, 320, 1, 1)
    got = oracle._swish(F.conv2d(x, wl, torch.from_numpy(ents["clast.b"])))
    assert (got - ref).abs().max() < 1e-4 * max(1.0, ref.abs().max().item())


def test_normalise_lut_is_bit_exact(pkg, oracle):
    from importlib import import_module
    w = import_module(pkg.__name__ + ".weights")
    lut = w.normalise_lut()
    img = np.arange(256, dtype=np.uint8).reshape(16, 16, 1).repeat(3, axis=2)
    ref = oracle.normalize_u8(img)  # centerface.py:32-34
    for c in range(3):
        assert np.array_equal(lut[c][img[..., c]], ref[c])


def test_layouts(pkg, weights_path):
    sdn = pkg.load_state_dict(weights_path)
    ents, _ = _entries(pkg, sdn)
    w = sdn["first_conv.0.1.weight"]
    assert ents["stem.w"].reshape(3, 3, 3, 32)[1, 2, 0, 5] == w[5, 0, 1, 2]
    w = sdn["layer1.0.conv.0.1.weight"]  # expand 16->96
    assert ents["b1.exp"].reshape(16, 96)[3, 40] == w[40, 3, 0, 0]
    w = sdn["layer2.0.conv.1.1.weight"]  # dw 5x5, 144 ch
    assert ents["b3.dw"].reshape(25, 144)[7, 100] == w[100, 0, 1, 2]
    w = sdn["layer6.0.conv.2.weight"]  # project 960->320
    assert ents["b11.proj"].reshape(960, 320)[900, 300] == w[300, 900, 0, 0]


def test_transform_matches_reference_formula(oracle):
    assert oracle.transform(478, 720) == (480, 736, 480 / 478, 736 / 720)
    assert oracle.transform(640, 640) == (640, 640, 1.0, 1.0)


def test_inverse_affine_matches_oracle_closed_form(pkg, oracle):
    """The host-side affine of ctdet_post_process (cv2.getAffineTransform on the reference's three point pairs)
    against the oracle's closed form."""
    from importlib import import_module
    eng = import_module(pkg.__name__ + ".engine")
    for c, s, wh in (((512.0, 304.5), 1024.0, (160, 160)), ((320.0, 240.0), 640.0, (160, 120))):
        t = eng.inverse_affine(np.array(c, np.float32), s, wh)
        assert np.allclose(t, oracle.inverse_affine(c, s, wh[0], wh[1]), rtol=0, atol=1e-4)


def test_resize_tables_reproduce_cv2(pkg, images):
    """cf_resize_tables (the C++ restatement of OpenCV's index/weight loops) drives a numpy version of the device
    kernel's arithmetic; the result must equal cv2.resize bit for bit."""
    import cv2
    lib = pkg._lib.load()
    img = images["8"]
    sh, sw = img.shape[:2]
    for dh, dw in ((384, 512), (640, 640), (320, 320), (sh + 7, sw + 13)):
        tab = np.empty((3 * dw + 4 * dh,), np.int32)
        area2 = C.c_int32()
        assert lib.cf_resize_tables(sh, sw, dh, dw, C.c_void_p(tab.ctypes.data), tab.size, C.byref(area2)) == 0
        assert area2.value == 0
        sx, a0, a1 = tab[:dw], tab[dw:2 * dw], tab[2 * dw:3 * dw]
        y0, y1, b0, b1 = (tab[3 * dw + k * dh:3 * dw + (k + 1) * dh] for k in range(4))
        sx1 = np.minimum(sx + 1, sw - 1)
        S = img.astype(np.int32)
        h0 = S[y0][:, sx] * a0[None, :, None] + S[y0][:, sx1] * a1[None, :, None]
        h1 = S[y1][:, sx] * a0[None, :, None] + S[y1][:, sx1] * a1[None, :, None]
        out = ((((b0[:, None, None] * (h0 >> 4)) >> 16) + ((b1[:, None, None] * (h1 >> 4)) >> 16) + 2) >> 2).astype(np.uint8)
        assert np.array_equal(out, cv2.resize(img, (dw, dh))), (dh, dw)
    tab = np.empty((3 * 245 + 4 * 176,), np.int32)
    assert lib.cf_resize_tables(352, 490, 176, 245, C.c_void_p(tab.ctypes.data), tab.size, C.byref(area2)) == 0 and area2.value == 1
    assert lib.cf_resize_tables(352, 490, 176, 245, C.c_void_p(tab.ctypes.data), 10, C.byref(area2)) == -5


def test_warp_affine_tables_match_oracle(pkg, oracle):
    """cf_warp_affine_tables (host side of the device letter-box) against the oracle's restatement of OpenCV's fp64 ->
    fixed-point table construction, incl. a rotated matrix and a singular one (D = 0 -> all-zero inverse)."""
    lib = pkg._lib.load()
    mats = [oracle.letterbox_matrix(480, 640, 640, 640), oracle.letterbox_matrix(898, 1600, 640, 640),
            np.array([[0.7, -0.3, 12.5], [0.3, 0.7, -40.25]]), np.array([[1.0, 2.0, 3.0], [2.0, 4.0, 5.0]])]
    for M in mats:
        for dw, dh in ((640, 640), (320, 256)):
            tab = np.empty((2 * dw + 2 * dh,), np.int32)
            Mh = np.ascontiguousarray(np.asarray(M, np.float64).reshape(6))
            assert lib.cf_warp_affine_tables(C.c_void_p(Mh.ctypes.data), dh, dw, C.c_void_p(tab.ctypes.data), tab.size) == 0
            a, b, x0, y0 = oracle.warp_affine_tables(M, dw, dh)
            assert np.array_equal(tab, np.concatenate([a, b, x0, y0]))
    assert lib.cf_warp_affine_tables(C.c_void_p(Mh.ctypes.data), 8, 8, C.c_void_p(tab.ctypes.data), 10) == -5
    assert np.array_equal(pkg.letterbox_matrix(609, 1024, 640, 640), oracle.letterbox_matrix(609, 1024, 640, 640))


def test_bench_launch_table_matches_work_model(pkg):
    """bench.py's per-launch algorithmic bytes (roofline.top_launches) add up to cf_work_model's layer-wise totals: 42
    launches per step, the network's 384.6 MB per 640x640 image, and the same per-class sums."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    L = pkg._lib
    for h, w in ((640, 640), (480, 640), (320, 256)):
        tab = bench.launch_table(h, w)
        assert len(tab) == 42  # 41 network launches + the decode launch (peak keep fused into the top-k kernel at these sizes)
        net = sum(b for n, b in tab if "top-k" not in n and n != "peak mask")
        want, _ = L.work_model(h, w, L.CF_IN_U8_HWC, L.CLS_ALL, L.CF_PW_TCGEN05_LAYERWISE)
        assert net == want, (h, w, net, want)
        dw = sum(b for n, b in tab if " dw" in n)
        assert dw == L.work_model(h, w, L.CF_IN_U8_HWC, L.CLS_DW, L.CF_PW_TCGEN05_LAYERWISE)[0]
        pw = sum(b for n, b in tab if "expand" in n or "project" in n or n.startswith(("conv_last", "up")))
        assert pw == L.work_model(h, w, L.CF_IN_U8_HWC, L.CLS_PW, L.CF_PW_TCGEN05_LAYERWISE)[0]
        # the default engine: every fused block is one launch credited with the bytes of the three it replaces
        fused, dwp = L.fused_blocks(L.CF_PW_TCGEN05), L.dwp_blocks(L.CF_PW_TCGEN05)
        tabf = bench.launch_table(h, w, fused, dwp)
        assert len(tabf) == 42 - sum(2 if bench.BLOCKS[i][2] != 1 else 1 for i in fused) - len(dwp)
        assert sum(b for n, b in tabf) == sum(b for n, b in tab)
        # own=True: a fused launch counts its block input + output only -- the figure cf_work_model reports for the fused class
        tabo = bench.launch_table(h, w, fused, dwp, own=True)
        assert [n for n, _ in tabo] == [n for n, _ in tabf]
        assert sum(b for n, b in tabo if "fused" in n) == L.work_model(h, w, L.CF_IN_U8_HWC, L.CLS_FUSED, L.CF_PW_TCGEN05)[0]
        assert all(bo == bf for (n, bo), (_, bf) in zip(tabo, tabf) if "fused" not in n)
