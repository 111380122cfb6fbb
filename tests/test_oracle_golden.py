"""The oracle (oracle/centerface_oracle.py) against the committed golden vectors, which were
produced by the reference itself (oracle/gen_golden.py asserts bit-equality at generation time).
Runs on CPU."""
import hashlib
import os

import numpy as np
import pytest
import torch

from conftest import IMGS


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_inputs_match_fixture_hashes(golden, f5_640):
    for n in IMGS:
        assert sha(f5_640[n]) == str(golden[f"f5_640/{n}/input_sha256"]), n


@pytest.mark.parametrize("n", IMGS)
def test_forward_heads(oracle, golden, oracle_heads_640, n):
    """EfficientNet.forward restated: hm within the oracle noise floor of the stored reference run
    (thread count may differ from generation time; SURVEY.md 8c noise floor 3.3e-6 on the logit)."""
    o = oracle_heads_640[n]
    assert np.abs(o["hm"][0].numpy() - golden[f"f5_640/{n}/hm"]).max() < 2e-5
    for k in ("wh", "lm", "reg"):
        assert abs(o[k].double().sum().item() - float(golden[f"f5_640/{n}/{k}_sum64"])) < 0.5
        assert abs(o[k].abs().max().item() - float(golden[f"f5_640/{n}/{k}_absmax"])) < 1e-3


@pytest.mark.parametrize("n", ["27", "17"])
def test_decode_paths_on_golden_heads(oracle, golden, n):
    """Decode restatements on the stored reference head maps: bit-exact."""
    g = lambda k: golden[f"f5_640/{n}/{k}"]  # noqa: E731
    hm = oracle.sigmoid_clamp(torch.from_numpy(g("hm"))[None])
    wh, lm, reg = (torch.from_numpy(g(k))[None] for k in ("wh", "lm", "reg"))
    da, la = oracle.decode_a(hm.numpy(), wh.numpy(), reg.numpy(), lm.numpy(), (640, 640))
    assert np.array_equal(np.asarray(da, np.float32), g("pathA_dets"))
    assert np.array_equal(np.asarray(la, np.float32), g("pathA_lms"))
    db = oracle.decode_b(hm.numpy()[0], wh.numpy()[0], reg.numpy()[0], (640, 640), 0.35)
    assert np.array_equal(np.asarray(db, np.float32), g("pathB_dets"))
    dc, ic = oracle.ctdet_decode(hm, wh, reg, K=100)
    assert np.array_equal(dc[0].numpy(), g("pathC_dets"))
    assert np.array_equal(ic[0].numpy().astype(np.int32), g("pathC_inds"))


def test_toy_path_b(oracle, golden):
    hm = np.full((1, 160, 160), 1e-4, np.float32)
    hm[0, 10, 20] = 0.9
    wh = np.full((2, 160, 160), 5.0, np.float32)
    rg = np.zeros((2, 160, 160), np.float32)
    rg[0], rg[1] = 0.25, 0.75
    out = oracle.decode_b(hm, wh, rg, (640, 640), 0.35)
    assert np.array_equal(out, golden["toy/pathB"])
    assert np.allclose(out, [[75, 33, 95, 53, 0.9]])


def test_detect_native_one_image(oracle, sd, golden, images):
    """CenterFace.__call__ restated end to end (resize, net, decode A, NMS, //scale) on 8.jpg."""
    n = "8"
    dets, lms = oracle.detect(sd, images[n])
    gd, gl = golden[f"native/{n}/dets"], golden[f"native/{n}/lms"]
    assert dets.shape == gd.shape
    # scores may wobble at the oracle noise floor between thread counts; coordinates are floor-divided
    assert np.abs(dets[:, 4] - gd[:, 4]).max() < 1e-5
    assert (np.abs(dets[:, :4] - gd[:, :4]) <= 1.0).all() and (dets[:, :4] == gd[:, :4]).mean() > 0.98
    assert (np.abs(lms - gl) <= 1.0).all()


def test_post_process(oracle, golden):
    """ctdet_post_process (utils/post_process.py:83-100) restated."""
    dets = golden["f5_640/27/pathC_dets"][None].copy()
    out = oracle.ctdet_post_process(dets, golden["post/27/c"], golden["post/27/s"], 160, 160)
    assert np.allclose(out[0], golden["post/27/dets"], rtol=0, atol=2e-3)


def test_topk_tie_rule(oracle):
    """(score desc, flat index asc) on ties -- the documented total order."""
    heat = torch.full((1, 1, 8, 8), 0.5)
    s, idx, cls, ys, xs = oracle.topk(heat, 5)
    assert idx[0].tolist() == [0, 1, 2, 3, 4] and cls[0].tolist() == [0] * 5
    # several classes: the lower class first, then the lower pixel (the class-major candidate order of centerface_ext.py:20)
    heat = torch.full((1, 3, 4, 4), 0.5)
    heat[0, 2, 1, 1] = 0.9
    s, idx, cls, ys, xs = oracle.topk(heat, 4)
    assert cls[0].tolist() == [2, 0, 0, 0] and idx[0].tolist() == [5, 0, 1, 2]


def test_ctdet_decode_classes_match_reference_golden(oracle):
    """C > 1 and cat_spec_wh (centerface_ext.py:11-27, :72-77) against outputs of the reference itself
    (oracle/gen_golden_multiclass.py -> tests/golden/multiclass_v1.npz): bit-exact."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "multiclass_v1.npz"))
    names = sorted({k.split("/")[0] for k in z.files})
    assert len(names) == 5
    for n in names:
        B, C, h, w, K, cat, with_reg = z[f"{n}/meta"].tolist()
        reg = torch.from_numpy(z[f"{n}/reg"]) if with_reg else None
        dets, inds = oracle.ctdet_decode(torch.from_numpy(z[f"{n}/heat"]), torch.from_numpy(z[f"{n}/wh"]), reg, K=K, cat_spec_wh=bool(cat))
        assert np.array_equal(dets.numpy(), z[f"{n}/dets"]), n
        assert np.array_equal(inds.numpy().astype(np.int32), z[f"{n}/inds"]), n


def test_peak_nms_plateau_and_border(oracle):
    heat = torch.zeros(1, 1, 4, 4)
    heat[0, 0, 0, 0] = 0.9   # border pixel competes only with in-bounds neighbours
    heat[0, 0, 2, 2] = 0.7
    heat[0, 0, 2, 3] = 0.7   # equal plateau neighbours are both kept
    out = oracle.peak_nms(heat)
    assert out[0, 0, 0, 0] == 0.9 and out[0, 0, 2, 2] == 0.7 and out[0, 0, 2, 3] == 0.7
    assert out[0, 0, 1, 1] == 0


def test_resize_restatement_is_bit_exact_against_cv2(oracle, images):
    """cv2.resize (INTER_LINEAR, 8UC3) restated: up-, down-, mixed scaling, the /32 sizes CenterFace.__call__ uses and
    the exact-2x case that OpenCV routes to INTER_AREA."""
    import cv2
    for n, img in images.items():
        h, w = img.shape[:2]
        sizes = [(640, 640), (480, 640), (int(np.ceil(h / 32) * 32), int(np.ceil(w / 32) * 32)), (320, 320), (h + 7, w + 13)]
        if h % 2 == 0 and w % 2 == 0:
            sizes.append((h // 2, w // 2))
        for dh, dw in sizes:
            assert np.array_equal(oracle.resize_linear_u8(img, dh, dw), cv2.resize(img, (dw, dh))), (n, (h, w), (dh, dw))


def _warp_cases(images):
    rng = np.random.RandomState(1)
    for name in ("27", "8", "1"):
        img = images[name]
        h, w = img.shape[:2]
        yield img, None, (640, 640)  # the loader's letter-box (matrix filled in by the caller)
        for _ in range(2):
            a, th = rng.uniform(0.3, 2.5), rng.uniform(-0.5, 0.5)
            M = np.array([[a * np.cos(th), -a * np.sin(th), rng.uniform(-200, 300)], [a * np.sin(th), a * np.cos(th), rng.uniform(-200, 300)]])
            yield img, M, (320, 256)


def test_warp_affine_restatement_is_bit_exact_against_cv2(oracle, images):
    """SURVEY.md 8f-1, letter-box variant (dataset/dataset.py:130-134): the numpy restatement of cv2.warpAffine(INTER_LINEAR,
    8UC3, constant border) equals OpenCV on the loader's own transform and on rotated / scaled / shifted ones."""
    import cv2
    n = 0
    for img, M, (dw, dh) in _warp_cases(images):
        if M is None:
            M = oracle.letterbox_matrix(img.shape[0], img.shape[1], dw, dh)
        want = cv2.warpAffine(img, M, (dw, dh), flags=cv2.INTER_LINEAR)
        assert np.array_equal(oracle.warp_affine_linear_u8(img, M, dw, dh), want)
        n += 1
    assert n == 9


def test_letterbox_matrix_equals_the_reference(oracle):
    """oracle.letterbox_matrix against get_affine_transform of the reference itself (utils/image.py:27-61), stored by
    oracle/gen_golden_widerface.py: every double equal."""
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "widerface_v1.npz"))
    for (h, w), M in zip(z["lb_sizes"], z["lb_mats"]):
        assert np.array_equal(oracle.letterbox_matrix(int(h), int(w), 640, 640), M), (h, w)


def test_letterbox_matrix_is_the_uniform_scale_about_the_centre(oracle):
    """get_affine_transform(c, max(h,w), 0, [W,H]) (utils/image.py:27-61): scale W/max(h,w), centre to centre."""
    for h, w in ((480, 640), (609, 1024), (898, 1600), (353, 490), (640, 480)):
        M = oracle.letterbox_matrix(h, w, 640, 640)
        a = 640.0 / max(h, w)
        want = np.array([[a, 0, 320 - a * w / 2], [0, a, 320 - a * h / 2]])
        assert np.allclose(M, want, rtol=0, atol=1e-4), (h, w, M, want)
