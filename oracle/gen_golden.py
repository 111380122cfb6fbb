"""Pin the oracle against the reference itself and emit the golden fixtures.

Runs ONLY in the build container (needs /root/reference, which does not exist on the
GPU box).  It imports the unmodified reference with three shims:
  * stub ``mlconfig`` / ``torchsummary`` modules (oracle/refshim/),
  * a ``CenterFace`` subclass that restates ``__init__`` without ``.cuda()``
    (centerface.py:20-23 hard-codes it),
  * ``model.centernet.ghost_net = efficientnet_b0`` so ``centerface_ext`` imports and its
    module-level ``ctdet_decode`` (centerface_ext.py:52-82) becomes reachable,
runs it on the bundled JPEGs, asserts that oracle/centerface_oracle.py reproduces every
output BIT-EXACTLY, and writes tests/golden/*.

    python oracle/gen_golden.py            # regenerate + verify
"""
import hashlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, os.path.join(HERE, "refshim"))
sys.path.insert(0, REF)
sys.path.insert(0, HERE)

import cv2  # noqa: E402
import centerface_oracle as O  # noqa: E402

IMGS = ["1", "17", "2", "27", "8"]


def import_reference():
    os.chdir(REF)  # weights are cwd-relative (centerface.py:23)
    import model.centernet as mc
    mc.ghost_net = mc.efficientnet_b0  # F4 in SURVEY.md
    import centerface as ref_cf
    import eval_widerface as ref_eval
    try:
        import centerface_ext as ref_ext
    except Exception:  # its module body is fine; only class construction needs missing weights
        raise

    class CPUCenterFace(ref_cf.CenterFace):
        def __init__(self, height, width, landmarks=True):  # restates centerface.py:16-27 on CPU
            self.landmarks = landmarks
            self.net = mc.efficientnet_b0()
            self.cuda = False
            self.net.load_state_dict(torch.load("weight/model_epoch_100.pt", map_location="cpu",
                                                weights_only=True))
            self.net.eval()
            self.img_h_new, self.img_w_new, self.scale_h, self.scale_w = self.transform(height, width)

    return mc, ref_cf, ref_eval, ref_ext, CPUCenterFace


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    torch.manual_seed(0)
    mc, ref_cf, ref_eval, ref_ext, CPUCenterFace = import_reference()
    os.makedirs(GOLD, exist_ok=True)

    # ---- weights: re-pack the checkpoint as npz (same names, fp32) -------------------
    sd = torch.load(os.path.join(REF, "weight/model_epoch_100.pt"), map_location="cpu", weights_only=True)
    np.savez(os.path.join(GOLD, "weights_e100.npz"), **{k: v.numpy() for k, v in sd.items()})
    net = mc.efficientnet_b0()
    net.load_state_dict(sd)
    net.eval()
    sd_o = O.load_weights(os.path.join(GOLD, "weights_e100.npz"))

    # ---- the five JPEGs travel as byte fixtures --------------------------------------
    jpeg = {}
    for n in IMGS:
        with open(os.path.join(REF, "imgs", n + ".jpg"), "rb") as f:
            jpeg[n] = np.frombuffer(f.read(), dtype=np.uint8)
    np.savez(os.path.join(GOLD, "images_jpeg.npz"), **{"img_" + n: jpeg[n] for n in IMGS})
    bgr = {n: cv2.imdecode(jpeg[n], cv2.IMREAD_COLOR) for n in IMGS}
    for n in IMGS:
        assert np.array_equal(bgr[n], cv2.imread(os.path.join(REF, "imgs", n + ".jpg")))

    out = {}

    def ref_forward(x):
        with torch.no_grad():
            return net(x)[0]

    def check_heads(tag, x):
        r = ref_forward(x)
        o = O.forward(sd_o, x)
        for k in ("hm", "wh", "lm", "reg"):
            assert torch.equal(r[k], o[k]), (tag, k, (r[k] - o[k]).abs().max())
        return r

    # ---- set f5_640: stretched to 640x640 (SURVEY.md 8d parity inputs) ---------------
    for n in IMGS:
        u8 = cv2.resize(bgr[n], (640, 640))
        x = torch.from_numpy(O.normalize_u8(u8)).unsqueeze(0)
        # the reference's own normalisation (centerface.py:32-34) must equal the oracle's
        xr = ((u8.astype(np.float32) / 255.) - ref_cf.CenterFace.mean) / ref_cf.CenterFace.std
        assert np.array_equal(xr.transpose(2, 0, 1), x[0].numpy())
        r = check_heads("f5_640/" + n, x)
        out[f"f5_640/{n}/input_sha256"] = np.array(sha(u8))
        hm_s = torch.clamp(r["hm"].clone().sigmoid_(), min=1e-4, max=1 - 1e-4)
        assert torch.equal(hm_s, O.sigmoid_clamp(r["hm"]))
        out[f"f5_640/{n}/hm"] = r["hm"][0].numpy()
        if n in ("27", "17"):
            for k in ("wh", "lm", "reg"):
                out[f"f5_640/{n}/{k}"] = r[k][0].numpy()
        for k in ("wh", "lm", "reg"):  # checksums for the rest
            out[f"f5_640/{n}/{k}_sum64"] = np.array(r[k].double().sum().item())
            out[f"f5_640/{n}/{k}_absmax"] = np.array(r[k].abs().max().item())
        # path A at 640x640 through the reference decode (centerface.py:73-109)
        cf = CPUCenterFace.__new__(CPUCenterFace)
        cf.landmarks = True
        da, la = cf.decode(hm_s.numpy(), r["wh"].numpy(), r["reg"].numpy(), r["lm"].numpy(), (640, 640), threshold=0.2)
        oa, ola = O.decode_a(hm_s.numpy(), r["wh"].numpy(), r["reg"].numpy(), r["lm"].numpy(), (640, 640))
        assert np.array_equal(np.asarray(da), np.asarray(oa)) and np.array_equal(np.asarray(la), np.asarray(ola)), n
        out[f"f5_640/{n}/pathA_dets"] = np.asarray(da, np.float32).reshape(-1, 5)
        out[f"f5_640/{n}/pathA_lms"] = np.asarray(la, np.float32).reshape(-1, 10)
        # path B (eval_widerface.py:92-110), threshold 0.35
        db = ref_eval.decode(hm_s.numpy()[0], r["wh"].numpy()[0], r["reg"].numpy()[0], None, (640, 640), threshold=0.35)
        ob = O.decode_b(hm_s.numpy()[0], r["wh"].numpy()[0], r["reg"].numpy()[0], (640, 640), 0.35)
        assert np.array_equal(np.asarray(db), np.asarray(ob)), n
        out[f"f5_640/{n}/pathB_dets"] = np.asarray(db, np.float32).reshape(-1, 5)
        # path C (centerface_ext.py:52-82), K=100
        dc = ref_ext.ctdet_decode(hm_s, r["wh"], r["reg"], K=100)
        oc, oi = O.ctdet_decode(hm_s, r["wh"], r["reg"], K=100)
        assert torch.equal(dc, oc), (n, (dc - oc).abs().max())
        out[f"f5_640/{n}/pathC_dets"] = dc[0].numpy()
        out[f"f5_640/{n}/pathC_inds"] = oi[0].numpy().astype(np.int32)
        print("f5_640", n, "ok: pathA", len(da), "pathB", len(db), "hm max", float(hm_s.max()))

    # ---- set c4_vga: 640x480 frame centred on a zero 640x640 canvas, path B ----------
    for n in IMGS:
        canvas = np.zeros((640, 640, 3), np.uint8)
        canvas[80:560] = cv2.resize(bgr[n], (640, 480))
        x = torch.from_numpy(O.normalize_u8(canvas)).unsqueeze(0)
        r = check_heads("c4_vga/" + n, x)
        hm_s = O.sigmoid_clamp(r["hm"])
        # through the reference's own batched entry (eval_widerface.py:76-90)
        db = ref_eval.get_detections({"input": x}, net, cuda=False, threshold=0.35)[0]
        ob = O.decode_b(hm_s.numpy()[0], r["wh"].numpy()[0], r["reg"].numpy()[0], (640, 640), 0.35)
        assert np.array_equal(np.asarray(db), np.asarray(ob)), n
        out[f"c4_vga/{n}/input_sha256"] = np.array(sha(canvas))
        out[f"c4_vga/{n}/hm"] = r["hm"][0].numpy()
        out[f"c4_vga/{n}/pathB_dets"] = np.asarray(db, np.float32).reshape(-1, 5)
        print("c4_vga", n, "ok: pathB", len(db))

    # ---- set c5_320: max side 320, CenterFace.__call__ stretch to /32 sizes ----------
    # ---- set native: CenterFace.__call__ at the JPEG's own size ----------------------
    for tag in ("c5_320", "native"):
        for n in IMGS:
            img = bgr[n]
            if tag == "c5_320":
                h, w = img.shape[:2]
                f = 320.0 / max(h, w)
                img = cv2.resize(img, (int(round(w * f)), int(round(h * f))))
            h, w = img.shape[:2]
            cf = CPUCenterFace(h, w)
            stdout = sys.stdout
            sys.stdout = io.StringIO()  # the reference prints "cpu times = ..."
            try:
                dets, lms = cf(img, threshold=0.2)
            finally:
                sys.stdout = stdout
            od, ol = O.detect(sd_o, img)
            assert np.array_equal(dets, od) and np.array_equal(lms, ol), (tag, n)
            out[f"{tag}/{n}/hw"] = np.array([h, w], np.int32)
            out[f"{tag}/{n}/input_sha256"] = np.array(sha(img))
            out[f"{tag}/{n}/dets"] = np.asarray(dets, np.float32).reshape(-1, 5)
            out[f"{tag}/{n}/lms"] = np.asarray(lms, np.float32).reshape(-1, 10)
            print(tag, n, (h, w), "ok:", len(dets), "dets")

    # ---- path C follow-up: ctdet_post_process (utils/post_process.py:83-100) ---------
    from utils.post_process import ctdet_post_process as ref_pp
    dets = out["f5_640/27/pathC_dets"][None].copy()
    c = np.array([[512.0, 304.5]], np.float32)
    s = np.array([1024.0], np.float32)
    rp = ref_pp(dets.copy(), c, s, 160, 160, 1)
    op = O.ctdet_post_process(dets.copy(), c, s, 160, 160)
    rp0 = np.asarray(rp[0][1], np.float32)
    assert np.allclose(rp0, op[0], rtol=0, atol=2e-3), np.abs(rp0 - op[0]).max()
    out["post/27/c"] = c
    out["post/27/s"] = s
    out["post/27/dets"] = rp0
    print("ctdet_post_process ok, max diff", np.abs(rp0 - op[0]).max())

    # ---- toy anchors from SURVEY.md 8c ------------------------------------------------
    hm = np.full((1, 160, 160), 1e-4, np.float32)
    hm[0, 10, 20] = 0.9
    wh = np.full((2, 160, 160), 5.0, np.float32)
    rg = np.zeros((2, 160, 160), np.float32)
    rg[0], rg[1] = 0.25, 0.75
    tb = ref_eval.decode(hm, wh, rg, None, (640, 640), threshold=0.35)
    assert np.array_equal(tb, O.decode_b(hm, wh, rg, (640, 640), 0.35))
    assert np.allclose(tb, [[75, 33, 95, 53, 0.9]]), tb
    out["toy/pathB"] = np.asarray(tb, np.float32)

    np.savez_compressed(os.path.join(GOLD, "golden_v1.npz"), **out)
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main()
