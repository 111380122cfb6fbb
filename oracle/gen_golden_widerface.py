"""Golden vectors for the WIDER-FACE host helpers, generated FROM THE REFERENCE.

Runs only in the build container (needs /root/reference).  Imports the unmodified ``eval_widerface`` (bbox_overlap :48-74,
evaluate :172-211) and restates the txt writer of ``demo.py:81-87`` verbatim around the reference's format strings (demo.py
itself cannot be imported: it opens a window at import time), runs them on seeded synthetic boxes and writes
tests/golden/widerface_v1.npz.

    python oracle/gen_golden_widerface.py
"""
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
sys.path.insert(0, os.path.join(HERE, "refshim"))
sys.path.insert(0, REF)


def synth_boxes(rng, n, dtype, size=640.0):
    xy = rng.uniform(0, size - 40, size=(n, 2))
    wh = rng.uniform(4, 120, size=(n, 2))
    b = np.concatenate([xy, xy + wh], axis=1)
    return b.astype(dtype)


def main():
    cwd = os.getcwd()
    os.chdir(REF)
    import model.centernet as mc
    mc.ghost_net = mc.efficientnet_b0
    import eval_widerface as ref
    os.chdir(cwd)
    rng = np.random.RandomState(20260101)
    out = {}
    # ---- bbox_overlap: float32 detections against float32 / float64 annotations, with exact overlaps and touching edges
    cases = []
    for ci, (n, k, dt_b, dt_q) in enumerate([(7, 5, np.float32, np.float32), (12, 9, np.float32, np.float64), (3, 1, np.float64, np.float64),
                                             (1, 6, np.float32, np.float32)]):
        b = synth_boxes(rng, n, dt_b)
        q = synth_boxes(rng, k, dt_q)
        q[0] = b[0].astype(dt_q)                       # identical box
        if k > 1:
            q[1] = (b[-1] + np.array([b[-1][2] - b[-1][0] + 1, 0, b[-1][2] - b[-1][0] + 1, 0], dt_b)).astype(dt_q)  # just past the edge
        if k > 2:
            q[2] = (b[-1] + np.array([b[-1][2] - b[-1][0], 0, b[-1][2] - b[-1][0], 0], dt_b)).astype(dt_q)          # one shared column
        out[f"ov{ci}_boxes"] = b
        out[f"ov{ci}_query"] = q
        out[f"ov{ci}_out"] = ref.bbox_overlap(b, q)
        cases.append(ci)
    # mixed dtypes in both directions, on a half-pixel grid as well (ties in min() / max(): the reference keeps the FIRST operand's
    # dtype there); a separate generator, so that the cases above keep their bytes
    rng2 = np.random.RandomState(20261017)
    for j, (n, k, dt_b, dt_q, grid) in enumerate([(9, 7, np.float32, np.float64, 0), (9, 7, np.float64, np.float32, 0),
                                                   (11, 6, np.float32, np.float64, 2), (11, 6, np.float64, np.float32, 2),
                                                   (6, 8, np.float32, np.float64, 1), (6, 8, np.float64, np.float32, 1)]):
        ci = 4 + j
        b = synth_boxes(rng2, n, np.float64, size=120.0)
        q = synth_boxes(rng2, k, np.float64, size=120.0)
        q[0] = b[0]
        if grid:
            b, q = np.round(b * grid) / grid, np.round(q * grid) / grid
        b, q = b.astype(dt_b), q.astype(dt_q)
        out[f"ov{ci}_boxes"] = b
        out[f"ov{ci}_query"] = q
        out[f"ov{ci}_out"] = ref.bbox_overlap(b, q)
        cases.append(ci)
    out["ov_cases"] = np.array(cases)
    # ---- evaluate: three batches of four images, stored detections stand in for the network
    batches = []
    for bi in range(3):
        dets, annots = [], []
        for j in range(4):
            n_gt = [3, 0, 5, 2][j] if bi != 1 else [0, 4, 1, 6][j]
            gt = np.full((8, 5), -1.0, dtype=np.float32)  # the loader pads with -1 rows
            g = synth_boxes(rng, n_gt, np.float32)
            gt[:n_gt, :4] = g
            gt[:n_gt, 4] = 0
            n_det = [4, 2, 0, 3][(j + bi) % 4]
            d = synth_boxes(rng, n_det, np.float32)
            for t in range(min(n_det, n_gt)):  # some detections sit on a ground-truth box
                if t % 2 == 0:
                    d[t] = g[t] + rng.uniform(-2, 2, size=4).astype(np.float32)
            d = np.concatenate([d, rng.uniform(0.3, 1, size=(n_det, 1)).astype(np.float32)], axis=1)
            dets.append(d if n_det > 0 else (None if j % 2 == 0 else np.zeros((0, 5), np.float32)))
            annots.append(gt)
        batches.append((dets, annots))
    val_data = [{"meta": {"gt_det": a}, "dets": d} for d, a in batches]
    ref.get_detections = lambda data, model, cuda=True, threshold=0.35: data["dets"]  # evaluate() looks it up in its module
    ref.tqdm = lambda x: x
    for thr in (0.5, 0.3):
        r, p = ref.evaluate(val_data, None, threshold=thr)
        out[f"eval_thr{thr}"] = np.array([r, p], dtype=np.float64)
    for bi, (d, a) in enumerate(batches):
        for j in range(4):
            out[f"ev{bi}_{j}_gt"] = a[j]
            out[f"ev{bi}_{j}_det"] = d[j] if d[j] is not None else np.zeros((0, 0), np.float32)  # (0,0) encodes None
    # ---- txt writer: the reference's statements (demo.py:81-87)
    dets = np.concatenate([synth_boxes(rng, 5, np.float32), rng.uniform(0.05, 1, size=(5, 1)).astype(np.float32)], axis=1)
    im_dir, im_name = "0--Parade", "0_Parade_marchingband_1_465"
    f = io.StringIO()
    f.write('{:s}\n'.format('%s/%s.jpg' % (im_dir, im_name)))
    f.write('{:d}\n'.format(len(dets)))
    for b in dets:
        x1, y1, x2, y2, s = b
        f.write('{:.1f} {:.1f} {:.1f} {:.1f} {:.3f}\n'.format(x1, y1, (x2 - x1 + 1), (y2 - y1 + 1), s))
    out["txt_dets"] = dets
    out["txt_bytes"] = np.frombuffer(f.getvalue().encode(), dtype=np.uint8)
    # ---- the loader's letter-box matrix (dataset/dataset.py:113-131 -> utils/image.py:27-61), from the reference's own function
    from utils.image import get_affine_transform
    sizes = [(353, 490), (609, 1024), (898, 1600), (480, 640), (640, 480), (1080, 1920), (333, 500)]
    mats = []
    for h, w in sizes:
        c = np.array([w / 2., h / 2.], dtype=np.float32)
        mats.append(get_affine_transform(c, max(h, w) * 1.0, 0, [640, 640]))
    out["lb_sizes"] = np.array(sizes, dtype=np.int32)
    out["lb_mats"] = np.stack(mats)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "widerface_v1.npz"), **out)
    print("wrote tests/golden/widerface_v1.npz", {k: v.shape for k, v in out.items() if k.startswith(("ov0", "eval"))})


if __name__ == "__main__":
    main()
