"""WIDER-FACE txt writer, bbox_overlap and evaluate() against vectors generated from the reference
(oracle/gen_golden_widerface.py; eval_widerface.py:48-74, :172-211, demo.py:81-87).  CPU only."""
import importlib
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "widerface_v1.npz")


@pytest.fixture(scope="module")
def wf():
    return importlib.import_module("lightweight-face-detection-centernet_b200.widerface")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_bbox_overlap_is_bit_exact(wf, gold):
    for ci in gold["ov_cases"]:
        got = wf.bbox_overlap(gold[f"ov{ci}_boxes"], gold[f"ov{ci}_query"])
        want = gold[f"ov{ci}_out"]
        assert got.dtype == np.float64 and got.shape == want.shape
        assert np.array_equal(got, want), f"case {ci}: max diff {np.abs(got - want).max()}"
        if ci < 4:
            assert want[0, 0] == 1.0  # the identical pair (cases 4..: the pair differs by the float32 rounding of one side)
    assert wf.bbox_overlap(np.zeros((0, 4), np.float32), np.zeros((3, 4), np.float32)).shape == (0, 3)
    assert wf.bbox_overlap(np.zeros((2, 4), np.float32), np.zeros((0, 4), np.float32)).shape == (2, 0)


def test_bbox_overlap_edges(wf, gold):
    # ov0: query 1 starts one pixel past the last box (no overlap), query 2 shares exactly one column (overlap > 0)
    out = gold["ov0_out"]
    assert out[-1, 1] == 0.0 and out[-1, 2] > 0.0


def _val_data(gold):
    data = []
    for bi in range(3):
        dets, gts = [], []
        for j in range(4):
            d = gold[f"ev{bi}_{j}_det"]
            dets.append(None if d.shape == (0, 0) else d)
            gts.append(gold[f"ev{bi}_{j}_gt"])
        data.append({"meta": {"gt_det": gts}, "dets": dets})
    return data


@pytest.mark.parametrize("thr", [0.5, 0.3])
def test_evaluate_matches_reference(wf, gold, thr):
    r, p = wf.evaluate(_val_data(gold), None, threshold=thr, get_detections=lambda data, model: data["dets"])
    want = gold[f"eval_thr{thr}"]
    assert r == want[0] and p == want[1], ((r, p), want)


def test_txt_writer_matches_reference_bytes(wf, gold, tmp_path):
    path = wf.write_detections_txt(str(tmp_path), "0--Parade", "0_Parade_marchingband_1_465", gold["txt_dets"])
    assert path.endswith(os.path.join("0--Parade", "0_Parade_marchingband_1_465.txt"))
    assert open(path, "rb").read() == gold["txt_bytes"].tobytes()
    empty = wf.write_detections_txt(str(tmp_path), "1--X", "e", np.zeros((0, 5), np.float32))
    assert open(empty).read() == "1--X/e.jpg\n0\n"
