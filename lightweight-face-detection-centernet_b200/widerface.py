"""WIDER-FACE result writer and recall/precision loop of the reference, host side.

    write_detections_txt   demo.py:81-87     one ``<event>/<image>.txt`` per image
    bbox_overlap           eval_widerface.py:48-74   inclusive-pixel IoU matrix
    evaluate               eval_widerface.py:172-211 mean per-batch recall / precision at an IoU threshold

The detections come from the CUDA path (``centerface.get_detections`` / ``CenterFace.__call__``); what is here is the
pure host arithmetic around them, vectorised with numpy in the operands' own dtypes so that every value equals the
reference's scalar loops bit for bit (pinned by tests/golden/widerface_v1.npz, generated from the reference itself).
"""
from __future__ import annotations

import os

import numpy as np


def write_detections_txt(save_path, im_dir, im_name, dets):
    """demo.py:81-87: header ``<im_dir>/<im_name>.jpg``, the count, then ``x y w h score`` (w, h inclusive: +1)."""
    os.makedirs(os.path.join(save_path, im_dir), exist_ok=True)
    path = os.path.join(save_path, im_dir, im_name + ".txt")
    with open(path, "w") as f:
        f.write("{:s}\n".format("%s/%s.jpg" % (im_dir, im_name)))
        f.write("{:d}\n".format(len(dets)))
        for b in dets:
            x1, y1, x2, y2, s = b
            f.write("{:.1f} {:.1f} {:.1f} {:.1f} {:.3f}\n".format(x1, y1, (x2 - x1 + 1), (y2 - y1 + 1), s))
    return path


def bbox_overlap(boxes, query_boxes):
    """eval_widerface.py:48-74 -> float64 [N,K].  The reference walks N x K Python scalars; the same expressions on whole
    arrays give the same bits when both inputs share a dtype (the loader and the decoders both produce float32)."""
    boxes = np.asarray(boxes)
    query_boxes = np.asarray(query_boxes)
    N, K = boxes.shape[0], query_boxes.shape[0]
    overlaps = np.zeros((N, K))
    if N == 0 or K == 0:
        return overlaps
    if boxes.dtype != query_boxes.dtype:
        return _bbox_overlap_mixed(boxes, query_boxes, overlaps)
    b = boxes[:, None, :]
    q = query_boxes[None, :, :]
    box_area = (q[..., 2] - q[..., 0] + 1) * (q[..., 3] - q[..., 1] + 1)            # :52-55
    iw = np.minimum(b[..., 2], q[..., 2]) - np.maximum(b[..., 0], q[..., 0]) + 1    # :57-60
    ih = np.minimum(b[..., 3], q[..., 3]) - np.maximum(b[..., 1], q[..., 1]) + 1    # :62-65
    ua = (b[..., 2] - b[..., 0] + 1) * (b[..., 3] - b[..., 1] + 1) + box_area - iw * ih  # :67-71 (float() is exact)
    hit = (iw > 0) & (ih > 0)
    with np.errstate(divide="ignore", invalid="ignore"):
        val = iw * ih / ua                                                          # :72
    overlaps[hit] = val[hit]
    return overlaps


def _bbox_overlap_mixed(boxes, query_boxes, overlaps):
    """Mixed input dtypes (float32 detections against float64 annotations, say).  The reference's Python min() / max() hand
    back one of their operands WITH ITS OWN dtype (the first one on a tie), so each difference `min - max + 1` is rounded in
    the detection dtype, the annotation dtype or their common type depending on which box wins each comparison, and the
    product iw * ih likewise.  Vectorised: every candidate expression is evaluated in its own dtype on whole arrays and the
    per-element winner pattern selects among them (values are exact when widened to float64, so the selection is lossless)."""
    tb, tq = boxes.dtype, query_boxes.dtype
    tr = np.result_type(tb, tq)
    b = boxes[:, None, :]
    q = query_boxes[None, :, :]
    one = 1  # a Python int: weak under NEP 50, keeps the operand dtype as in the reference's scalar arithmetic

    def extent(lo, hi):
        """min(b[hi], q[hi]) - max(b[lo], q[lo]) + 1 and the dtype class it was rounded in (0 detection, 1 annotation, 2 common)."""
        mn_b = ~(q[..., hi] < b[..., hi])   # min(a, b) returns a unless b < a
        mx_b = ~(q[..., lo] > b[..., lo])   # max(a, b) returns a unless b > a
        shape = np.broadcast(b[..., hi], q[..., hi]).shape
        bb = np.broadcast_to((b[..., hi] - b[..., lo] + one).astype(np.float64), shape)
        qq = np.broadcast_to((q[..., hi] - q[..., lo] + one).astype(np.float64), shape)
        bq = (b[..., hi].astype(tr) - q[..., lo].astype(tr) + one).astype(np.float64)
        qb = (q[..., hi].astype(tr) - b[..., lo].astype(tr) + one).astype(np.float64)
        val = np.where(mn_b, np.where(mx_b, bb, bq), np.where(mx_b, qb, qq))
        cls = np.where(mn_b & mx_b, 0, np.where(~mn_b & ~mx_b, 1, 2))
        return val, cls

    iw, cw = extent(0, 2)                                                           # :57-60
    ih, ch = extent(1, 3)                                                           # :62-65
    pc = np.where((cw == 0) & (ch == 0), 0, np.where((cw == 1) & (ch == 1), 1, 2))  # dtype class of iw * ih
    prod = np.where(pc == 0, (iw.astype(tb) * ih.astype(tb)).astype(np.float64),
                    np.where(pc == 1, (iw.astype(tq) * ih.astype(tq)).astype(np.float64),
                             (iw.astype(tr) * ih.astype(tr)).astype(np.float64)))
    # a product of two same-class extents stays in that class; any other pairing is computed in the common type -- as is every
    # later step, because the detection area (detection dtype) meets the annotation area (annotation dtype) in :67-71
    area_b = ((b[..., 2] - b[..., 0] + one) * (b[..., 3] - b[..., 1] + one)).astype(tr)
    area_q = ((q[..., 2] - q[..., 0] + one) * (q[..., 3] - q[..., 1] + one)).astype(tr)  # :52-55
    ua = (area_b + area_q - prod.astype(tr)).astype(np.float64)
    hit = (iw > 0) & (ih > 0)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        # :72 divides the numpy scalar iw * ih by a PYTHON float (ua went through float()): the quotient keeps the scalar's dtype
        val = np.where(pc == 0, (prod.astype(tb) / ua.astype(tb)).astype(np.float64),
                       np.where(pc == 1, (prod.astype(tq) / ua.astype(tq)).astype(np.float64), prod / ua))
    overlaps[hit] = val[hit]
    return overlaps


def evaluate(val_data, model, threshold=0.5, get_detections=None):
    """eval_widerface.py:172-211: (recall, precision) averaged over the batches of `val_data`.  Each element of
    `val_data` is the loader's dict (``'input'``, ``'meta': {'gt_det': [B] arrays [n,>=4], -1 rows = padding}``);
    `get_detections(data, model)` defaults to the CUDA path B of this package.  The reference's naming quirk is kept:
    its "recall" counts DETECTIONS whose best ground-truth overlap passes the threshold, divided by the ground-truth count."""
    if get_detections is None:
        from .centerface import get_detections
    recall = 0.
    precision = 0.
    n_batches = 0
    for data in val_data:
        n_batches += 1
        annots = data["meta"]["gt_det"]
        picked_boxes = get_detections(data, model)
        recall_iter = 0.
        precision_iter = 0.
        for j, boxes in enumerate(picked_boxes):
            annot_boxes = np.asarray(annots[j])
            annot_boxes = annot_boxes[annot_boxes[:, 0] != -1]
            if boxes is None and annot_boxes.shape[0] == 0:
                continue
            elif (boxes is None or len(boxes) < 1) and annot_boxes.shape[0] != 0:
                recall_iter += 0.
                precision_iter += 1.
                continue
            elif boxes is not None and annot_boxes.shape[0] == 0:
                recall_iter += 1.
                precision_iter += 0.
                continue
            overlap = bbox_overlap(boxes, annot_boxes).astype(np.float32)  # torch.FloatTensor(overlap), :197
            detected_num = int((overlap.max(axis=1) > threshold).sum())   # :198-200
            recall_iter += detected_num / annot_boxes.shape[0]
            true_positives = int((overlap.max(axis=0) > threshold).sum())  # :203-205
            precision_iter += true_positives / np.asarray(boxes).shape[0]
        recall += recall_iter / len(picked_boxes)
        precision += precision_iter / len(picked_boxes)
    return recall / n_batches, precision / n_batches
