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
        # Python's min()/max() return one of their operands WITH ITS OWN dtype, so with mixed float32/float64 inputs the
        # reference's precision depends on which box wins each comparison: keep its scalar loop for that (unusual) case.
        for k in range(K):
            box_area = (query_boxes[k, 2] - query_boxes[k, 0] + 1) * (query_boxes[k, 3] - query_boxes[k, 1] + 1)
            for n in range(N):
                iw = min(boxes[n, 2], query_boxes[k, 2]) - max(boxes[n, 0], query_boxes[k, 0]) + 1
                if iw > 0:
                    ih = min(boxes[n, 3], query_boxes[k, 3]) - max(boxes[n, 1], query_boxes[k, 1]) + 1
                    if ih > 0:
                        ua = float((boxes[n, 2] - boxes[n, 0] + 1) * (boxes[n, 3] - boxes[n, 1] + 1) + box_area - iw * ih)
                        overlaps[n, k] = iw * ih / ua
        return overlaps
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
