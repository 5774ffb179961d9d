"""TEST INFRASTRUCTURE — numpy restatement of the reference's validation matching (never imported by the product).

box_iou                 ultralytics/utils/metrics.py:52-73
match_predictions       ultralytics/engine/validator.py:224-264 (use_scipy=False branch)

Pinned against the real reference functions by tests/golden/metrics.npz (oracle/gen_golden.py: gen_metrics).
The AP integration (ap_per_class) is host-side numpy in the product itself (specyolo/utils/metrics.py) and is pinned
by the same fixture file.
"""
from __future__ import annotations

import numpy as np


def box_iou(box1: np.ndarray, box2: np.ndarray, eps: float = 1e-7) -> np.ndarray:
    """[N,4] x [M,4] xyxy -> [N,M] IoU in fp32, operation order of metrics.py:68-73."""
    b1, b2 = box1.astype(np.float32), box2.astype(np.float32)
    a1, a2 = b1[:, None, :2], b1[:, None, 2:]
    c1, c2 = b2[None, :, :2], b2[None, :, 2:]
    inter = np.clip(np.minimum(a2, c2) - np.maximum(a1, c1), 0, None).prod(2)
    return (inter / ((a2 - a1).prod(2) + (c2 - c1).prod(2) - inter + np.float32(eps))).astype(np.float32)


def match_predictions(pred_classes: np.ndarray, true_classes: np.ndarray, iou: np.ndarray, iouv) -> np.ndarray:
    """iou [L labels, D detections] -> correct [D, len(iouv)] bool (validator.py:237-264)."""
    correct = np.zeros((pred_classes.shape[0], len(iouv)), dtype=bool)
    iou = iou * (true_classes[:, None] == pred_classes[None, :])
    for i, thr in enumerate(iouv):
        li, di = np.nonzero(iou >= thr)
        if li.shape[0] == 0:
            continue
        m = np.stack((li, di), 1)
        if m.shape[0] > 1:
            m = m[iou[m[:, 0], m[:, 1]].argsort()[::-1]]                 # by IoU, descending
            m = m[np.unique(m[:, 1], return_index=True)[1]]              # best label per detection
            m = m[np.unique(m[:, 0], return_index=True)[1]]              # first (lowest-index) detection per label
        correct[m[:, 1].astype(int), i] = True
    return correct
