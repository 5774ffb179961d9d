"""TEST INFRASTRUCTURE — numpy restatement of the reference's NMS path.

  * non_max_suppression  <- ultralytics/utils/ops.py:181-332
  * nms (greedy IoU suppression) <- torchvision.ops.nms CPU kernel (third-party, torchvision>=0.9 unpinned in
    pyproject.toml:73; installed 0.26.0).  Published algorithm: stable descending sort by score, then for
    each not-yet-suppressed box i suppress every later box j with  inter/(area_i+area_j-inter) > thr, all in
    the box dtype (fp32), the threshold being a double.  The reference tree restates the same algorithm in
    examples/YOLOv8-LibTorch-CPP-Inference/main.cc:81-136.

All arithmetic is done with numpy float32 scalars/arrays so that every rounding step matches the fp32
kernel (numpy never contracts a*b+c into an FMA).  Pinned against torchvision.ops.nms itself and the real
reference function by tests/test_oracle_vs_reference.py / tests/golden/nms_*.npz.

Only tests/, smoke() and bench.py's CPU-baseline legs may import this.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

f32 = np.float32


def xywh2xyxy(x: np.ndarray) -> np.ndarray:
    """ops.py:432-449 (fp32)."""
    x = x.astype(f32, copy=False)
    y = np.empty_like(x)
    xy = x[..., :2]
    wh = x[..., 2:] / f32(2)
    y[..., :2] = xy - wh
    y[..., 2:] = xy + wh
    return y


def nms(boxes: np.ndarray, scores: np.ndarray, iou_threshold: float) -> np.ndarray:
    """torchvision.ops.nms semantics: kept indices in descending-score order (ties: lower index first)."""
    boxes = boxes.astype(f32, copy=False)
    n = boxes.shape[0]
    if n == 0:
        return np.zeros((0,), dtype=np.int64)
    order = np.argsort(-scores.astype(f32), kind="stable")
    x1, y1, x2, y2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    areas = (x2 - x1) * (y2 - y1)  # fp32
    suppressed = np.zeros(n, dtype=bool)
    keep = []
    thr = float(iou_threshold)
    for _i in range(n):
        i = order[_i]
        if suppressed[i]:
            continue
        keep.append(i)
        rest = order[_i + 1:]
        if rest.size == 0:
            break
        xx1 = np.maximum(x1[i], x1[rest])
        yy1 = np.maximum(y1[i], y1[rest])
        xx2 = np.minimum(x2[i], x2[rest])
        yy2 = np.minimum(y2[i], y2[rest])
        w = np.maximum(f32(0), xx2 - xx1)
        h = np.maximum(f32(0), yy2 - yy1)
        inter = w * h
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = inter / (areas[i] + areas[rest] - inter)
        suppressed[rest[ovr.astype(np.float64) > thr]] = True
    return np.asarray(keep, dtype=np.int64)


def non_max_suppression(prediction: np.ndarray, conf_thres: float = 0.25, iou_thres: float = 0.45,
                        classes: Optional[Sequence[int]] = None, agnostic: bool = False, multi_label: bool = False,
                        max_det: int = 300, nc: int = 0, max_nms: int = 30000, max_wh: float = 7680,
                        return_indices: bool = False):
    """ops.py:181-332 for detection models (no masks, not rotated, not end2end).

    prediction: [B, 4+nc, A] fp32.  Returns list of [n,6] arrays (x1,y1,x2,y2,conf,cls); with
    return_indices also the kept indices into the post-threshold candidate list (`i` of ops.py:312-313).
    """
    assert 0 <= conf_thres <= 1, f"Invalid Confidence threshold {conf_thres}, valid values are between 0.0 and 1.0"
    assert 0 <= iou_thres <= 1, f"Invalid IoU {iou_thres}, valid values are between 0.0 and 1.0"
    prediction = np.asarray(prediction, dtype=f32)
    bs = prediction.shape[0]
    nc = nc or (prediction.shape[1] - 4)
    mi = 4 + nc
    conf_t = f32(conf_thres)
    xc = prediction[:, 4:mi].max(1) > conf_t                     # ops.py:250
    multi_label &= nc > 1                                         # ops.py:255
    pred = np.transpose(prediction, (0, 2, 1)).copy()             # ops.py:257
    pred[..., :4] = xywh2xyxy(pred[..., :4])                      # ops.py:259-260
    output: List[np.ndarray] = [np.zeros((0, 6), dtype=f32) for _ in range(bs)]
    indices: List[np.ndarray] = [np.zeros((0,), dtype=np.int64) for _ in range(bs)]
    for xi in range(bs):
        x = pred[xi][xc[xi]]                                      # ops.py:269
        if not x.shape[0]:
            continue
        box, cls = x[:, :4], x[:, 4:mi]
        if multi_label:                                           # ops.py:286-288
            i, j = np.nonzero(cls > conf_t)
            x = np.concatenate((box[i], x[i, 4 + j, None], j[:, None].astype(f32)), 1)
        else:                                                     # ops.py:289-291
            j = cls.argmax(1)                                     # first maximum on ties, like torch.max
            conf = cls[np.arange(cls.shape[0]), j]
            x = np.concatenate((box, conf[:, None], j[:, None].astype(f32)), 1)[conf > conf_t]
        if classes is not None:                                   # ops.py:294-295
            x = x[np.isin(x[:, 5].astype(np.int64), np.asarray(classes))]
        n = x.shape[0]
        if not n:
            continue
        if n > max_nms:                                           # ops.py:301-302
            x = x[np.argsort(-x[:, 4], kind="stable")[:max_nms]]
        c = x[:, 5:6] * f32(0 if agnostic else max_wh)            # ops.py:305
        boxes = x[:, :4] + c                                      # ops.py:311 (fp32 add)
        i = nms(boxes, x[:, 4], iou_thres)[:max_det]              # ops.py:312-313
        output[xi] = x[i]
        indices[xi] = i
    return (output, indices) if return_indices else output


def scale_boxes(img1_shape, boxes: np.ndarray, img0_shape) -> np.ndarray:
    """ops.py:92-127 + clip_boxes ops.py:335-354 (xyxy, padding=True)."""
    gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
    pad = (round((img1_shape[1] - img0_shape[1] * gain) / 2 - 0.1), round((img1_shape[0] - img0_shape[0] * gain) / 2 - 0.1))
    b = boxes.astype(f32).copy()
    b[..., 0] -= f32(pad[0])
    b[..., 1] -= f32(pad[1])
    b[..., 2] -= f32(pad[0])
    b[..., 3] -= f32(pad[1])
    b[..., :4] /= f32(gain)
    b[..., [0, 2]] = b[..., [0, 2]].clip(0, img0_shape[1])
    b[..., [1, 3]] = b[..., [1, 3]].clip(0, img0_shape[0])
    return b
