"""Detection metrics of the validation path: host-side numpy, like the reference, which also computes
them on the CPU after the device work (ultralytics/utils/metrics.py:547-552 smooth, :605-634 compute_ap, :637-729
ap_per_class, :840-905 Metric / DetMetrics results).  The device part — IoU + matching — is ops.match_predictions."""
from __future__ import annotations

from typing import Dict

import numpy as np


def smooth(y: np.ndarray, f: float = 0.05) -> np.ndarray:
    """Box filter over a fraction f of the curve, edges padded with the end values (metrics.py:547-552)."""
    nf = round(len(y) * f * 2) // 2 + 1
    pad = np.ones(nf // 2)
    yp = np.concatenate((pad * y[0], y, pad * y[-1]), 0)
    return np.convolve(yp, np.ones(nf) / nf, mode="valid")


def compute_ap(recall: np.ndarray, precision: np.ndarray):
    """COCO 101-point interpolated AP of one precision/recall curve (metrics.py:605-634)."""
    mrec = np.concatenate(([0.0], recall, [1.0]))
    mpre = np.concatenate(([1.0], precision, [0.0]))
    mpre = np.flip(np.maximum.accumulate(np.flip(mpre)))          # precision envelope
    x = np.linspace(0, 1, 101)
    y = np.interp(x, mrec, mpre)
    ap = float(np.sum((y[1:] + y[:-1]) * 0.5 * np.diff(x)))       # trapezoid rule
    return ap, mpre, mrec


def ap_per_class(tp: np.ndarray, conf: np.ndarray, pred_cls: np.ndarray, target_cls: np.ndarray, eps: float = 1e-16):
    """Per-class AP at every IoU threshold plus P / R / F1 at the max-F1 confidence (metrics.py:637-729).

    tp [n, niou] bool, conf [n], pred_cls [n], target_cls [m].  Returns (tp_count, fp_count, p, r, f1, ap [nc, niou],
    unique_classes) — the plotting curves of the reference are not returned.
    """
    order = np.argsort(-conf)
    tp, conf, pred_cls = tp[order], conf[order], pred_cls[order]
    classes, n_labels = np.unique(target_cls, return_counts=True)
    nc = classes.shape[0]
    x = np.linspace(0, 1, 1000)
    ap = np.zeros((nc, tp.shape[1]))
    p_curve, r_curve = np.zeros((nc, 1000)), np.zeros((nc, 1000))
    for ci, c in enumerate(classes):
        sel = pred_cls == c
        n_l, n_p = n_labels[ci], int(sel.sum())
        if n_p == 0 or n_l == 0:
            continue
        fpc = (1 - tp[sel]).cumsum(0)
        tpc = tp[sel].cumsum(0)
        recall = tpc / (n_l + eps)
        r_curve[ci] = np.interp(-x, -conf[sel], recall[:, 0], left=0)       # conf decreases: interpolate on -conf
        precision = tpc / (tpc + fpc)
        p_curve[ci] = np.interp(-x, -conf[sel], precision[:, 0], left=1)
        for j in range(tp.shape[1]):
            ap[ci, j], _, _ = compute_ap(recall[:, j], precision[:, j])
    f1_curve = 2 * p_curve * r_curve / (p_curve + r_curve + eps)
    i = smooth(f1_curve.mean(0), 0.1).argmax() if nc else 0
    p, r, f1 = p_curve[:, i], r_curve[:, i], f1_curve[:, i]
    tp_count = (r * n_labels).round()
    fp_count = (tp_count / (p + eps) - tp_count).round()
    return tp_count, fp_count, p, r, f1, ap, classes.astype(int)


def results_dict(tp, conf, pred_cls, target_cls) -> Dict[str, float]:
    """The reference's `metrics.results_dict` keys for the detect task (metrics.py:860-905, DetMetrics.keys)."""
    keys = ["metrics/precision(B)", "metrics/recall(B)", "metrics/mAP50(B)", "metrics/mAP50-95(B)"]
    if len(tp) == 0 or not tp.any():
        return {k: 0.0 for k in keys} | {"fitness": 0.0}
    _, _, p, r, _, ap, _ = ap_per_class(tp, conf, pred_cls, target_cls)
    mp, mr, map50, map_ = float(p.mean()), float(r.mean()), float(ap[:, 0].mean()), float(ap.mean())
    return {keys[0]: mp, keys[1]: mr, keys[2]: map50, keys[3]: map_,
            "fitness": float(np.dot([0.0, 0.0, 0.1, 0.9], [mp, mr, map50, map_]))}


def box_iou(box1: np.ndarray, box2: np.ndarray, eps: float = 1e-7) -> np.ndarray:
    """[N,4] x [M,4] xyxy -> [N,M] IoU in float32, the operation order of ultralytics/utils/metrics.py:52-71."""
    b1, b2 = np.asarray(box1, np.float32), np.asarray(box2, np.float32)
    a1, a2 = b1[:, None, :2], b1[:, None, 2:]
    c1, c2 = b2[None, :, :2], b2[None, :, 2:]
    inter = np.clip(np.minimum(a2, c2) - np.maximum(a1, c1), 0, None).prod(2)
    return inter / ((a2 - a1).prod(2) + (c2 - c1).prod(2) - inter + np.float32(eps))


class ConfusionMatrix:
    """Detection confusion matrix (ultralytics/utils/metrics.py:394-493, without the plot): rows = predicted class, columns =
    true class, index nc = background.  Host-side numpy like the reference (a few hundred boxes per image)."""

    def __init__(self, nc: int, conf: float = 0.25, iou_thres: float = 0.45, task: str = "detect"):
        if task != "detect":
            raise NotImplementedError("specyolo implements the detect task only")
        self.task = task
        self.matrix = np.zeros((nc + 1, nc + 1))
        self.nc = nc
        self.conf = 0.25 if conf in {None, 0.001} else conf          # default val conf -> 0.25 (metrics.py:411)
        self.iou_thres = iou_thres

    def process_batch(self, detections, gt_bboxes, gt_cls):
        """detections [N,6] (x1,y1,x2,y2,conf,cls) | None, gt_bboxes [M,4] xyxy, gt_cls [M] (metrics.py:426-482)."""
        gt_cls = np.asarray(gt_cls).reshape(-1)
        if gt_cls.shape[0] == 0:
            if detections is not None:
                detections = np.asarray(detections, np.float32)
                for dc in detections[detections[:, 4] > self.conf][:, 5].astype(int):
                    self.matrix[dc, self.nc] += 1            # false positives
            return
        if detections is None:
            for gc in gt_cls.astype(int):
                self.matrix[self.nc, gc] += 1                # background FN
            return
        detections = np.asarray(detections, np.float32)
        detections = detections[detections[:, 4] > self.conf]
        gt_classes = gt_cls.astype(int)
        detection_classes = detections[:, 5].astype(int)
        iou = box_iou(np.asarray(gt_bboxes, np.float32).reshape(-1, 4), detections[:, :4])
        x = np.nonzero(iou > self.iou_thres)
        if x[0].shape[0]:
            matches = np.concatenate((np.stack(x, 1).astype(np.float32), iou[x[0], x[1]][:, None]), 1)
            if x[0].shape[0] > 1:
                matches = matches[matches[:, 2].argsort()[::-1]]
                matches = matches[np.unique(matches[:, 1], return_index=True)[1]]
                matches = matches[matches[:, 2].argsort()[::-1]]
                matches = matches[np.unique(matches[:, 0], return_index=True)[1]]
        else:
            matches = np.zeros((0, 3))
        n = matches.shape[0] > 0
        m0, m1, _ = matches.transpose().astype(int)
        for i, gc in enumerate(gt_classes):
            j = m0 == i
            if n and j.sum() == 1:
                self.matrix[detection_classes[m1[j]], gc] += 1   # correct
            else:
                self.matrix[self.nc, gc] += 1                    # true background
        for i, dc in enumerate(detection_classes):
            if not (m1 == i).any():
                self.matrix[dc, self.nc] += 1                    # predicted background

    def tp_fp(self):
        tp = self.matrix.diagonal()
        fp = self.matrix.sum(1) - tp
        return tp[:-1], fp[:-1]


class Metric:
    """Per-class and mean detection metrics (ultralytics/utils/metrics.py:726-873), plots / curves left out."""

    def __init__(self):
        self.p, self.r, self.f1, self.all_ap, self.ap_class_index = [], [], [], [], []
        self.nc = 0

    @property
    def ap50(self):
        return self.all_ap[:, 0] if len(self.all_ap) else []

    @property
    def ap(self):
        return self.all_ap.mean(1) if len(self.all_ap) else []

    @property
    def mp(self):
        return self.p.mean() if len(self.p) else 0.0

    @property
    def mr(self):
        return self.r.mean() if len(self.r) else 0.0

    @property
    def map50(self):
        return self.all_ap[:, 0].mean() if len(self.all_ap) else 0.0

    @property
    def map75(self):
        return self.all_ap[:, 5].mean() if len(self.all_ap) else 0.0

    @property
    def map(self):
        return self.all_ap.mean() if len(self.all_ap) else 0.0

    def mean_results(self):
        return [self.mp, self.mr, self.map50, self.map]

    def class_result(self, i):
        return self.p[i], self.r[i], self.ap50[i], self.ap[i]

    @property
    def maps(self):
        maps = np.zeros(self.nc) + self.map
        for i, c in enumerate(self.ap_class_index):
            maps[c] = self.ap[i]
        return maps

    def fitness(self):
        w = [0.0, 0.0, 0.1, 0.9]
        return (np.array(self.mean_results()) * w).sum()

    def update(self, results):
        self.p, self.r, self.f1, self.all_ap, self.ap_class_index = results


class DetMetrics:
    """`validator.metrics` of the reference (ultralytics/utils/metrics.py:898-965) without the plots."""

    def __init__(self, names=None):
        self.names = names or {}
        self.box = Metric()
        self.speed = {"preprocess": 0.0, "inference": 0.0, "loss": 0.0, "postprocess": 0.0}
        self.task = "detect"

    def process(self, tp, conf, pred_cls, target_cls):
        _, _, p, r, f1, ap, classes = ap_per_class(tp, conf, pred_cls, target_cls)
        self.box.nc = len(self.names)
        self.box.update((p, r, f1, ap, classes))

    @property
    def keys(self):
        return ["metrics/precision(B)", "metrics/recall(B)", "metrics/mAP50(B)", "metrics/mAP50-95(B)"]

    def mean_results(self):
        return self.box.mean_results()

    def class_result(self, i):
        return self.box.class_result(i)

    @property
    def maps(self):
        return self.box.maps

    @property
    def fitness(self):
        return self.box.fitness()

    @property
    def ap_class_index(self):
        return self.box.ap_class_index

    @property
    def results_dict(self):
        return dict(zip(self.keys + ["fitness"], self.mean_results() + [self.fitness]))
