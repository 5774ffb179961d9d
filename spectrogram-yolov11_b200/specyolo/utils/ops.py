"""Drop-in for the box utilities of ultralytics/utils/ops.py that sit on the hot path."""
from __future__ import annotations

from typing import List, Optional

import torch

from .. import ops as _ops


# torchvision.ops.nms (the call at ultralytics/utils/ops.py:312) has two kernels that disagree on pairs whose IoU sits
# within one ulp of the threshold: the CPU kernel compares the fp32 IoU with the threshold as a DOUBLE, the CUDA kernel
# takes the threshold as a FLOAT (measured on the B200: profiles/parity_nms_vs_torchvision_cuda.json — boxes at exactly
# IoU 1/3 or 3/5 are suppressed by the CPU kernel and kept by the CUDA kernel).  "cpu" (default) is the semantics the
# golden vectors of the real reference pin (tests/golden/nms_cases.npz); "cuda" reproduces what the reference returns
# when it runs on a GPU.  Everything else (stable descending sort, fp32 IoU arithmetic, strict >) is common.
TORCHVISION_NMS_SEMANTICS = "cpu"


def _iou_threshold(iou_thres: float) -> float:
    if TORCHVISION_NMS_SEMANTICS == "cuda":
        return float(torch.tensor(iou_thres, dtype=torch.float32).item())   # (double)(float)thr: compare as floats
    return float(iou_thres)


def xywh2xyxy(x: torch.Tensor) -> torch.Tensor:
    """ultralytics/utils/ops.py:432-449 (plumbing; the kernels do this conversion themselves)."""
    assert x.shape[-1] == 4, f"input shape last dimension expected 4 but input shape is {x.shape}"
    y = torch.empty_like(x)
    xy, wh = x[..., :2], x[..., 2:] / 2
    y[..., :2] = xy - wh
    y[..., 2:] = xy + wh
    return y


def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False,
                        multi_label=False, labels=(), max_det=300, nc=0, max_time_img=0.05, max_nms=30000,
                        max_wh=7680, in_place=True, rotated=False, end2end=False,
                        return_idxs: bool = False) -> List[torch.Tensor]:
    """Same signature, argument meaning and error behaviour as ultralytics/utils/ops.py:181-332.

    One batched kernel launch replaces the per-image Python loop + torchvision.ops.nms; results are
    bit-identical (kept indices, class ids, boxes, scores) to the reference on the same `prediction`.
    Not implemented (outside the detection hot path, raise instead of falling back): apriori `labels`,
    `rotated` boxes, mask coefficients (nm > 0).  `in_place` is accepted and ignored: the input tensor
    is never modified.  There is no time limit (`max_time_img` is ignored): the kernel cannot hit it.
    """
    assert 0 <= conf_thres <= 1, f"Invalid Confidence threshold {conf_thres}, valid values are between 0.0 and 1.0"
    assert 0 <= iou_thres <= 1, f"Invalid IoU {iou_thres}, valid values are between 0.0 and 1.0"
    if isinstance(prediction, (list, tuple)):
        prediction = prediction[0]
    if prediction.shape[-1] == 6 or end2end:  # end-to-end model output (B, N, 6): plain filtering
        output = [pred[pred[:, 4] > conf_thres][:max_det] for pred in prediction]
        if classes is not None:
            cl = torch.tensor(classes, device=prediction.device)
            output = [pred[(pred[:, 5:6] == cl).any(1)] for pred in output]
        return output
    if rotated or (labels and any(len(l) for l in labels)):
        raise NotImplementedError("rotated boxes / apriori labels are outside the specyolo hot path")
    if not prediction.is_cuda:
        raise RuntimeError("specyolo.non_max_suppression needs a CUDA tensor; there is no CPU fallback")
    bs = prediction.shape[0]
    nc = nc or (prediction.shape[1] - 4)
    nm = prediction.shape[1] - nc - 4
    if nm != 0:
        raise NotImplementedError("mask coefficients (nm > 0) are outside the specyolo hot path")
    pred = prediction.to(torch.float32).contiguous()
    A = pred.shape[2]
    cls_t: Optional[torch.Tensor] = None
    if classes is not None:
        cls_t = torch.tensor(list(classes), device=pred.device, dtype=torch.int32)
    out, cnt, keep, _ = _ops.nms(prediction=pred, B=bs, nc=nc, A=A, conf_thres=conf_thres, iou_thres=_iou_threshold(iou_thres),
                                 agnostic=agnostic, multi_label=multi_label, max_det=max_det, max_nms=max_nms,
                                 max_wh=float(max_wh), classes=cls_t)
    counts = cnt.tolist()  # the one device->host sync of the whole batch
    result = [out[b, : counts[b]] for b in range(bs)]
    if return_idxs:
        return result, [keep[b, : counts[b]].to(torch.int64) for b in range(bs)]
    return result


def scale_boxes(img1_shape, boxes, img0_shape, ratio_pad=None, padding=True, xywh=False):
    """ultralytics/utils/ops.py:92-127 on a [n, >=4] tensor (plumbing version; the predictor uses the kernel)."""
    if ratio_pad is None:
        gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
        pad = (round((img1_shape[1] - img0_shape[1] * gain) / 2 - 0.1),
               round((img1_shape[0] - img0_shape[0] * gain) / 2 - 0.1))
    else:
        gain, pad = ratio_pad[0][0], ratio_pad[1]
    if padding:
        boxes[..., 0] -= pad[0]
        boxes[..., 1] -= pad[1]
        if not xywh:
            boxes[..., 2] -= pad[0]
            boxes[..., 3] -= pad[1]
    boxes[..., :4] /= gain
    return clip_boxes(boxes, img0_shape)


def clip_boxes(boxes, shape):
    """ultralytics/utils/ops.py:335-354."""
    boxes[..., 0] = boxes[..., 0].clamp(0, shape[1])
    boxes[..., 1] = boxes[..., 1].clamp(0, shape[0])
    boxes[..., 2] = boxes[..., 2].clamp(0, shape[1])
    boxes[..., 3] = boxes[..., 3].clamp(0, shape[0])
    return boxes
