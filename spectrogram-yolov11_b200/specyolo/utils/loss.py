"""Detection criterion (first slice): host-side mirror of `v8DetectionLoss`
(ultralytics/utils/loss.py:163-275) over the CUDA criterion `specyolo_det_loss` (csrc/det_loss.cu) — assigner, class /
box / DFL losses and their gradients with respect to the head outputs in six launches, wrapped in a
`torch.autograd.Function` so that it drops into a PyTorch training step (`loss, items = criterion(preds, batch)`;
`loss.backward()` continues into whatever produced `preds`).  Same constructor, call signature and return values as the
reference: `(loss.sum() * batch_size, loss.detach())` with `loss = (box, cls, dfl)` after the `hyp` gains.

There is no CPU path: CPU tensors raise (use the reference package on the CPU).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Sequence, Tuple

import torch

from .. import _lib
from .._lib import DetLossArgs, check


def _hyp(h, name: str, default: float) -> float:
    if h is None:
        return default
    if isinstance(h, dict):
        return float(h.get(name, default))
    return float(getattr(h, name, default))


class _DetLossFn(torch.autograd.Function):
    """out = specyolo_det_loss(pred_distri, pred_scores, ...): returns (total, items[3], aux[2] = tss, positives)."""

    @staticmethod
    def forward(ctx, pred_distri, pred_scores, hw, strides, gt_boxes, gt_labels, gt_count, M, reg_max, tal, gains):
        if not pred_distri.is_cuda:
            raise RuntimeError("specyolo has no CPU criterion; use the reference package for CPU training")
        _lib.init_device()
        lib = _lib.load()
        pd = pred_distri.detach().float().contiguous()
        ps = pred_scores.detach().float().contiguous()
        B, A, nc = ps.shape
        nl = len(hw)
        need_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        a = DetLossArgs()
        a.nl = nl
        for i, ((h, w), s) in enumerate(zip(hw, strides)):
            a.h[i], a.w[i], a.stride[i] = int(h), int(w), float(s)
        assert sum(int(h) * int(w) for h, w in hw) == A and pd.shape == (B, A, 4 * reg_max)
        a.B, a.nc, a.reg_max = B, nc, reg_max
        a.pred_distri, a.pred_scores = pd.data_ptr(), ps.data_ptr()
        a.M = int(M)
        a.gt_boxes = gt_boxes.data_ptr() if M else None
        a.gt_labels = gt_labels.data_ptr() if M else None
        a.gt_count = gt_count.data_ptr()
        a.topk, a.alpha, a.beta, a.tal_eps = int(tal[0]), float(tal[1]), float(tal[2]), float(tal[3])
        a.gain_box, a.gain_cls, a.gain_dfl = (float(g) for g in gains)
        out = torch.empty(6, device=pd.device, dtype=torch.float32)
        gd = torch.empty_like(pd) if need_grad else None
        gs = torch.empty_like(ps) if need_grad else None
        hs = (C.c_int * nl)(*[int(h) for h, _ in hw])
        ws_ = (C.c_int * nl)(*[int(w) for _, w in hw])
        ws = torch.empty(lib.specyolo_det_loss_ws_bytes(B, hs, ws_, nl, int(M), int(tal[0])) + 256, device=pd.device, dtype=torch.uint8)
        off = (-ws.data_ptr()) % 256
        a.out, a.ws = out.data_ptr(), ws.data_ptr() + off
        a.grad_distri = gd.data_ptr() if need_grad else None
        a.grad_scores = gs.data_ptr() if need_grad else None
        check(lib.specyolo_det_loss(C.byref(a), _lib.stream_ptr()))
        ctx.save_for_backward(*( [gd, gs] if need_grad else [] ))
        ctx.in_dtypes = (pred_distri.dtype, pred_scores.dtype)
        total, items, aux = out[3].clone(), out[:3].clone(), out[4:].clone()
        ctx.mark_non_differentiable(items, aux)
        return total, items, aux

    @staticmethod
    def backward(ctx, g_total, _g_items, _g_aux):
        gd, gs = ctx.saved_tensors
        return ((gd * g_total).to(ctx.in_dtypes[0]), (gs * g_total).to(ctx.in_dtypes[1]),
                None, None, None, None, None, None, None, None, None)


def pack_targets(batch: Dict[str, torch.Tensor], B: int, img_hw: Tuple[float, float], device) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, int]:
    """loss.py:193-209 (`preprocess`) + `mask_gt` (:243) without the per-image Python loop: ground truth grouped by image,
    normalised xywh -> xyxy in input pixels, boxes whose coordinates sum to <= 0 dropped.  Returns gt_boxes [B, M, 4],
    gt_labels [B, M] (int32), gt_count [B] (int32) on `device`, and M."""
    bi = batch["batch_idx"].reshape(-1).long()
    cls = batch["cls"].reshape(-1)
    bb = batch["bboxes"].reshape(-1, 4).float()
    H, W = img_hw
    scale = torch.tensor([W, H, W, H], dtype=torch.float32, device=bb.device)
    xywh = bb * scale
    xyxy = torch.cat((xywh[:, :2] - xywh[:, 2:] / 2, xywh[:, :2] + xywh[:, 2:] / 2), 1)
    keep = (xyxy.sum(1) > 0) & (bi >= 0) & (bi < B)
    bi, cls, xyxy = bi[keep], cls[keep], xyxy[keep]
    count = torch.bincount(bi, minlength=B)
    M = int(count.max()) if bi.numel() else 0                      # the labels arrive on the host (the trainer moves only the images)
    gt_boxes = torch.zeros((B, max(M, 1), 4), dtype=torch.float32, device=bb.device)
    gt_labels = torch.zeros((B, max(M, 1)), dtype=torch.int32, device=bb.device)
    if M:
        order = torch.argsort(bi, stable=True)
        bs = bi[order]
        start = torch.cumsum(count, 0) - count
        slot = torch.arange(bs.numel(), device=bb.device) - start[bs]
        gt_boxes[bs, slot] = xyxy[order]
        gt_labels[bs, slot] = cls[order].to(torch.int32)
    return gt_boxes.to(device), gt_labels.to(device), count.to(torch.int32).to(device), M


def pack_batch_targets(batch: Dict[str, torch.Tensor], B: int, img_hw: Tuple[float, float], device, max_boxes: int = 0):
    """`pack_targets` ahead of the criterion call, with the per-image slot count padded up to `max_boxes`: put the result
    under `batch["packed_targets"]` and the criterion does no host work and no host synchronisation at all — which is
    what lets a whole training step (forward, criterion, backward, optimizer) be captured in ONE CUDA graph: the packed
    tensors are static buffers, refill them in place (`copy_`) between replays.  `max_boxes` must not be smaller than the
    largest number of boxes of any image that will ever be copied in."""
    gt_boxes, gt_labels, gt_count, M = pack_targets(batch, B, img_hw, device)
    if max_boxes > M:
        gb = torch.zeros((B, max_boxes, 4), dtype=torch.float32, device=gt_boxes.device)
        gl = torch.zeros((B, max_boxes), dtype=torch.int32, device=gt_boxes.device)
        gb[:, : gt_boxes.shape[1]] = gt_boxes
        gl[:, : gt_labels.shape[1]] = gt_labels
        gt_boxes, gt_labels, M = gb, gl, max_boxes
    return gt_boxes, gt_labels, gt_count, M


class v8DetectionLoss:
    """Criterion class for computing training losses (ultralytics/utils/loss.py:163-275)."""

    def __init__(self, model, tal_topk: int = 10):
        m = model.model[-1]                                         # Detect()
        self.hyp = getattr(model, "args", None)
        self.stride = m.stride
        self.nc = m.nc
        self.reg_max = m.reg_max
        self.no = m.nc + m.reg_max * 4
        self.device = next(model.parameters()).device
        self.tal = (tal_topk, 0.5, 6.0, 1e-9)                       # TaskAlignedAssigner(topk, num_classes, alpha=0.5, beta=6.0)

    def __call__(self, preds, batch):
        return criterion_call(self, preds, batch)


def criterion_call(self, preds, batch):
    """v8DetectionLoss.__call__ (loss.py:222-275) on the CUDA criterion.  `self` is this module's v8DetectionLoss or the
    REFERENCE's own instance (specyolo.ultralytics_shim): only attributes both have are read (stride, nc, reg_max, no, hyp,
    and the assigner's topk / alpha / beta / eps)."""
    feats = preds[1] if isinstance(preds, tuple) else preds
    B = feats[0].shape[0]
    cat = torch.cat([xi.view(B, self.no, -1) for xi in feats], 2)
    pred_distri = cat[:, : self.reg_max * 4].permute(0, 2, 1).contiguous()
    pred_scores = cat[:, self.reg_max * 4:].permute(0, 2, 1).contiguous()
    hw = [tuple(int(v) for v in f.shape[2:]) for f in feats]
    strides = getattr(self, "_strides_host", None)               # `stride` is a (device) tensor on the reference's criterion:
    if strides is None:                                          # read it once — float(tensor) is a host synchronisation
        strides = self._strides_host = [float(s) for s in self.stride]
    img_hw = (hw[0][0] * strides[0], hw[0][1] * strides[0])
    packed = batch.get("packed_targets") if isinstance(batch, dict) else None
    gt_boxes, gt_labels, gt_count, M = packed if packed is not None else pack_targets(batch, B, img_hw, pred_distri.device)
    gains = (_hyp(self.hyp, "box", 7.5), _hyp(self.hyp, "cls", 0.5), _hyp(self.hyp, "dfl", 1.5))
    asg = getattr(self, "assigner", None)
    tal = getattr(self, "tal", None) or (asg.topk, asg.alpha, asg.beta, asg.eps)
    total, items, aux = _DetLossFn.apply(pred_distri, pred_scores, hw, strides, gt_boxes, gt_labels, gt_count, M,
                                         self.reg_max, tal, gains)
    self.last_aux = aux                                          # (max(target_scores.sum(), 1), number of positives)
    return total, items
