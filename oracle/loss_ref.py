"""TEST INFRASTRUCTURE — CPU restatement of the reference's detection criterion (first slice):

    v8DetectionLoss.__call__          ultralytics/utils/loss.py:222-275
    BboxLoss / DFLoss                 ultralytics/utils/loss.py:66-129
    TaskAlignedAssigner               ultralytics/utils/tal.py:14-297   (topk 10, alpha 0.5, beta 6.0)
    bbox_iou(CIoU=True)               ultralytics/utils/metrics.py:171-228
    make_anchors / dist2bbox / bbox2dist   ultralytics/utils/tal.py:334-364

written image by image with explicit loops over the ground-truth boxes (the reference broadcasts [B, n_max, A] tensors).
Pinned by tests/golden/loss_cases.npz = outputs of the REAL criterion (losses, autograd gradients with respect to the
head maps, assigner targets), `oracle/gen_golden.py loss`.  Only tests/ may import this module.

Remark on ties (tal.py:143-169): `torch.topk` over the alignment metric may return anchors whose metric is exactly 0
(fewer than `topk` anchors with positive metric inside a box).  Such an anchor becomes "foreground" with target score 0:
it contributes nothing to any loss term or gradient, and which zero-metric anchors are returned is implementation
defined.  Comparisons therefore look at the losses, the gradients and the targets where the target score is positive.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F


def make_anchors(hw: Sequence[Tuple[int, int]], strides: Sequence[float], offset: float = 0.5):
    """tal.py:334-346 -> anchor centres [A, 2] in grid units (x, y) and the stride of every anchor [A, 1]."""
    pts, st = [], []
    for (h, w), s in zip(hw, strides):
        sx = torch.arange(w, dtype=torch.float32) + offset
        sy = torch.arange(h, dtype=torch.float32) + offset
        yy, xx = torch.meshgrid(sy, sx, indexing="ij")
        pts.append(torch.stack((xx, yy), -1).view(-1, 2))
        st.append(torch.full((h * w, 1), float(s)))
    return torch.cat(pts), torch.cat(st)


def ciou(b1: torch.Tensor, b2: torch.Tensor, eps: float = 1e-7) -> torch.Tensor:
    """metrics.py:199-228 with xywh=False, CIoU=True; boxes [..., 4] xyxy, result [...]."""
    x11, y11, x12, y12 = b1.unbind(-1)
    x21, y21, x22, y22 = b2.unbind(-1)
    w1, h1 = x12 - x11, y12 - y11 + eps
    w2, h2 = x22 - x21, y22 - y21 + eps
    inter = (torch.minimum(x12, x22) - torch.maximum(x11, x21)).clamp(min=0) * \
            (torch.minimum(y12, y22) - torch.maximum(y11, y21)).clamp(min=0)
    union = w1 * h1 + w2 * h2 - inter + eps
    iou = inter / union
    cw = torch.maximum(x12, x22) - torch.minimum(x11, x21)
    ch = torch.maximum(y12, y22) - torch.minimum(y11, y21)
    c2 = cw.pow(2) + ch.pow(2) + eps
    rho2 = ((x21 + x22 - x11 - x12).pow(2) + (y21 + y22 - y11 - y12).pow(2)) / 4
    v = (4 / math.pi ** 2) * (torch.atan(w2 / h2) - torch.atan(w1 / h1)).pow(2)
    with torch.no_grad():
        alpha = v / (v - iou + (1 + eps))
    return iou - (rho2 / c2 + v * alpha)


def decode_boxes(pred_distri: torch.Tensor, anchors: torch.Tensor, reg_max: int = 16) -> torch.Tensor:
    """loss.py:211-220: softmax expectation over the reg_max bins of each side, dist2bbox -> xyxy in grid units."""
    b, a, c = pred_distri.shape
    proj = torch.arange(reg_max, dtype=pred_distri.dtype)
    d = pred_distri.view(b, a, 4, c // 4).softmax(3).matmul(proj)
    return torch.cat((anchors - d[..., :2], anchors + d[..., 2:]), -1)


def assign_image(scores: torch.Tensor, boxes: torch.Tensor, anc: torch.Tensor, gt_labels: torch.Tensor,
                 gt_boxes: torch.Tensor, nc: int, topk: int = 10, alpha: float = 0.5, beta: float = 6.0, eps: float = 1e-9):
    """One image of TaskAlignedAssigner._forward (tal.py:79-118).  scores [A, nc] (sigmoid), boxes [A, 4] and anc [A, 2]
    in pixels, gt_labels [M] (long), gt_boxes [M, 4] (valid boxes only).  Returns target_boxes [A, 4] (pixels),
    target_scores [A, nc], fg [A] bool, gt_idx [A]."""
    A, M = boxes.shape[0], gt_boxes.shape[0]
    target_scores = torch.zeros((A, nc))
    target_boxes = torch.zeros((A, 4))
    fg = torch.zeros(A, dtype=torch.bool)
    gt_idx = torch.zeros(A, dtype=torch.long)
    if M == 0:
        return target_boxes, target_scores, fg, gt_idx
    overlaps = torch.zeros((M, A))
    align = torch.zeros((M, A))
    in_gt = torch.zeros((M, A), dtype=torch.bool)
    for g in range(M):
        x1, y1, x2, y2 = gt_boxes[g]
        d = torch.stack((anc[:, 0] - x1, anc[:, 1] - y1, x2 - anc[:, 0], y2 - anc[:, 1]), 1)
        in_gt[g] = d.amin(1) > eps                                                   # tal.py:254-273
        sel = in_gt[g]
        if sel.any():
            ov = ciou(gt_boxes[g].expand(int(sel.sum()), 4), boxes[sel]).clamp(min=0)  # tal.py:137-141
            overlaps[g, sel] = ov
            align[g, sel] = scores[sel, gt_labels[g]].pow(alpha) * ov.pow(beta)
    mask_pos = torch.zeros((M, A), dtype=torch.bool)
    for g in range(M):                                                                # tal.py:143-169
        idx = torch.topk(align[g], min(topk, A)).indices
        mask_pos[g, idx] = True
    mask_pos &= in_gt
    count = mask_pos.sum(0)
    multi = count > 1                                                                 # tal.py:276-297
    if multi.any():
        best = overlaps.argmax(0)
        only = torch.zeros_like(mask_pos)
        only[best, torch.arange(A)] = True
        mask_pos = torch.where(multi[None, :], only, mask_pos)
    fg = mask_pos.any(0)
    gt_idx = mask_pos.float().argmax(0)
    target_boxes = gt_boxes[gt_idx]
    am = align * mask_pos
    pos_align = am.amax(1, keepdim=True)
    pos_ov = (overlaps * mask_pos).amax(1, keepdim=True)
    norm = (am * pos_ov / (pos_align + eps)).amax(0)                                   # tal.py:112-116
    onehot = F.one_hot(gt_labels[gt_idx].clamp(min=0), nc).float() * fg[:, None]
    target_scores = onehot * norm[:, None]
    return target_boxes, target_scores, fg, gt_idx


def detection_loss(feats: List[torch.Tensor], batch: Dict[str, torch.Tensor], strides: Sequence[float], nc: int,
                   reg_max: int = 16, gains=(7.5, 0.5, 1.5), topk: int = 10):
    """loss.py:222-275.  feats: per-level head maps [B, 4*reg_max + nc, h, w] (may require grad); batch: 'batch_idx' [n],
    'cls' [n] or [n, 1], 'bboxes' [n, 4] normalised xywh.  Returns (loss.sum() * B, loss_items [3] = (box, cls, dfl) with
    gains, extras dict with the assigner's targets)."""
    B = feats[0].shape[0]
    no = nc + 4 * reg_max
    cat = torch.cat([f.view(B, no, -1) for f in feats], 2)
    pred_distri = cat[:, : 4 * reg_max].permute(0, 2, 1).contiguous()
    pred_scores = cat[:, 4 * reg_max:].permute(0, 2, 1).contiguous()
    hw = [tuple(f.shape[2:]) for f in feats]
    H, W = hw[0][0] * strides[0], hw[0][1] * strides[0]
    anchors, stride_t = make_anchors(hw, strides)
    pred_boxes = decode_boxes(pred_distri, anchors, reg_max)                          # grid units
    bi = batch["batch_idx"].view(-1).long()
    cls = batch["cls"].view(-1).long()
    bb = batch["bboxes"].view(-1, 4).float()
    scale = torch.tensor([W, H, W, H], dtype=torch.float32)
    A = anchors.shape[0]
    t_boxes = torch.zeros((B, A, 4))
    t_scores = torch.zeros((B, A, nc))
    fg = torch.zeros((B, A), dtype=torch.bool)
    gt_idx = torch.zeros((B, A), dtype=torch.long)
    for b in range(B):
        sel = bi == b
        xywh = bb[sel] * scale                                                        # loss.py:193-209
        xyxy = torch.cat((xywh[:, :2] - xywh[:, 2:] / 2, xywh[:, :2] + xywh[:, 2:] / 2), 1)
        labels = cls[sel]
        valid = xyxy.sum(1) > 0                                                       # mask_gt, loss.py:243
        tb, ts, f, gi = assign_image(pred_scores[b].detach().sigmoid(), (pred_boxes[b].detach() * stride_t),
                                     anchors * stride_t, labels[valid], xyxy[valid], nc, topk)
        t_boxes[b], t_scores[b], fg[b], gt_idx[b] = tb, ts, f, gi
    tss = max(float(t_scores.sum()), 1.0)
    loss = torch.zeros(3)
    loss_cls = F.binary_cross_entropy_with_logits(pred_scores, t_scores, reduction="none").sum() / tss
    loss_box = torch.zeros(())
    loss_dfl = torch.zeros(())
    if fg.any():
        tb_grid = t_boxes / stride_t
        weight = t_scores.sum(-1)[fg]
        iou = ciou(pred_boxes[fg], tb_grid[fg])
        loss_box = ((1.0 - iou) * weight).sum() / tss
        anc_b = anchors.expand(B, A, 2)[fg]
        ltrb = torch.cat((anc_b - tb_grid[fg][:, :2], tb_grid[fg][:, 2:] - anc_b), 1).clamp(0, reg_max - 1 - 0.01)
        tl = ltrb.long()
        wl = (tl + 1).float() - ltrb
        wr = 1.0 - wl
        logits = pred_distri[fg].view(-1, reg_max)
        ce_l = F.cross_entropy(logits, tl.view(-1), reduction="none").view(tl.shape)
        ce_r = F.cross_entropy(logits, (tl + 1).view(-1), reduction="none").view(tl.shape)
        loss_dfl = ((ce_l * wl + ce_r * wr).mean(-1) * weight).sum() / tss
    loss = torch.stack((loss_box * gains[0], loss_cls * gains[1], loss_dfl * gains[2]))
    extras = {"target_boxes": t_boxes, "target_scores": t_scores, "fg": fg, "gt_idx": gt_idx, "tss": tss}
    return loss.sum() * B, loss.detach(), extras


def loss_case(seed: int, B: int, H: int, W: int, nc: int, n_gt, dense: bool = False):
    """Seeded head maps + labels for the criterion fixtures / tests: logits whose decoded boxes are a few grid cells wide
    (so that many anchors have a positive CIoU with some box), ground truth of mixed sizes, optionally piled up so that
    anchors are claimed by several boxes (tal.py:276-297)."""
    g = torch.Generator().manual_seed(seed)
    no = 64 + nc
    feats = []
    for s in (8, 16, 32):
        f = torch.randn((B, no, H // s, W // s), generator=g)
        f[:, :64] *= 2.0
        f[:, 64:] = f[:, 64:] * 1.5 - 2.0
        feats.append(f)
    bi, cls, bb = [], [], []
    for b in range(B):
        n = n_gt[b]
        for _ in range(n):
            w = float(torch.rand((), generator=g)) * (0.5 if not dense else 0.3) + 0.04
            h = float(torch.rand((), generator=g)) * (0.5 if not dense else 0.3) + 0.04
            cx = (0.5 + (float(torch.rand((), generator=g)) - 0.5) * 0.2) if dense else float(torch.rand((), generator=g)) * (1 - w) + w / 2
            cy = (0.5 + (float(torch.rand((), generator=g)) - 0.5) * 0.2) if dense else float(torch.rand((), generator=g)) * (1 - h) + h / 2
            bi.append(b)
            cls.append(int(torch.randint(0, nc, (), generator=g)))
            bb.append([cx, cy, w, h])
    batch = {"batch_idx": torch.tensor(bi, dtype=torch.float32), "cls": torch.tensor(cls, dtype=torch.float32).view(-1, 1),
             "bboxes": torch.tensor(bb, dtype=torch.float32).view(-1, 4)}
    return feats, batch
