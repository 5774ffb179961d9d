"""`LetterBox` of the reference's ingest path (ultralytics/data/augment.py:1477-1601) on the device.

Same constructor and geometry; `__call__(image=<HWC uint8 ndarray | CUDA tensor>)` returns the letterboxed HWC image
like the reference (as a CUDA tensor, or an ndarray when given one), bit-exact with cv2.resize(INTER_LINEAR) +
copyMakeBorder(114).  `to_network_input` is what the predictor uses: letterbox + BGR->RGB + HWC->CHW in one kernel
launch for a batch of same-sized images (ultralytics/engine/predictor.py:125-136, 147-163).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from .. import ops


class LetterBox:
    def __init__(self, new_shape=(640, 640), auto=False, scale_fill=False, scaleup=True, center=True, stride=32):
        self.new_shape = (new_shape, new_shape) if isinstance(new_shape, int) else tuple(new_shape)
        self.auto = auto
        self.scaleFill = scale_fill
        self.scaleup = scaleup
        self.stride = stride
        self.center = center

    def geometry(self, h: int, w: int) -> Tuple[int, int, int, int, int, int]:
        """(new_w, new_h, left, top, out_h, out_w) for an h x w image (augment.py:1566-1591)."""
        nh, nw = self.new_shape
        r = min(nh / h, nw / w)
        if not self.scaleup:
            r = min(r, 1.0)
        new_w, new_h = int(round(w * r)), int(round(h * r))
        dw, dh = nw - new_w, nh - new_h
        if self.auto:                           # minimum rectangle
            dw, dh = dw % self.stride, dh % self.stride
        elif self.scaleFill:                    # stretch
            dw, dh = 0.0, 0.0
            new_w, new_h = nw, nh
        if self.center:
            dw /= 2
            dh /= 2
        top, bottom = (int(round(dh - 0.1)) if self.center else 0), int(round(dh + 0.1))
        left, right = (int(round(dw - 0.1)) if self.center else 0), int(round(dw + 0.1))
        return new_w, new_h, left, top, new_h + top + bottom, new_w + left + right

    def _run(self, imgs: torch.Tensor, swap_rb: bool, chw: bool, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        B, H, W, C = imgs.shape
        if C != 3 or imgs.dtype != torch.uint8:
            raise ValueError("LetterBox expects HWC uint8 images with 3 channels")
        new_w, new_h, left, top, oh, ow = self.geometry(H, W)
        return ops.letterbox_u8(imgs, (new_w, new_h, left, top, oh, ow), swap_rb=swap_rb, chw=chw, out=out)

    def __call__(self, labels=None, image=None):
        if labels:
            raise NotImplementedError("label transforms belong to the training pipeline (SURVEY 8 f2)")
        is_np = isinstance(image, np.ndarray)
        t = torch.from_numpy(np.ascontiguousarray(image)).cuda() if is_np else image
        y = self._run(t.unsqueeze(0).contiguous(), swap_rb=False, chw=False)[0]
        return y.cpu().numpy() if is_np else y

    def to_network_input(self, imgs: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """[B,H,W,3] uint8 BGR (CUDA) -> [B,3,out_h,out_w] uint8 RGB, letterboxed."""
        return self._run(imgs.contiguous(), swap_rb=True, chw=True, out=out)
