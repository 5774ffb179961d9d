"""TEST INFRASTRUCTURE — numpy restatement of the reference's uint8 LetterBox (never imported by the product).

LetterBox.__call__            ultralytics/data/augment.py:1535-1601 (geometry :1566-1591)
cv2.resize(INTER_LINEAR) on uint8 — third-party: opencv-python >= 4.6.0 (pyproject.toml), installed 4.13; its 8-bit
bilinear arithmetic (modules/imgproc/src/resize.cpp: 11-bit fixed-point coefficients, HResizeLinear / VResizeLinear,
the exact-2x-shrink area fast path) is restated here and was checked bit for bit against cv2 4.13 in the build container.
Pinned by tests/golden/letterbox_u8.npz = outputs of the real LetterBox (oracle/gen_golden.py: gen_letterbox_u8).
"""
from __future__ import annotations

import numpy as np


def geometry(h, w, new_shape=(640, 640), auto=False, scale_fill=False, scaleup=True, center=True, stride=32):
    nh, nw = new_shape
    r = min(nh / h, nw / w)
    if not scaleup:
        r = min(r, 1.0)
    new_w, new_h = int(round(w * r)), int(round(h * r))
    dw, dh = nw - new_w, nh - new_h
    if auto:
        dw, dh = dw % stride, dh % stride
    elif scale_fill:
        dw, dh, new_w, new_h = 0.0, 0.0, nw, nh
    if center:
        dw, dh = dw / 2, dh / 2
    top, bottom = (int(round(dh - 0.1)) if center else 0), int(round(dh + 0.1))
    left, right = (int(round(dw - 0.1)) if center else 0), int(round(dw + 0.1))
    return new_w, new_h, left, top, new_h + top + bottom, new_w + left + right


def _coef_x(src, dst):
    scale = 1.0 / (dst / src)
    s = np.empty(dst, np.int64); a0 = np.empty(dst, np.int64); a1 = np.empty(dst, np.int64)
    for d in range(dst):
        f = np.float32((d + 0.5) * scale - 0.5)
        i = int(np.floor(f)); f = np.float32(f - np.float32(i))
        if i < 0:
            f, i = np.float32(0), 0
        if i >= src - 1:
            f, i = np.float32(0), src - 1
        a0[d] = int(np.rint(np.float32((np.float32(1) - f) * np.float32(2048))))
        a1[d] = int(np.rint(np.float32(f * np.float32(2048))))
        s[d] = i
    return s, a0, a1


def _coef_y(src, dst):
    scale = 1.0 / (dst / src)
    r0 = np.empty(dst, np.int64); r1 = np.empty(dst, np.int64); b0 = np.empty(dst, np.int64); b1 = np.empty(dst, np.int64)
    for d in range(dst):
        f = np.float32((d + 0.5) * scale - 0.5)
        i = int(np.floor(f)); f = np.float32(f - np.float32(i))
        b0[d] = int(np.rint(np.float32((np.float32(1) - f) * np.float32(2048))))
        b1[d] = int(np.rint(np.float32(f * np.float32(2048))))
        r0[d], r1[d] = min(max(i, 0), src - 1), min(max(i + 1, 0), src - 1)
    return r0, r1, b0, b1


def resize_u8(img: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR) for HWC uint8."""
    H, W, _ = img.shape
    x = img.astype(np.int64)
    if (W, H) == (dw, dh):
        return img.copy()
    if W == 2 * dw and H == 2 * dh:
        return ((x[0::2, 0::2] + x[0::2, 1::2] + x[1::2, 0::2] + x[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    sx, a0, a1 = _coef_x(W, dw)
    r0, r1, b0, b1 = _coef_y(H, dh)
    rows = x[:, sx, :] * a0[None, :, None] + x[:, np.minimum(sx + 1, W - 1), :] * a1[None, :, None]
    out = (((b0[:, None, None] * (rows[r0] >> 4)) >> 16) + ((b1[:, None, None] * (rows[r1] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def letterbox(img: np.ndarray, **kw) -> np.ndarray:
    h, w = img.shape[:2]
    new_w, new_h, left, top, oh, ow = geometry(h, w, **kw)
    out = np.full((oh, ow, 3), 114, np.uint8)
    out[top:top + new_h, left:left + new_w] = resize_u8(img, new_w, new_h)
    return out
