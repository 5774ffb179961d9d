"""TEST INFRASTRUCTURE — float64 specification of the IQ -> spectrogram -> letterbox front end.

PARITY UNPINNED: the reference has no IQ / STFT code at all (README.md:7 mentions spectrogram
images; cfg/datasets/Spectrogram.yaml:1-4 points at pre-rendered image folders), so there is nothing
to follow or to pin against.  This file freezes the convention that csrc/stft.cu implements; the only
reference-derived part is the letterbox geometry, which restates LetterBox.__call__
(ultralytics/data/augment.py:1566-1591) with auto=False, scaleup=True, center=True, and that part IS
pinned against the real LetterBox/cv2 by tests (geometry: new_unpad, top/left; sampling: cv2.INTER_LINEAR
half-pixel taps, compared in float).

Convention (see csrc/stft.cu header): frames of nfft samples every hop (center=False), periodic Hann,
forward DFT, fftshift so row 0 = -fs/2, power in dB relative to a unit-amplitude tone ((sum w)^2),
affine map of [db_min, db_max] to [0,1] with clamping, then bilinear letterbox of the nfft x T image,
padding value 114/255, three identical channels.

Known-answer checks live in tests/test_stft_oracle.py (tone -> peak row, impulse -> flat spectrum,
Parseval) together with a cross-check of the STFT convention against torch.stft and scipy.fft on random samples
(independent library implementations — not the reference, which has none: the header above stands).
"""
from __future__ import annotations

import numpy as np


def hann_periodic(n: int) -> np.ndarray:
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def stft_power(iq: np.ndarray, nfft: int = 1024, hop: int = 256) -> np.ndarray:
    """iq complex [L] -> |X|^2 [nfft, T] (float64), rows fftshifted (row 0 = -fs/2)."""
    iq = np.asarray(iq, dtype=np.complex128)
    L = iq.shape[0]
    T = 1 + (L - nfft) // hop
    idx = np.arange(nfft)[None, :] + hop * np.arange(T)[:, None]
    frames = iq[idx] * hann_periodic(nfft)[None, :]
    X = np.fft.fftshift(np.fft.fft(frames, axis=1), axes=1)
    return (X.real ** 2 + X.imag ** 2).T


def normalise_db(power: np.ndarray, nfft: int, db_min: float = -100.0, db_max: float = 0.0) -> np.ndarray:
    pref = hann_periodic(nfft).sum() ** 2
    db = 10.0 * np.log10(np.maximum(power, 1e-20) / pref)
    return np.clip((db - db_min) / (db_max - db_min), 0.0, 1.0)


def letterbox_geometry(shape_hw, new_shape_hw):
    """augment.py:1566-1588 (auto=False, scaleFill=False, scaleup=True, center=True).
    Returns (new_w, new_h, left, top)."""
    h, w = shape_hw
    r = min(new_shape_hw[0] / h, new_shape_hw[1] / w)
    new_w, new_h = int(round(w * r)), int(round(h * r))
    dw, dh = (new_shape_hw[1] - new_w) / 2, (new_shape_hw[0] - new_h) / 2
    top, left = int(round(dh - 0.1)), int(round(dw - 0.1))
    return new_w, new_h, left, top


def bilinear_resize(img: np.ndarray, new_w: int, new_h: int) -> np.ndarray:
    """Half-pixel bilinear resize without antialiasing (cv2.INTER_LINEAR geometry, float arithmetic)."""
    H, W = img.shape
    fy = np.maximum((np.arange(new_h) + 0.5) * (H / new_h) - 0.5, 0.0)
    fx = np.maximum((np.arange(new_w) + 0.5) * (W / new_w) - 0.5, 0.0)
    y0 = np.minimum(np.floor(fy).astype(np.int64), H - 1)
    x0 = np.minimum(np.floor(fx).astype(np.int64), W - 1)
    y1 = np.minimum(y0 + 1, H - 1)
    x1 = np.minimum(x0 + 1, W - 1)
    wy = (fy - y0)[:, None]
    wx = (fx - x0)[None, :]
    a = img[y0][:, x0]
    b = img[y0][:, x1]
    c = img[y1][:, x0]
    d = img[y1][:, x1]
    top = a + wx * (b - a)
    bot = c + wx * (d - c)
    return top + wy * (bot - top)


def iq_to_letterbox(iq: np.ndarray, nfft: int = 1024, hop: int = 256, db_min: float = -100.0, db_max: float = 0.0,
                    out_hw=(640, 640), pad_value: float = 114.0 / 255.0) -> np.ndarray:
    """iq complex [B, L] -> float64 [B, 3, out_h, out_w]."""
    iq = np.asarray(iq)
    out = np.full((iq.shape[0], 3, out_hw[0], out_hw[1]), pad_value, dtype=np.float64)
    for b in range(iq.shape[0]):
        v = normalise_db(stft_power(iq[b], nfft, hop), nfft, db_min, db_max)
        new_w, new_h, left, top = letterbox_geometry(v.shape, out_hw)
        out[b, :, top:top + new_h, left:left + new_w] = bilinear_resize(v, new_w, new_h)[None]
    return out
