"""Deterministic synthetic weights ("random-init weights of that architecture" for benches and parity
tests — there is no network for checkpoints).  The recipe follows SURVEY 8(d): BatchNorm running stats
and affine terms are randomised so the BN fold is exercised, GCT gamma/beta are non-zero so the Fusion
gates are not the identity they are at init, and the Detect class bias is raised so that a few percent
of the anchors clear conf=0.25 and NMS has real work.

Everything is generated on the CPU with a seeded torch.Generator, key by key in state_dict order, so the
same (cfg, seed) gives bit-identical tensors in the build container and on the GPU box.
"""
from __future__ import annotations

import json
import math
from pathlib import Path
from typing import Dict

import torch


def _load_scales(model) -> Dict[str, float]:
    """Per-conv RMS table measured offline by tools/calibrate_synth.py (plain data, cfg/<name>.synth.json)."""
    name = Path(str(getattr(model, "yaml", {}).get("yaml_file", ""))).stem
    f = Path(__file__).resolve().parent.parent / "cfg" / f"{name}.synth.json"
    return json.loads(f.read_text()) if name and f.is_file() else {}


def synth_state_dict(model: torch.nn.Module, seed: int = 0, cls_bias: float | None = None,
                     calibrated: bool = True) -> Dict[str, torch.Tensor]:
    if cls_bias is None:   # ~1-3 % of the anchors clear conf=0.25 for 2 classes as for 80
        nc = int(getattr(model, "yaml", {}).get("nc", 2))
        cls_bias = -2.4 - 0.45 * math.log(max(nc, 2) / 2.0)
    g = torch.Generator().manual_seed(seed)
    scales = _load_scales(model) if calibrated else {}
    out: Dict[str, torch.Tensor] = {}
    for k, v in model.state_dict().items():
        shape = tuple(v.shape)
        if k.endswith("num_batches_tracked"):
            t = torch.zeros(shape, dtype=v.dtype)
        elif k.endswith(".dfl.conv.weight"):
            t = torch.arange(shape[1], dtype=torch.float32).view(shape)
        elif k.endswith("bn.running_mean") or k.endswith("bn.bias"):
            t = torch.randn(shape, generator=g) * 0.1
        elif k.endswith("bn.running_var") or k.endswith("bn.weight"):
            t = torch.rand(shape, generator=g) + 0.5
        elif k.endswith(".alpha"):
            t = torch.rand(shape, generator=g) * 0.5 + 0.75
        elif k.endswith(".gamma") or k.endswith(".beta"):
            t = torch.randn(shape, generator=g) * 0.3
        elif k.endswith("sab.cv1.weight"):
            t = torch.randn(shape, generator=g) * 0.5
        elif k.endswith(".weight") and len(shape) == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            t = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)
        elif k.endswith(".bias"):
            # last conv of a Detect branch: cv2 (box, 64 outputs) or cv3 (classes)
            is_cls = ".cv3." in k
            t = torch.randn(shape, generator=g) * 0.1 + (cls_bias if is_cls else 1.0)
        else:
            t = torch.randn(shape, generator=g) * 0.1
        # BatchNorm running stats follow the measured RMS of the conv output (what training would give);
        # the last conv of each Detect branch is normalised to unit-RMS logits (x1.5 for the DFL bins)
        if k.endswith("bn.running_var") and k[: -len(".bn.running_var")] in scales:
            t = t * scales[k[: -len(".bn.running_var")]] ** 2
        elif k.endswith("bn.running_mean") and k[: -len(".bn.running_mean")] in scales:
            t = t * scales[k[: -len(".bn.running_mean")]]
        elif k.endswith(".2.weight") and k[: -len(".weight")] in scales:
            t = t * ((1.5 if ".cv2." in k else 1.0) / scales[k[: -len(".weight")]])
        out[k] = t.to(v.dtype)
    return out


def synth_det_batch(batch: int, size: int = 640, nc: int = 2, boxes_per_image: int = 8, seed: int = 0, head_maps: bool = False):
    """Synthetic detection labels in the reference's collated-batch format (`batch_idx` [n], `cls` [n, 1], `bboxes` [n, 4]
    normalised xywh; data/dataset.py:232-252) and, on request, seeded Detect head maps `[B, 64 + nc, size/s, size/s]` for
    s = 8, 16, 32 whose decoded boxes span a few grid cells (inputs of the criterion benchmark).  Bench / demo data only."""
    g = torch.Generator().manual_seed(5000 + seed)
    n = batch * boxes_per_image
    wh = torch.rand((n, 2), generator=g) * 0.5 + 0.04
    ctr = torch.rand((n, 2), generator=g) * (1 - wh) + wh / 2
    labels = {"batch_idx": torch.arange(batch, dtype=torch.float32).repeat_interleave(boxes_per_image),
              "cls": torch.randint(0, nc, (n, 1), generator=g).float(), "bboxes": torch.cat((ctr, wh), 1)}
    if not head_maps:
        return labels
    feats = []
    for s in (8, 16, 32):
        f = torch.randn((batch, 64 + nc, size // s, size // s), generator=g)
        f[:, :64] *= 2.0
        f[:, 64:] = f[:, 64:] * 1.5 - 2.0
        feats.append(f)
    return feats, labels


def synth_images(batch: int, size: int = 640, seed: int = 0, dtype=torch.float32) -> torch.Tensor:
    """Spectrogram-like frames: noise floor + a few bright axis-aligned rectangles."""
    g = torch.Generator().manual_seed(1000 + seed)
    x = (torch.randn((batch, 1, size, size), generator=g) * 0.05 + 0.25)
    for b in range(batch):
        n = int(torch.randint(4, 13, (1,), generator=g))
        for _ in range(n):
            w = int(torch.randint(max(2, size // 32), max(3, size * 15 // 32), (1,), generator=g))
            h = int(torch.randint(max(2, size // 80), max(3, size * 3 // 16), (1,), generator=g))
            x0 = int(torch.randint(0, size - w, (1,), generator=g))
            y0 = int(torch.randint(0, size - h, (1,), generator=g))
            lvl = float(torch.rand((1,), generator=g)) * 0.3 + 0.6
            x[b, 0, y0:y0 + h, x0:x0 + w] = lvl + torch.randn((h, w), generator=g) * 0.03
    x = x.clamp_(0, 1).expand(batch, 3, size, size).contiguous()
    if dtype == torch.uint8:
        return (x * 255).round().to(torch.uint8)
    return x.to(dtype)


def synth_iq(batch: int, length: int = 1 << 20, seed: int = 0) -> torch.Tensor:
    """complex64 bursts: complex Gaussian noise + a few band-limited on/off carriers."""
    out = torch.empty((batch, length), dtype=torch.complex64)
    t = torch.arange(length, dtype=torch.float64)
    for b in range(batch):
        g = torch.Generator().manual_seed(1000 + seed * 7919 + b)
        x = torch.complex(torch.randn(length, generator=g, dtype=torch.float64),
                          torch.randn(length, generator=g, dtype=torch.float64)) * (0.1 / math.sqrt(2))
        n = int(torch.randint(2, 7, (1,), generator=g))
        for _ in range(n):
            fc = (float(torch.rand((1,), generator=g)) - 0.5) * 0.9
            bw = float(torch.rand((1,), generator=g)) * 0.19 + 0.01
            t0 = int(torch.randint(0, length // 2, (1,), generator=g))
            dur = int(torch.randint(length // 16, length // 2, (1,), generator=g))
            amp = 10 ** (float(torch.rand((1,), generator=g)) * 1.0 - 0.5) * 0.1
            # band-limited noise-like carrier: sum of a few tones inside the band
            seg = torch.zeros(dur, dtype=torch.complex128)
            tt = t[:dur]
            for _k in range(8):
                f = fc + (float(torch.rand((1,), generator=g)) - 0.5) * bw
                ph = float(torch.rand((1,), generator=g)) * 2 * math.pi
                seg += torch.exp(1j * (2 * math.pi * f * tt + ph))
            x[t0:t0 + dur] += seg[: max(0, min(dur, length - t0))] * (amp / math.sqrt(8))
        out[b] = x.to(torch.complex64)
    return out


# dBFS window that maps `synth_iq_emissions` bursts onto the grey levels the synthetic weights were calibrated on
# (noise floor ~0.25, emissions 0.6 - 0.9): pass as predict_iq(..., db_min=, db_max=)
EMISSION_DB_RANGE = (-83.0, -23.0)
# class-logit bias of the synthetic Detect head for the IQ workload: the random-weight head scores letterboxed
# spectrograms (content only in the middle 160 rows) lower than the calibration images; +0.8 on the default -2.4 puts
# ~1.5 % of the anchors above conf = 0.25 again (17-67 detections per burst) so decode / NMS see a realistic load
IQ_CLS_BIAS = -1.6
# same for stock yolo11s (nc = 80) at 1280^2 (~400-1000 candidates per image at 1280^2)
YOLO11S_1280_CLS_BIAS = -2.05


def synth_iq_emissions(batch: int, length: int = 1 << 20, seed: int = 0) -> torch.Tensor:
    """complex64 bursts that look like the 5G / LTE captures the detector is meant for: a complex Gaussian
    noise floor (sigma 0.01: -68 dBFS per 1024-point bin under a Hann window) plus 4-12 band-limited, time-gated
    noise-like emissions (flat spectrum over 1-24 % of fs — an OFDM-like block — lasting 2-20 % of the burst) whose
    spectral density sits 21-39 dB above the floor: axis-aligned bright rectangles in the spectrogram, like the images
    the synthetic weights are calibrated on.  Deterministic in (seed, burst index)."""
    out = torch.empty((batch, length), dtype=torch.complex64)
    floor = 0.01
    for b in range(batch):
        g = torch.Generator().manual_seed(2000 + seed * 7919 + b)
        x = torch.complex(torch.randn(length, generator=g, dtype=torch.float32),
                          torch.randn(length, generator=g, dtype=torch.float32)) * (floor / math.sqrt(2))
        for _ in range(int(torch.randint(4, 13, (1,), generator=g))):
            bw = float(torch.rand((1,), generator=g)) * 0.23 + 0.01
            fc = (float(torch.rand((1,), generator=g)) - 0.5) * (0.96 - bw)
            dur = int((float(torch.rand((1,), generator=g)) * 0.18 + 0.02) * length) // 2 * 2
            t0 = int(torch.randint(0, length - dur, (1,), generator=g))
            gain_db = float(torch.rand((1,), generator=g)) * 18.0 + 21.0
            # white noise of the emission's duration, brick-wall filtered to [fc - bw/2, fc + bw/2] in the frequency domain
            w = torch.complex(torch.randn(dur, generator=g, dtype=torch.float32),
                              torch.randn(dur, generator=g, dtype=torch.float32)) * (floor / math.sqrt(2))
            W = torch.fft.fft(w)
            f = torch.fft.fftfreq(dur)
            W[((f - fc + 0.5) % 1.0 - 0.5).abs() > bw / 2] = 0
            x[t0:t0 + dur] += torch.fft.ifft(W) * (10.0 ** (gain_db / 20.0))
        out[b] = x
    return out
