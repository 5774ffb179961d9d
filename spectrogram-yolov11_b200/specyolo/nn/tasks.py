"""YAML -> layer graph -> forward interpreter: the host-side mirror of ultralytics/nn/tasks.py
(parse_model :963-1168, DetectionModel :329-375, BaseModel._predict_once :161-188, fuse :223-251)
restricted to the modules of the Spectrogram-YOLOv11 and stock YOLO11 detection configs.

Differences by design: there is no CPU forward at build time (the reference probes strides with a
256x256 CPU forward, tasks.py:365 — here they follow from the graph), and `fuse()` folds BatchNorm and
repacks weights on the device for the tcgen05 kernels instead of rewriting nn.Conv2d modules.
"""
from __future__ import annotations

import ast
import contextlib
import math
import re
from copy import deepcopy
from pathlib import Path
from typing import List, Optional

import torch
import torch.nn as nn
import yaml

from .. import ops
from .modules import (C2PSA, C3, SPPF, Bottleneck, C2f, C3k, C3k2, C3k2GC, C3x, Concat, Conv, ConvHCA, DDWConv, Detect, DWConv,
                      Fusion, Upsample2x, UpsampledView)

CFG_DIR = Path(__file__).resolve().parent.parent / "cfg"

_MODULES = {m.__name__: m for m in (Conv, ConvHCA, DWConv, DDWConv, Bottleneck, C2f, C3, C3k, C3k2, C3k2GC, C3x, SPPF, C2PSA, Concat,
                                    Fusion, Detect)}
_BASE = {Conv, ConvHCA, DWConv, DDWConv, Bottleneck, C2f, C3, C3k, C3k2, C3k2GC, C3x, SPPF, C2PSA}
_REPEAT = {C2f, C3, C3k, C3k2, C3k2GC, C3x, C2PSA}
_STRIDE2 = {Conv, ConvHCA, DWConv, DDWConv}


def make_divisible(x, divisor):
    return math.ceil(x / divisor) * divisor


def guess_model_scale(path) -> str:
    """Scale letter embedded in the file name, e.g. yolo11s_fusion_sand3_new.yaml -> 's' (tasks.py:1187-1202)."""
    m = re.search(r"yolo[v]?\d+([nslmx])", Path(path).stem)
    return m.group(1) if m else ""


def yaml_model_load(path) -> dict:
    """Resolve 'yolo11s_fusion_sand3_new.yaml' -> cfg/yolo11_fusion_sand3_new.yaml + scale (tasks.py:1171-1184)."""
    path = Path(path)
    unified = re.sub(r"(\d+)([nslmx])(.+)?$", r"\1\3", path.stem)
    candidates = [path, path.with_name(unified + path.suffix), CFG_DIR / path.name, CFG_DIR / (unified + ".yaml")]
    for c in candidates:
        if c.is_file():
            d = yaml.safe_load(c.read_text())
            d["scale"] = guess_model_scale(path)
            d["yaml_file"] = str(path)
            return d
    raise FileNotFoundError(f"model config '{path}' not found (searched {[str(c) for c in candidates]})")


def parse_model(d: dict, ch: int, verbose: bool = False):
    """Build the nn.Sequential of blocks and the save list; also returns each layer's cumulative stride."""
    legacy = True
    max_channels = float("inf")
    nc, scales = d.get("nc"), d.get("scales")
    depth, width = d.get("depth_multiple", 1.0), d.get("width_multiple", 1.0)
    scale = d.get("scale")
    if scales:
        if not scale:
            scale = tuple(scales.keys())[0]
        depth, width, max_channels = scales[scale]
    chans = [ch]
    strides_in = [1]
    layers, save, lstride = [], [], []
    c2 = ch
    for i, (f, n, m, args) in enumerate(d["backbone"] + d["head"]):
        args = list(args)
        if m == "nn.Upsample":
            mod = Upsample2x
        elif m in _MODULES:
            mod = _MODULES[m]
        else:
            raise NotImplementedError(f"module '{m}' is outside the hot path this package implements")
        for j, a in enumerate(args):
            if isinstance(a, str):
                with contextlib.suppress(ValueError, SyntaxError):
                    args[j] = nc if a == "nc" else ast.literal_eval(a)
        n = n_ = max(round(n * depth), 1) if n > 1 else n
        fin = f if isinstance(f, int) else f[0]
        s_in = strides_in[fin] if fin != -1 else strides_in[-1]
        s_out = s_in
        if mod in _BASE:
            c1, c2 = chans[f], args[0]
            if c2 != nc:
                c2 = make_divisible(min(c2, max_channels) * width, 8)
            args = [c1, c2, *args[1:]]
            if mod in _REPEAT:
                args.insert(2, n)
                n = 1
            if mod is C3k2:
                legacy = False
                if scale in "mlx":
                    args[3] = True
            if mod is C3k2GC and scale in "mlx":        # tasks.py:1110-1112
                args[3] = True
            if mod in _STRIDE2:
                s = args[3] if len(args) > 3 else (1 if mod not in (DDWConv, ConvHCA) else 2)
                s_out = s_in * s
        elif mod is Upsample2x:
            c2 = chans[f]
            s_out = s_in // 2
        elif mod is Concat:
            c2 = sum(chans[x] for x in f)
        elif mod is Fusion:
            c2 = chans[f[0]]
            args = [[chans[x] for x in f], "ESChannel"]   # tasks.py:1132-1135
        elif mod is Detect:
            args.append([chans[x] for x in f])
            Detect.legacy = legacy      # read by Detect.__init__ only, which copies it onto the instance (head.py:43-58)
        m_ = nn.Sequential(*(mod(*args) for _ in range(n))) if n > 1 else mod(*args)
        m_.np = sum(x.numel() for x in m_.parameters())
        m_.i, m_.f, m_.type = i, f, m
        m_.c_out, m_.s_out = c2, s_out          # output channels / cumulative stride (concat planning, _concat_plan)
        save.extend(x % i for x in ([f] if isinstance(f, int) else f) if x != -1)
        layers.append(m_)
        if i == 0:
            chans, strides_in = [], []
        chans.append(c2)
        strides_in.append(s_out)
        lstride.append(s_out)
        if verbose:
            print(f"{i:>3}{str(f):>20}{n_:>3}{m_.np:10.0f}  {m:<20}{str(args):<30}")
    return nn.Sequential(*layers), sorted(save), lstride


class DetectionModel(nn.Module):
    """Detection model built from a YAML (ultralytics/nn/tasks.py:329-375)."""

    def __init__(self, cfg="yolo11s_fusion_sand3_new.yaml", ch=3, nc=None, verbose=False):
        super().__init__()
        self.yaml = cfg if isinstance(cfg, dict) else yaml_model_load(cfg)
        ch = self.yaml["ch"] = self.yaml.get("ch", ch)
        if nc and nc != self.yaml["nc"]:
            self.yaml["nc"] = nc
        self.model, self.save, lstride = parse_model(deepcopy(self.yaml), ch=ch, verbose=verbose)
        self.names = {i: f"{i}" for i in range(self.yaml["nc"])}
        self.inplace = True
        self.end2end = False
        m = self.model[-1]
        if isinstance(m, Detect):
            m.stride = torch.tensor([float(lstride[x]) for x in m.f])
            self.stride = m.stride
            m.bias_init()
        else:
            self.stride = torch.tensor([32.0])
        self.task = "detect"

    # ---- weight prep ---------------------------------------------------------------------------
    def fuse(self, verbose=False):
        """Fold BN + repack every conv now (otherwise done lazily on first forward)."""
        for m in self.modules():
            if hasattr(m, "packed"):
                m.packed()
        return self

    def is_fused(self, thresh=10):
        return all(getattr(m, "_packed", 1) is not None for m in self.modules())

    # ---- forward -------------------------------------------------------------------------------
    def forward(self, x, *args, **kwargs):
        return self.predict(x, *args, **kwargs)

    def predict(self, x, profile=False, visualize=False, augment=False, embed=None):
        if augment or visualize or embed or profile:
            raise NotImplementedError("augment / visualize / embed / profile are outside the hot path")
        return self._predict_once(x)

    def _parallel_groups(self):
        """Runs of consecutive trunk layers that read only earlier saved outputs (never each other or the running
        `x`), e.g. the five lateral / DDWConv layers 11-15 of the Spectrogram cfg: they can run concurrently."""
        if getattr(self, "_groups", None) is None:
            groups, i, n = {}, 0, len(self.model) - 1
            while i < n:
                f = self.model[i].f
                if isinstance(f, int) and f != -1:
                    j = i
                    while j + 1 < n:
                        fn = self.model[j + 1].f
                        if isinstance(fn, int) and fn != -1 and fn % (j + 1) < i:
                            j += 1
                        else:
                            break
                    if j > i and f % i < i:
                        groups[i] = j
                        i = j + 1
                        continue
                i += 1
            self._groups = groups
        return self._groups

    def _concat_plan(self):
        """Concat layers whose buffer can be planned ahead (stock YOLO11 necks: `Upsample -> Concat`, `Conv -> Concat`):
        {producer layer: (concat layer, channel offset, channels)} for inputs whose producer can write its output straight
        into the concat buffer's channel slice (`out=`), and {concat layer: [(input layer, offset, channels, how)]}.
        torch.cat (conv.py:1810-1820) then moves no bytes: the only kernel left is the x2 nearest upsample, which writes
        into its slice directly."""
        if getattr(self, "_cplan", None) is None:
            writers, cats = {}, {}
            n = len(self.model) - 1
            takes_out = (Conv, DWConv, DDWConv, C2f, C3, C3k, C3k2, SPPF, C2PSA, Fusion)
            for i in range(n):
                m = self.model[i]
                if not isinstance(m, Concat) or not isinstance(m.f, (list, tuple)):
                    continue
                srcs = [(i - 1) if j == -1 else j for j in m.f]
                if any(getattr(self.model[j], "c_out", None) is None or self.model[j].c_out % 8 for j in srcs):
                    continue
                off, entries = 0, []
                for j in srcs:
                    pm, c = self.model[j], self.model[j].c_out
                    if isinstance(pm, Upsample2x):
                        how = "upsample"
                    elif isinstance(pm, takes_out) and j not in writers and j != 0 and not (j == 1 and isinstance(self.model[0], Conv)):
                        how = "inplace"                     # (layers 0 / 1 may run as the fused stem pair: left alone)
                        writers[j] = (i, off, c)
                    else:
                        how = "copy"
                    entries.append((j, off, c, how))
                    off += c
                cats[i] = (entries, off, m.s_out)
            self._cplan = (writers, cats)
        return self._cplan

    def _run_trunk(self, x):
        """All layers except Detect; returns the list of Detect inputs."""
        y = []
        groups = DetectionModel._parallel_groups(self)      # (self may be the reference's model behind the shim)
        layers = list(self.model[:-1])
        # concat buffers planned ahead (our own layer classes only; behind the ultralytics shim the reference's Concat runs)
        plan_ok = isinstance(self, DetectionModel) and x.dim() == 4 and x.shape[2] % 32 == 0 and x.shape[3] % 32 == 0
        writers, cats = DetectionModel._concat_plan(self) if plan_ok else ({}, {})
        cat_bufs = {}
        in_b, in_h, in_w, in_dev = (x.shape[0], x.shape[2], x.shape[3], x.device) if plan_ok else (0, 0, 0, None)

        def cat_buf(ci):
            if ci not in cat_bufs:
                _, ctot, s_out = cats[ci]
                cat_bufs[ci] = ops.new_act(in_b, ctot, in_h // s_out, in_w // s_out, in_dev)
            return cat_bufs[ci]
        i = 0
        # stem pair: layer 0 writes its output 2x2-blocked and layer 1 (3x3 / s2) reads it as a 2x2 / s1 conv over
        # 4c channels — one 128-byte-row TMA box per tile instead of nine strided boxes (same arithmetic)
        if (len(layers) > 1 and hasattr(layers[0], "is_stem") and hasattr(layers[1], "takes_blocked")
                and type(layers[1]).__name__ == "Conv" and layers[0].is_stem()
                and layers[1].takes_blocked() and layers[1].f == -1 and 0 not in self.save and x.dim() == 4
                and x.stride(1) != 1 and x.shape[2] % 4 == 0 and x.shape[3] % 4 == 0
                and layers[0].packed().s2d is not None):
            if ops.stem_pair_ok(x, layers[0].packed(), layers[1].packed_from_blocked()):
                # uint8 input: both layers in one kernel, the layer-0 map never leaves shared memory
                x = ops.stem_pair(x, layers[0].packed(), layers[1].packed_from_blocked())
            else:
                xb = ops.stem_conv(x, layers[0].packed(), blocked_out=True)
                x = ops.conv2d(xb, layers[1].packed_from_blocked())
            y.extend([None, x if 1 in self.save else None])
            i = 2
        while i < len(layers):
            m = layers[i]
            if i in groups:                     # independent layers: fork / join on side streams
                j = groups[i]
                outs = ops.run_concurrently([lambda m=layers[k]: m(y[m.f]) for k in range(i, j + 1)])
                for k, o in zip(range(i, j + 1), outs):
                    y.append(o if layers[k].i in self.save or k == j else None)
                x = outs[-1]
                i = j + 1
                continue
            if m.f != -1:
                x = y[m.f] if isinstance(m.f, int) else [x if j == -1 else y[j] for j in m.f]
            if isinstance(x, UpsampledView) and type(m).__name__ != "Fusion":
                x = x.materialise()
            if i in cats:
                # planned concat: in-place inputs are already there, upsampled inputs are written by the upsample kernel
                buf = cat_buf(i)
                for (j, off, c, how), t in zip(cats[i][0], x):
                    if how == "upsample":
                        t.materialise(out=buf[:, off: off + c])
                    elif how == "copy" or t.data_ptr() != buf[:, off: off + c].data_ptr():
                        buf[:, off: off + c].copy_(t.materialise() if isinstance(t, UpsampledView) else t)
                x = buf
            elif i in writers:
                ci, off, c = writers[i]
                x = m(x, out=cat_buf(ci)[:, off: off + c])
            else:
                x = m(x)       # (the Concat fallback handles UpsampledView inputs itself)
            y.append(x if m.i in self.save else None)
            i += 1
        det = self.model[-1]
        return [x if j == -1 else y[j] for j in det.f]

    def _predict_once(self, x):
        """tasks.py:161-188 — the layer interpreter."""
        if not x.is_cuda:
            raise RuntimeError("specyolo runs on a CUDA device (B200); there is no CPU fallback")
        feats = self._run_trunk(x)
        return self.model[-1](feats)

    @torch.no_grad()
    def detect_fused(self, x, conf_thres=0.25, iou_thres=0.7, agnostic=False, max_det=300, max_nms=30000,
                     max_wh=7680.0, classes: Optional[torch.Tensor] = None, clip: bool = True):
        """Trunk -> head logits -> fused decode+threshold -> batched NMS, never materialising the dense
        [B, 4+nc, A] prediction.  Returns device tensors (out [B,max_det,6], count [B]).  `clip`: kept boxes are
        clamped to the network input's (H, W) as they are written — the clip_boxes every construct_result ends with
        (detect/predict.py:59-73 -> ops.scale_boxes :124-127), also when gain is 1 and the padding 0."""
        det: Detect = self.model[-1]
        feats = self._run_trunk(x)
        bufs, hw = det.head_logits(feats)
        _, cand, seg = ops.detect_decode(bufs, hw, [float(s) for s in det.stride], det.nc, want_dense=False,
                                         conf_thres=conf_thres)
        A = sum(h * w for h, w in hw)
        out, cnt, _, _ = ops.nms(cand=cand, seg_count=seg, B=bufs[0].shape[0], nc=det.nc, A=A, conf_thres=conf_thres,
                                 iou_thres=iou_thres, agnostic=agnostic, max_det=max_det, max_nms=max_nms,
                                 max_wh=max_wh, classes=classes,
                                 clip_hw=tuple(x.shape[2:]) if clip else None)
        return out, cnt


def attempt_load_one_weight(weight, device=None, inplace=True, fuse=False):
    """ultralytics/nn/tasks.py:937-960 (see specyolo.nn.checkpoint)."""
    from .checkpoint import attempt_load_one_weight as f
    return f(weight, device=device, inplace=inplace, fuse=fuse)


def torch_safe_load(weight):
    """ultralytics/nn/tasks.py:846-898 (see specyolo.nn.checkpoint)."""
    from .checkpoint import torch_safe_load as f
    return f(weight)
