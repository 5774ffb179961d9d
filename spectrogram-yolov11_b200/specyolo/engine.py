"""`YOLO(cfg).predict()` facade, predictor loop and result containers — the host-side mirror of
ultralytics/engine/model.py:499-558 (Model.predict), engine/predictor.py:118-306 (BasePredictor),
models/yolo/detect/predict.py:23-73 (DetectionPredictor.postprocess) and engine/results.py (Results / Boxes),
reduced to what the detection hot path returns.

The whole step (stem -> ... -> Detect head -> fused decode+threshold -> batched NMS) is captured once per
(input shape, thresholds) in a CUDA graph and replayed: ~170 kernel launches, zero host synchronisation
inside, one small device->host copy of [B, max_det, 6] + counts at the end.
"""
from __future__ import annotations

import collections
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Union

import numpy as np
import torch

from . import _lib, ops
from .nn.tasks import DetectionModel


class Boxes:
    """[n, 6] detections (x1, y1, x2, y2, conf, cls) in original-image pixels (engine/results.py:1015)."""

    def __init__(self, data: torch.Tensor, orig_shape):
        self.data = data
        self.orig_shape = orig_shape

    @property
    def xyxy(self):
        return self.data[:, :4]

    @property
    def conf(self):
        return self.data[:, 4]

    @property
    def cls(self):
        return self.data[:, 5]

    @property
    def xywh(self):
        b = self.xyxy
        return torch.cat(((b[:, :2] + b[:, 2:]) / 2, b[:, 2:] - b[:, :2]), 1)

    @property
    def xyxyn(self):
        """Boxes normalised by the original image size (engine/results.py:1185-1206)."""
        xyxy = self.xyxy.clone() if isinstance(self.xyxy, torch.Tensor) else np.copy(self.xyxy)
        xyxy[..., [0, 2]] /= self.orig_shape[1]
        xyxy[..., [1, 3]] /= self.orig_shape[0]
        return xyxy

    @property
    def xywhn(self):
        """(cx, cy, w, h) normalised by the original image size (engine/results.py:1209-1229)."""
        xywh = self.xywh
        xywh = xywh.clone() if isinstance(xywh, torch.Tensor) else np.copy(xywh)
        xywh[..., [0, 2]] /= self.orig_shape[1]
        xywh[..., [1, 3]] /= self.orig_shape[0]
        return xywh

    def __len__(self):
        return self.data.shape[0]

    def cpu(self):
        return Boxes(self.data.cpu(), self.orig_shape)

    def numpy(self):
        return Boxes(self.data.cpu().numpy(), self.orig_shape)


class Results:
    """One image's detections (engine/results.py:187).  `orig_img` is kept by reference only when the caller
    passed host images; tensor sources are not copied back to the host (the reference does, predict.py:37-38)."""

    def __init__(self, orig_shape, boxes: torch.Tensor, names: Dict[int, str], path: str = "", orig_img=None,
                 speed: Optional[dict] = None):
        self.orig_shape = orig_shape
        self.boxes = Boxes(boxes, orig_shape)
        self.names = names
        self.path = path
        self.orig_img = orig_img
        self.speed = speed or {}

    def __len__(self):
        return len(self.boxes)

    # ---- egress formats of the reference (engine/results.py) ------------------------------------
    def verbose(self) -> str:
        """'2 wifis, 1 bluetooth, ' (engine/results.py:630-666)."""
        if len(self) == 0:
            return "(no detections), "
        cls = self.boxes.data[:, 5]
        s = ""
        for c in torch.unique(cls.cpu()) if isinstance(cls, torch.Tensor) else np.unique(cls):
            n = int((cls == c).sum())
            s += f"{n} {self.names[int(c)]}{'s' * (n > 1)}, "
        return s

    def summary(self, normalize: bool = False, decimals: int = 5) -> list:
        """[{name, class, confidence, box{x1,y1,x2,y2}}] (engine/results.py:759-824)."""
        d = self.boxes.data
        d = d.cpu() if isinstance(d, torch.Tensor) else torch.from_numpy(np.asarray(d))
        h, w = self.orig_shape if normalize else (1, 1)
        out = []
        for row in d:
            class_id, conf = int(row[5]), round(row[4].item(), decimals)
            box = row[:4].reshape(-1, 2).tolist()
            xy = {}
            for j, b in enumerate(box):
                xy[f"x{j + 1}"] = round(b[0] / w, decimals)
                xy[f"y{j + 1}"] = round(b[1] / h, decimals)
            out.append({"name": self.names[class_id], "class": class_id, "confidence": conf, "box": xy})
        return out

    def to_json(self, normalize: bool = False, decimals: int = 5) -> str:
        """engine/results.py:907-944."""
        import json

        return json.dumps(self.summary(normalize=normalize, decimals=decimals), indent=2)

    tojson = to_json

    def to_df(self, normalize: bool = False, decimals: int = 5):
        """engine/results.py:826-849."""
        import pandas as pd

        return pd.DataFrame(self.summary(normalize=normalize, decimals=decimals))

    def to_csv(self, normalize: bool = False, decimals: int = 5, *args, **kwargs):
        """engine/results.py:851-877."""
        return self.to_df(normalize=normalize, decimals=decimals).to_csv(*args, **kwargs)

    def save_txt(self, txt_file, save_conf: bool = False):
        """YOLO label lines `cls cx cy w h [conf]`, normalised, appended to txt_file (engine/results.py:669-722)."""
        d = self.boxes.data
        d = d.cpu() if isinstance(d, torch.Tensor) else torch.from_numpy(np.asarray(d))
        nb = Boxes(d, self.orig_shape)
        texts = []
        for row, xywhn in zip(d, nb.xywhn):
            line = (int(row[5]), *xywhn.view(-1)) + ((float(row[4]),) if save_conf else ())
            texts.append(("%g " * len(line)).rstrip() % line)
        if texts:
            Path(txt_file).parent.mkdir(parents=True, exist_ok=True)
            with open(txt_file, "a") as f:
                f.writelines(t + "\n" for t in texts)


class _GraphStep:
    """A captured forward for one static input signature."""

    def __init__(self, model: DetectionModel, example: torch.Tensor, conf, iou, agnostic, max_det, classes, front=None):
        self.static_in = example.clone()
        self.front = front          # optional device-side front end captured with the detector (IQ -> spectrogram image)
        # the class-filter tensor's address is baked into the captured NMS launch: the graph owns its own copy
        self.classes = None if classes is None else classes.detach().clone()
        args = dict(conf_thres=conf, iou_thres=iou, agnostic=agnostic, max_det=max_det, classes=self.classes)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        fwd = (lambda x: model.detect_fused(x, **args)) if front is None else \
            (lambda x: model.detect_fused(front(x), **args))
        with torch.cuda.stream(side):      # warm-up: lazy weight packing, attribute setup, allocator
            for _ in range(2):
                fwd(self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        n0 = _lib.load().specyolo_launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out, self.cnt = fwd(self.static_in)
        self.launches = int(_lib.load().specyolo_launch_count() - n0)

    def run(self, x: torch.Tensor):
        if x.data_ptr() != self.static_in.data_ptr():       # callers may fill the static input in place (zero copy)
            self.static_in.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.out, self.cnt

    def clone(self, model: DetectionModel, conf, iou, agnostic, max_det, classes) -> "_GraphStep":
        """A second, independent instance (own static input / outputs) for double-buffered streaming."""
        return _GraphStep(model, self.static_in, conf, iou, agnostic, max_det, classes, front=self.front)


class IQFrontEnd:
    """IQ bursts -> letterboxed spectrogram images on the device (absent in the reference): the callable a
    `_GraphStep` captures in front of the detector, so one graph replay = STFT kernel + ~95 detector launches.
    Input: float32 [B, L, 2] (complex64 viewed as real), output bf16 NCHW [B, 3, H, W] in [0, 1]."""

    def __init__(self, nfft=1024, hop=256, db_min=-100.0, db_max=0.0, out_hw=(640, 640)):
        self.nfft, self.hop, self.db_min, self.db_max, self.out_hw = nfft, hop, db_min, db_max, tuple(out_hw)

    def key(self):
        return ("iq", self.nfft, self.hop, self.db_min, self.db_max, self.out_hw)

    @staticmethod
    def as_real(iq: torch.Tensor) -> torch.Tensor:
        if iq.dim() == 1 or (iq.dim() == 2 and not iq.is_complex() and iq.shape[-1] == 2):
            iq = iq.unsqueeze(0)
        if iq.is_complex():
            if iq.dtype != torch.complex64:
                raise TypeError("IQ bursts must be complex64")
            iq = torch.view_as_real(iq)
        if iq.dtype != torch.float32 or iq.dim() != 3 or iq.shape[-1] != 2:
            raise ValueError(f"IQ bursts must be complex64 [B, L] (or float32 [B, L, 2]); got {tuple(iq.shape)} {iq.dtype}")
        return iq.contiguous()

    def __call__(self, iq_real: torch.Tensor) -> torch.Tensor:
        return ops.iq_to_letterbox(iq_real, self.nfft, self.hop, self.db_min, self.db_max, self.out_hw)


def _classes_tensor(classes, device) -> Optional[torch.Tensor]:
    """`classes=` of predict (int, or list of ints; ops.py:294-295) as an int32 device tensor."""
    if classes is None:
        return None
    if isinstance(classes, (int, np.integer)):
        classes = [int(classes)]
    return torch.tensor([int(c) for c in classes], device=device, dtype=torch.int32)


def _classes_key(classes):
    if classes is None:
        return None
    return (int(classes),) if isinstance(classes, (int, np.integer)) else tuple(int(c) for c in classes)


class DetectionPredictor:
    """preprocess -> inference -> postprocess (engine/predictor.py:221-306), all on the device."""

    def __init__(self, model: DetectionModel, overrides: Optional[dict] = None, front: Optional[IQFrontEnd] = None):
        self.model = model
        self.args = dict(conf=0.25, iou=0.7, max_det=300, agnostic_nms=False, classes=None, imgsz=640, half=False,
                         use_graph=True, stream_slots=3)
        self.args.update(overrides or {})
        self.front = front            # device-side front end (IQ -> image) run inside the captured step
        self._graphs: Dict[tuple, _GraphStep] = {}
        self._stream_steps: List[Optional[_GraphStep]] = [None, None]   # double-buffered graph instances
        self._stream_host: list = [None, None]                           # their pinned host result buffers
        self._compute_streams: list = []                                 # one compute stream per instance
        self.last_launches = 0

    # -- sources ---------------------------------------------------------------------------------
    def preprocess(self, source) -> (torch.Tensor, list, list):
        """Returns (device tensor [B,3,H,W] (fp32/bf16 in 0..1 or uint8), original shapes, host images)."""
        dev = next(self.model.parameters()).device
        if self.front is not None:
            if not isinstance(source, torch.Tensor):
                raise TypeError("the IQ front end takes complex64 tensors [B, L]")
            iq = IQFrontEnd.as_real(source).to(dev, non_blocking=True)
            return iq, [self.front.out_hw] * iq.shape[0], [None] * iq.shape[0]
        if isinstance(source, torch.Tensor):
            # LoadTensor._single_check (data/loaders.py:548-566): BCHW, stride-32 sizes
            im = source if source.dim() == 4 else source.unsqueeze(0)
            if im.dim() != 4 or im.shape[1] != 3:
                raise ValueError(f"WARNING ⚠️ torch.Tensor inputs should be BCHW i.e. shape(1, 3, 640, 640) "
                                 f"divisible by stride 32. Input shape{tuple(im.shape)} is incompatible.")
            if im.shape[2] % 32 or im.shape[3] % 32:
                raise ValueError(f"WARNING ⚠️ torch.Tensor inputs should be BCHW i.e. shape(1, 3, 640, 640) "
                                 f"divisible by stride 32. Input shape{tuple(im.shape)} is incompatible.")
            im = im.to(dev, non_blocking=True)
            return im, [tuple(im.shape[2:])] * im.shape[0], [None] * im.shape[0]
        if isinstance(source, np.ndarray):
            source = [source]
        if isinstance(source, (list, tuple)) and all(isinstance(s, np.ndarray) for s in source):
            # HWC BGR uint8 images of any size: pre_transform (predictor.py:147-163) = LetterBox(imgsz, auto=all images
            # share one shape, stride) per image, then BGR->RGB, HWC->CHW (predictor.py:129-131) — one kernel per shape
            from .data import LetterBox

            if any(s.ndim != 3 or s.shape[2] != 3 or s.dtype != np.uint8 for s in source):
                raise ValueError("ndarray sources must be HWC uint8 images with 3 channels")
            imgsz = self.args.get("imgsz", 640)
            same = len({s.shape for s in source}) == 1
            lb = LetterBox(imgsz, auto=same, stride=int(self.model.stride.max()))
            if same:
                batch = torch.from_numpy(np.ascontiguousarray(np.stack(source))).to(dev, non_blocking=True)
                im = lb.to_network_input(batch)
            else:
                im = torch.empty((len(source), 3) + tuple(lb.new_shape), device=dev, dtype=torch.uint8)
                for i, s in enumerate(source):
                    lb.to_network_input(torch.from_numpy(np.ascontiguousarray(s)).to(dev, non_blocking=True)[None], out=im[i:i + 1])
            return im, [tuple(s.shape[:2]) for s in source], list(source)
        if isinstance(source, (list, tuple)) and source and all(
                isinstance(s, torch.Tensor) and s.dim() == 3 and s.shape[2] == 3 and s.dtype == torch.uint8 for s in source):
            # HWC BGR uint8 images already in device memory (specyolo.data.LoadImagesAndVideos: nvJPEG-decoded files)
            from .data import LetterBox

            imgsz = self.args.get("imgsz", 640)
            same = len({tuple(s.shape) for s in source}) == 1
            lb = LetterBox(imgsz, auto=same, stride=int(self.model.stride.max()))
            src = [s.to(dev, non_blocking=True).contiguous() for s in source]
            if same:
                im = lb.to_network_input(torch.stack(src))
            else:
                im = torch.empty((len(src), 3) + tuple(lb.new_shape), device=dev, dtype=torch.uint8)
                for i, s_ in enumerate(src):
                    lb.to_network_input(s_[None], out=im[i:i + 1])
            return im, [tuple(s.shape[:2]) for s in source], list(source)
        raise TypeError(f"unsupported source type {type(source)}")   # data/build.py:183

    # -- one batch -----------------------------------------------------------------------------
    @torch.no_grad()
    def infer(self, im: torch.Tensor):
        a = self.args
        classes = _classes_tensor(a["classes"], im.device)
        if a["use_graph"]:
            key = (tuple(im.shape), im.dtype, a["conf"], a["iou"], a["agnostic_nms"], a["max_det"],
                   _classes_key(a["classes"]), None if self.front is None else self.front.key())
            step = self._graphs.get(key)
            if step is None:
                step = self._graphs[key] = _GraphStep(self.model, im, a["conf"], a["iou"], a["agnostic_nms"],
                                                      a["max_det"], classes, front=self.front)
            self.last_launches = step.launches
            return step.run(im)
        n0 = _lib.load().specyolo_launch_count()
        if self.front is not None:
            im = self.front(im)
        r = self.model.detect_fused(im, conf_thres=a["conf"], iou_thres=a["iou"], agnostic=a["agnostic_nms"],
                                    max_det=a["max_det"], classes=classes)
        self.last_launches = int(_lib.load().specyolo_launch_count() - n0)
        return r

    def _net_hw(self, im: torch.Tensor):
        """(H, W) of the detector's input for the batch `im` (the front end's output size when there is one)."""
        return tuple(im.shape[2:]) if self.front is None else self.front.out_hw

    def _streams(self, dev, n: int = 2):
        while len(self._compute_streams) < n:
            self._compute_streams.append(torch.cuda.Stream(device=dev))
        return self._compute_streams

    def _ensure_stream_step(self, slot: int, im: torch.Tensor, classes):
        a = self.args
        steps, host_out = self._stream_steps, self._stream_host
        if steps[slot] is None or steps[slot].static_in.shape != im.shape or steps[slot].static_in.dtype != im.dtype:
            example = im.to(next(self.model.parameters()).device) if not im.is_cuda else im
            steps[slot] = _GraphStep(self.model, example, a["conf"], a["iou"], a["agnostic_nms"], a["max_det"], classes,
                                     front=self.front)
            host_out[slot] = (torch.empty(steps[slot].out.shape, dtype=torch.float32).pin_memory(),
                              torch.empty(steps[slot].cnt.shape, dtype=torch.int32).pin_memory())
            self.last_launches = steps[slot].launches
        return steps[slot]

    @torch.no_grad()
    def infer_pipelined(self, im: torch.Tensor, steps: int, inflight: int = 3):
        """`steps` forward + NMS passes over the device-resident batch `im` with `inflight` batches in flight: that many
        captured graph instances replay in turn on their own compute streams, so the low-occupancy phases of one step (the 20x20
        level, decode, NMS: grids far below 148 SMs) overlap the wide kernels of the other.  Returns the (out, cnt)
        device tensors of the last step of each instance.  Everything is enqueued behind the caller's current
        stream and joined back into it."""
        classes = _classes_tensor(self.args["classes"], im.device)
        main = torch.cuda.current_stream(im.device)
        cs = self._streams(im.device, inflight)
        while len(self._stream_steps) < inflight:
            self._stream_steps.append(None)
            self._stream_host.append(None)
        inst = [self._ensure_stream_step(s, im, classes) for s in range(inflight)]
        for s in range(inflight):
            if inst[s].static_in.data_ptr() != im.data_ptr():
                inst[s].static_in.copy_(im, non_blocking=True)      # once, outside the per-step work
            cs[s].wait_stream(main)
        for i in range(steps):
            with torch.cuda.stream(cs[i % inflight]):
                inst[i % inflight].graph.replay()
        for s in range(inflight):
            main.wait_stream(cs[s])
        return [(g.out, g.cnt) for g in inst]

    @torch.no_grad()
    def stream(self, batches):
        """`predict(source, stream=True)`: generator over an iterable of batches (engine/predictor.py:169-175 yields per
        batch too).  `stream_slots` (default 3) captured graph instances take the batches in turn, each on its own
        stream: the host->device copy of batch i+1 (copy stream) overlaps the forward + NMS of the batches in
        flight, and the device->host copy of a [B, max_det, 6] result lands in pinned memory while later batches are
        already running.  Yields List[Results] per batch, in order."""
        a = self.args
        dev = next(self.model.parameters()).device
        n = max(2, int(a.get("stream_slots", 3)))                        # graph instances = batches in flight
        copy_stream = torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream(dev)
        cs = self._streams(dev, n)                                       # one compute stream per graph instance
        for st in cs[:n]:
            st.wait_stream(main)
        while len(self._stream_steps) < n:
            self._stream_steps.append(None)
            self._stream_host.append(None)
        steps, host_out = self._stream_steps, self._stream_host          # captured once, reused across calls
        uploaded = [None] * n
        last_done = [None] * n       # event: the slot's most recent launch (graph + result copies) has completed
        consumed = [None] * n        # event: the step has copied its staged input into the graph's static input
        staging: list = [None] * n   # device staging buffers the host uploads into (decoupled from the graph inputs:
                                     # the upload of batch i+n may start as soon as step i has STARTED, not finished)
        classes = _classes_tensor(a["classes"], dev)

        def prep(item):
            if self.front is not None:      # IQ bursts: stay where they are (pinned host): the copy stream uploads them
                iq = IQFrontEnd.as_real(item)
                return iq, [self.front.out_hw] * iq.shape[0], [None] * iq.shape[0]
            if isinstance(item, torch.Tensor):
                im = item if item.dim() == 4 else item.unsqueeze(0)
                if im.dim() != 4 or im.shape[1] != 3 or im.shape[2] % 32 or im.shape[3] % 32:
                    raise ValueError(f"WARNING ⚠️ torch.Tensor inputs should be BCHW i.e. shape(1, 3, 640, 640) "
                                     f"divisible by stride 32. Input shape{tuple(im.shape)} is incompatible.")
                return im, [tuple(im.shape[2:])] * im.shape[0], [None] * im.shape[0]
            im, shapes, imgs = self.preprocess(item)
            return im, shapes, imgs

        def enqueue_upload(slot, item):
            """Start the host->device copy of `item` into the slot's staging buffer; returns the batch's metadata."""
            im, shapes, imgs = prep(item)
            st_old = steps[slot]
            if st_old is not None and last_done[slot] is not None and \
                    (st_old.static_in.shape != im.shape or st_old.static_in.dtype != im.dtype):
                last_done[slot].synchronize()       # the slot is about to be re-captured for a new input signature:
                                                    # its previous batch must have left the old graph's buffers
            self._ensure_stream_step(slot, im, classes)
            if staging[slot] is None or staging[slot].shape != im.shape or staging[slot].dtype != im.dtype:
                staging[slot] = torch.empty_like(steps[slot].static_in)
                staging[slot].record_stream(copy_stream)      # written on the copy stream, read on the slot's stream:
                staging[slot].record_stream(cs[slot])         # the allocator must not recycle it under either
            if consumed[slot] is not None:
                copy_stream.wait_event(consumed[slot])      # the previous step of this slot has taken its input
            if im.is_cuda:
                # device-side preprocess (LetterBox) produced `im` on the current stream; it is read on the copy stream
                # after this function has dropped its reference
                copy_stream.wait_stream(torch.cuda.current_stream(dev))
                im.record_stream(copy_stream)
            with torch.cuda.stream(copy_stream):
                staging[slot].copy_(im, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            uploaded[slot] = ev
            return (shapes, imgs, im.shape[0], self._net_hw(im))

        def launch(slot, meta):
            # every instance runs on its own stream: consecutive batches overlap on the GPU
            with torch.cuda.stream(cs[slot]):
                cs[slot].wait_event(uploaded[slot])
                steps[slot].static_in.copy_(staging[slot], non_blocking=True)
                consumed[slot] = torch.cuda.Event()
                consumed[slot].record(cs[slot])
                steps[slot].graph.replay()
                host_out[slot][0].copy_(steps[slot].out, non_blocking=True)
                host_out[slot][1].copy_(steps[slot].cnt, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(cs[slot])
            last_done[slot] = ev
            # the metadata and the pinned result buffers travel with the launch: the slot may be re-armed (even
            # re-captured with new buffers) before this batch is collected
            return (host_out[slot], ev, meta)

        def collect(rec):
            (h_out, h_cnt), ev, (shapes, imgs, B, img1) = rec
            ev.synchronize()
            counts = h_cnt.tolist()
            out_h = h_out.clone()                   # one copy out of the pinned buffer; per-image rows are views of it
            for b in range(B):                      # letterboxed sources: boxes back to original-image coordinates
                if tuple(shapes[b]) != tuple(img1):
                    from .utils.ops import scale_boxes
                    scale_boxes(img1, out_h[b, : counts[b], :4], shapes[b])
            return [Results(shapes[b], out_h[b, : counts[b]], self.model.names, orig_img=imgs[b])
                    for b in range(B)]

        it = iter(batches)
        try:
            first = next(it)
        except StopIteration:
            return
        meta_next = enqueue_upload(0, first)
        inflight = collections.deque()
        i = 0
        while True:
            slot = i % n
            inflight.append(launch(slot, meta_next))
            nxt = next(it, None)
            if nxt is not None:
                meta_next = enqueue_upload((i + 1) % n, nxt)   # overlaps the graphs already launched
            # batch i - (n - 1) is collected here, i.e. before launch(i + 1) re-uses its slot's pinned result buffers
            if len(inflight) > n - 1:
                yield collect(inflight.popleft())
            if nxt is None:
                break
            i += 1
        while inflight:
            yield collect(inflight.popleft())
        for st in cs[:n]:
            main.wait_stream(st)

    def __call__(self, source) -> List[Results]:
        im, shapes, host_imgs = self.preprocess(source)
        out, cnt = self.infer(im)
        # postprocess (detect/predict.py:59-73): boxes back to original-image coordinates
        img1 = self._net_hw(im)
        if any(s != img1 for s in shapes):      # boxes back to original-image coordinates (ops.py:92-127, 335-354)
            out = out.clone()
            for b, s in enumerate(shapes):
                if s != img1:
                    ops.scale_boxes_(out[b:b + 1], cnt[b:b + 1], img1, s)
        out_h = out.cpu()                       # single D2H of [B, max_det, 6]
        counts = cnt.cpu().tolist()
        return [Results(shapes[b], out_h[b, : counts[b]], self.model.names, orig_img=host_imgs[b])
                for b in range(im.shape[0])]


class DetectionValidator:
    """The validation loop of the reference (models/yolo/detect/val.py:93-193, engine/validator.py:224-264) on the device:
    dense forward -> `non_max_suppression(conf=0.001, iou=0.7, multi_label=True, max_det=300)` (the batched NMS kernel
    in its multi-label mode: ~10x the candidates of predict) -> IoU + matching kernel at the ten mAP thresholds ->
    host-side AP integration (numpy, as in the reference).

    `batches` yields the reference's collated dicts: {"img": [B,3,H,W] uint8 | float in 0..1, "cls": [n] or [n,1],
    "bboxes": [n,4] normalised xywh, "batch_idx": [n]} and, from the dataset loader (specyolo.data.YOLODataset), "ori_shape"
    + "ratio_pad" per image: predictions and labels are then compared in original-image coordinates like the reference
    does; without them the images are taken to be at the network size (the scale_boxes round trip is the identity)."""

    def __init__(self, model: DetectionModel, overrides: Optional[dict] = None):
        self.model = model
        self.args = dict(conf=0.001, iou=0.7, max_det=300, agnostic_nms=False, single_cls=False)
        self.args.update(overrides or {})
        self.iouv = [0.5 + 0.05 * i for i in range(10)]            # val.py:72
        self.stats = dict(tp=[], conf=[], pred_cls=[], target_cls=[])
        self.seen = 0
        self.confusion_matrix = None                               # val.py:77 / :166: filled when plots=True
        if self.args.get("plots"):
            from .utils.metrics import ConfusionMatrix

            self.confusion_matrix = ConfusionMatrix(nc=len(model.names), conf=self.args["conf"])

    @torch.no_grad()
    def update(self, batch: dict):
        dev = next(self.model.parameters()).device
        img = batch["img"].to(dev, non_blocking=True)
        B, _, H, W = img.shape
        y, _ = self.model(img)
        a = self.args
        out, cnt, _, _ = ops.nms(prediction=y.contiguous(), B=B, nc=y.shape[1] - 4, A=y.shape[2], conf_thres=a["conf"],
                                 iou_thres=a["iou"], agnostic=bool(a["single_cls"] or a["agnostic_nms"]),
                                 multi_label=True, max_det=a["max_det"])
        if a["single_cls"]:
            out[..., 5] = 0
        self.match(out, cnt, batch, H, W)

    @torch.no_grad()
    def match(self, out: torch.Tensor, cnt: torch.Tensor, batch: dict, H: int, W: int):
        """NMS output [B, max_det, 6] (letterboxed pixels) + counts vs the batch's labels -> stats (val.py:107-193)."""
        B = out.shape[0]
        # labels -> pixel xyxy grouped by image (val.py:107-118)
        cls = batch["cls"].reshape(-1).to(torch.float32)
        bidx = batch["batch_idx"].reshape(-1).to(torch.int64)
        box = batch["bboxes"].to(torch.float32).reshape(-1, 4)
        order = torch.argsort(bidx, stable=True)
        cls, bidx, box = cls[order], bidx[order], box[order]
        scale = torch.tensor([W, H, W, H], dtype=torch.float32)
        xyxy = torch.cat((box[:, :2] - box[:, 2:] / 2, box[:, :2] + box[:, 2:] / 2), 1) * scale
        rp, osh = batch.get("ratio_pad"), batch.get("ori_shape")
        if rp is not None and osh is not None:
            # batches of the dataset loader: predictions and labels go back to original-image coordinates and are clipped
            # there (val.py:107-128: scale_boxes with the recorded ratio_pad on both)
            for b in range(B):
                ops.scale_boxes_(out[b:b + 1], cnt[b:b + 1], (H, W), osh[b], ratio_pad=rp[b])
            for b in range(B):
                m = bidx == b
                if bool(m.any()):
                    (gain, _), (pw, ph) = rp[b]
                    bb = xyxy[m]
                    bb[:, [0, 2]] -= pw
                    bb[:, [1, 3]] -= ph
                    bb /= gain
                    bb[:, [0, 2]] = bb[:, [0, 2]].clamp(0, osh[b][1])
                    bb[:, [1, 3]] = bb[:, [1, 3]].clamp(0, osh[b][0])
                    xyxy[m] = bb
        else:
            # images at the network size: _prepare_pred still ends in clip_boxes (val.py:121-128 -> ops.py:124-127)
            ops.scale_boxes_(out, cnt, (H, W), (H, W))
        per_img = torch.bincount(bidx, minlength=B)
        off = torch.zeros(B + 1, dtype=torch.int32)
        off[1:] = torch.cumsum(per_img, 0)
        labels = torch.cat((cls[:, None], xyxy), 1)
        correct = ops.match_predictions(out, cnt, labels, off, int(per_img.max()) if len(per_img) else 0, self.iouv)
        counts = cnt.tolist()
        out_h, corr_h = out.cpu(), correct.cpu()
        for b in range(B):
            self.seen += 1
            n = counts[b]
            if n == 0 and per_img[b] == 0:
                continue
            self.stats["tp"].append(corr_h[b, :n].numpy())
            self.stats["conf"].append(out_h[b, :n, 4].numpy())
            self.stats["pred_cls"].append(out_h[b, :n, 5].numpy())
            self.stats["target_cls"].append(cls[off[b]:off[b + 1]].numpy())
            if self.confusion_matrix is not None:
                lb = labels[off[b]:off[b + 1]].numpy()
                self.confusion_matrix.process_batch(out_h[b, :n].numpy() if n else None, lb[:, 1:], lb[:, 0])

    def results(self) -> dict:
        """`metrics.results_dict` of the reference; the full `DetMetrics` mirror stays on `self.metrics`
        (`.box.maps`, `.class_result(i)`, `.ap_class_index`, ... : val.py:195-211, utils/metrics.py:898-965)."""
        from .utils.metrics import DetMetrics

        self.metrics = DetMetrics(names=self.model.names)
        if self.stats["tp"]:
            st = {k: np.concatenate(v, 0) for k, v in self.stats.items()}
            if len(st["tp"]) and st["tp"].any():
                self.metrics.process(st["tp"], st["conf"], st["pred_cls"], st["target_cls"])
        return {k: float(v) for k, v in self.metrics.results_dict.items()}

    def __call__(self, batches) -> dict:
        for batch in batches:
            self.update(batch)
        return self.results()


class YOLO:
    """`YOLO(cfg_or_weights).predict(source)` (ultralytics/models/yolo/model.py:11, engine/model.py:82-149, 499-558)."""

    def __init__(self, model: Union[str, Path, DetectionModel] = "yolo11s_fusion_sand3_new.yaml", task=None,
                 verbose=False, nc: Optional[int] = None):
        if task not in (None, "detect"):
            raise NotImplementedError(f"task '{task}' is not supported: specyolo implements the detect path only")
        self.task = "detect"
        self.overrides: dict = {}
        self.predictor: Optional[DetectionPredictor] = None
        if isinstance(model, DetectionModel):
            self.model = model
        else:
            p = Path(str(model))
            if p.suffix in (".yaml", ".yml"):
                self.model = DetectionModel(str(p), nc=nc, verbose=verbose)
            elif p.suffix == ".pt":
                # the reference's pickled checkpoints (trainer.save_model) or this package's {'cfg','nc','state_dict'}
                from .nn.checkpoint import attempt_load_one_weight
                self.model, self.ckpt = attempt_load_one_weight(str(p))
                self.overrides["imgsz"] = self.model.args.get("imgsz", 640) if isinstance(self.model.args, dict) else 640
            else:
                raise NotImplementedError(f"unsupported model spec '{model}'")
        self.model.eval()

    @property
    def names(self):
        return self.model.names

    def to(self, device):
        self.model.to(device)
        self.predictor = self._iq_predictor = None
        return self

    def load_state_dict(self, sd, strict=True):
        r = self.model.load_state_dict(sd, strict=strict)
        self.predictor = self._iq_predictor = None
        return r

    def fuse(self):
        self.model.fuse()
        return self

    def predict(self, source=None, stream=False, predictor=None, **kwargs) -> List[Results]:
        if source is None:
            raise ValueError("source is required (the reference's default assets are not shipped)")
        custom = {"conf": 0.25}            # engine/model.py:545
        args = {**self.overrides, **custom, **kwargs}
        dev = args.pop("device", None)
        args.pop("verbose", None)
        if dev is not None:
            if str(dev) == "cpu":
                raise RuntimeError("specyolo has no CPU path; use the reference package for CPU inference")
            self.model.to(torch.device(dev if isinstance(dev, (str, torch.device)) else f"cuda:{dev}"))
        elif not next(self.model.parameters()).is_cuda:
            self.model.to("cuda")
        if self.predictor is None or any(self.predictor.args.get(k) != v for k, v in args.items()):
            self.predictor = (predictor or DetectionPredictor)(self.model, args)
        if isinstance(source, (str, Path)) or (isinstance(source, (list, tuple)) and source and
                                               all(isinstance(s, (str, Path)) for s in source)):
            # image / video files, directory, glob, *.txt list (data/build.py:158-205 -> LoadImagesAndVideos): JPEGs are decoded
            # by nvJPEG straight into device memory (video frames: cv2 on the host, uploaded), letterboxed and run batch by
            # batch through the streaming loop
            from .data import LoadImagesAndVideos

            loader = LoadImagesAndVideos([str(s) for s in source] if isinstance(source, (list, tuple)) else str(source),
                                         batch=int(args.get("batch", 1)), vid_stride=int(args.get("vid_stride", 1)))
            paths: List[str] = []

            def batches():
                for pths, imgs, _ in loader:
                    paths.extend(pths)
                    yield imgs

            def with_paths():
                k = 0
                for res in self.predictor.stream(batches()):
                    for r in res:
                        r.path = paths[k]
                        k += 1
                    yield res

            if stream:
                return with_paths()
            return [r for res in with_paths() for r in res]
        if stream:      # `source` is an iterable of batches; generator of List[Results], copies overlapped with compute
            return self.predictor.stream(source)
        return self.predictor(source)

    __call__ = predict

    def val(self, data=None, validator=None, **kwargs) -> dict:
        """`model.val(...)` (engine/model.py:601-656) over an iterable of collated batches (see DetectionValidator);
        returns the reference's `results_dict` (precision, recall, mAP50, mAP50-95, fitness)."""
        if data is None:
            raise ValueError("data is required: a dataset YAML or an iterable of collated batches")
        if not next(self.model.parameters()).is_cuda:
            self.model.to("cuda")
        if isinstance(data, (str, Path)):
            # dataset YAML -> rectangular validation batches (engine/model.py:639 custom = {"rect": True};
            # validator.get_dataloader -> build_yolo_dataset(mode="val")), decoded and letterboxed on the device
            from .data import build_yolo_dataset, check_det_dataset

            d = check_det_dataset(data)
            split = kwargs.pop("split", "val")
            cfg = {"imgsz": kwargs.pop("imgsz", 640), "rect": kwargs.pop("rect", True),
                   "single_cls": kwargs.get("single_cls", False), "classes": kwargs.pop("classes", None)}
            data = build_yolo_dataset(cfg, d[split], int(kwargs.pop("batch", 16)), d, mode="val",
                                      stride=max(int(self.model.stride.max()), 32))
        self.validator = (validator or DetectionValidator)(self.model, kwargs)     # kept: .confusion_matrix, .stats
        return self.validator(data)

    def predict_iq(self, iq, nfft: int = 1024, hop: int = 256, db_min: float = -100.0, db_max: float = 0.0,
                   stream: bool = False, **kwargs):
        """Raw IQ bursts -> boxes (BASELINE configs[2]).  `iq`: complex64 [B, L] (or float32 [B, L, 2]) — or, with
        `stream=True`, an iterable of such batches (pinned host tensors are uploaded on the copy stream while earlier
        batches compute).  The STFT / log / letterbox kernel runs INSIDE the captured step: one CUDA-graph replay per batch
        covers IQ samples -> spectrogram image -> detector -> decode -> NMS, all on the device."""
        if not next(self.model.parameters()).is_cuda:
            self.model.to("cuda")
        imgsz = kwargs.pop("imgsz", 640)
        front = IQFrontEnd(nfft, hop, db_min, db_max, (imgsz, imgsz) if isinstance(imgsz, int) else tuple(imgsz))
        args = {**self.overrides, "conf": 0.25, **kwargs}
        args.pop("device", None), args.pop("verbose", None)
        p = getattr(self, "_iq_predictor", None)
        if p is None or p.front.key() != front.key() or any(p.args.get(k) != v for k, v in args.items()):
            p = self._iq_predictor = DetectionPredictor(self.model, args, front=front)
        if stream:
            return p.stream(iq)
        return p(iq)
