"""Executable drop-in: binds the B200 kernels INTO the reference package, so that the reference's own

    from ultralytics import YOLO
    YOLO("yolo11s_fusion_sand3_new.yaml").predict(x, device=0)

runs libspecyolo (SURVEY 8(b) hook 5).  `install()` rebinds, in the reference's module namespaces
(`ultralytics.nn.tasks` — where `parse_model` resolves YAML names, tasks.py:1074-1080 — and `ultralytics.nn.modules[.conv /
.block / .head]`, where the blocks construct their sub-blocks), the classes

    Conv DWConv DDWConv ConvHCA SobelSpatialAttention MSCSpatialAttention Bottleneck BottleNect C3 C3k C3x C2f C3k2 C3k2GC SPPF Attention PSABlock C2PSA Fusion Detect

to SUBCLASSES of the reference's own classes: constructor, parameters, `state_dict` keys, `fuse()` and every `isinstance`
check stay the reference's, and instances pickle as the reference's classes without this package's caches (a checkpoint the
reference trainer writes with the shim installed loads in a plain reference install); only `forward` changes — CUDA tensors go through the specyolo forward of the same
block (the code in specyolo/nn/modules.py, which touches only attributes the reference classes have), anything else
(the 256 x 256 CPU stride probe of DetectionModel.__init__, tasks.py:365; training mode) falls through to the reference's
PyTorch definition.  `ultralytics.utils.ops.non_max_suppression` is rebound to the NMS kernel for CUDA predictions, and
`BaseModel._predict_once` (tasks.py:161-188) to the specyolo layer interpreter (fused stem, concurrent lateral branches)
for CUDA inputs.  `uninstall()` restores everything.

There is no CPU compute in this package: the fall-through is the reference executing its own code on its own tensors.
"""
from __future__ import annotations

import importlib
from typing import Dict, List, Tuple

import torch
import torch.nn as nn

from . import ops
from .nn import modules as M
from .nn.tasks import DetectionModel as _MyModel
from .utils import ops as my_ops

_SAVED: List[Tuple[object, str, object]] = []
_INSTALLED = False

# reference module (relative to ultralytics.nn.modules) -> class names defined there that get a shim
_WHERE = {
    "conv": ["Conv", "DWConv", "DDWConv", "Fusion", "ConvHCA", "SobelSpatialAttention", "MSCSpatialAttention"],
    "block": ["Bottleneck", "C3", "C3k", "C3x", "C2f", "C3k2", "C3k2GC", "BottleNect", "SPPF", "Attention", "PSABlock", "C2PSA"],
    "head": ["Detect"],
}


# attributes this package caches on module instances (weight packs, folded stencils, cache keys)
_CACHE_ATTRS = ("_packed", "_packed_blocked", "_folded_dw", "_shim_key", "_pe_f32", "_pe_key", "_w18", "_f32", "_specyolo_plan")


def _as_reference_class(cls) -> None:
    """Make instances pickle / deepcopy as the REFERENCE's class: the trainer saves `deepcopy(ema.ema).half()` with
    torch.save (engine/trainer.py:500-530), and a checkpoint written while the shim is installed must load in a plain
    reference install.  The shim class takes the module path and name of the class it derives from (where it is bound
    while installed, so pickle's identity check holds) and drops this package's caches from the pickled state."""
    ref = next(b for b in cls.__mro__[1:] if b.__module__.startswith("ultralytics.") and not getattr(b, "_specyolo_shim_class", False))
    cls._specyolo_shim_class = True
    cls.__module__, cls.__qualname__, cls.__name__ = ref.__module__, ref.__qualname__, ref.__name__

    def __getstate__(self):
        state = self.__dict__.copy()
        for k in _CACHE_ATTRS:
            state.pop(k, None)
        return state

    cls.__getstate__ = __getstate__


def _is_cuda(x) -> bool:
    t = x[0] if isinstance(x, (list, tuple)) else x
    t = getattr(t, "src", t)                      # UpsampledView
    return isinstance(t, torch.Tensor) and t.is_cuda


def _weights_key(mod: nn.Module):
    """Identity of the parameters a cached pack was built from: the reference moves / casts / fuses modules in place
    (model.to(), .half(), BaseModel.fuse()), and none of that goes through specyolo's own invalidation hooks."""
    c = getattr(mod, "conv", mod)
    w = c.weight
    return (w.data_ptr(), M.tensor_version(w), w.dtype, w.device, hasattr(mod, "bn"), None if c.bias is None else c.bias.data_ptr())


def _fresh(mod: nn.Module, attrs):
    key = _weights_key(mod)
    if getattr(mod, "_shim_key", None) != key:
        for a in attrs:
            if hasattr(mod, a):
                setattr(mod, a, None)
        mod._shim_key = key


def _make_conv(ref_cls):
    mine = M.Conv

    class _ShimConv(ref_cls):
        # specyolo.nn.modules.Conv's weight-prep helpers, run on the reference instance
        _bn_terms = mine._bn_terms
        is_depthwise3x3 = mine.is_depthwise3x3
        is_stem = mine.is_stem
        takes_blocked = mine.takes_blocked

        def packed(self):
            _fresh(self, ("_packed", "_packed_blocked", "_folded_dw"))
            return mine.packed(self)

        def packed_from_blocked(self):
            _fresh(self, ("_packed", "_packed_blocked", "_folded_dw"))
            return mine.packed_from_blocked(self)

        def folded_depthwise(self):
            _fresh(self, ("_packed", "_packed_blocked", "_folded_dw"))
            return mine.folded_depthwise(self)

        def forward(self, x, out=None, residual=None):
            if _is_cuda(x) and not self.training:
                return mine.forward(self, x, out, residual)
            return ref_cls.forward(self, x)

        def forward_fuse(self, x, out=None, residual=None):
            if _is_cuda(x) and not self.training:
                return mine.forward(self, x, out, residual)
            return ref_cls.forward_fuse(self, x)

    return _ShimConv


def _make_block(ref_cls, mine_cls, extra=()):
    mine_forward = mine_cls.forward

    def forward(self, x, *a, **k):
        if _is_cuda(x) and not self.training:
            return mine_forward(self, x, *a, **k)
        return ref_cls.forward(self, x)

    ns: Dict[str, object] = {"forward": forward}
    for name in extra:
        ns[name] = getattr(mine_cls, name)
    return type(ref_cls.__name__, (ref_cls,), ns)


def _make_detect(ref_cls):
    mine = M.Detect

    class _ShimDetect(ref_cls):
        head_logits = mine.head_logits

        def forward(self, x):
            if _is_cuda(x) and not self.training and not self.end2end and not self.export:
                return mine.forward(self, x)        # (the plain nn.Conv2d tails are packed outside the modules: _TAIL_CACHE)
            return ref_cls.forward(self, x)

    return _ShimDetect


def _make_attention(ref_cls):
    mine = M.Attention

    class _ShimAttention(ref_cls):
        def _pe_weights(self):
            _fresh(self.pe, ("_folded_dw",))
            if getattr(self, "_pe_key", None) != self.pe._shim_key:
                self._pe_f32, self._pe_key = None, self.pe._shim_key
            return mine._pe_weights(self)

        def forward(self, x, out=None, residual=None):
            if _is_cuda(x) and not self.training:
                return mine.forward(self, x, out, residual)
            return ref_cls.forward(self, x)

    return _ShimAttention


def _predict_once_factory(ref_predict_once):
    def _predict_once(self, x, profile=False, visualize=False, embed=None):
        det = self.model[-1]
        if (isinstance(x, torch.Tensor) and x.is_cuda and not self.training and not profile and not visualize
                and embed is None and type(det).__name__ == "Detect" and getattr(det, "_specyolo_shim", False)):
            if x.dtype == torch.float16:
                x = x.float()
            feats = _MyModel._run_trunk(self, x)        # specyolo's interpreter over the reference's layer list
            return det(feats)
        return ref_predict_once(self, x, profile, visualize, embed)
    return _predict_once


def _nms_factory(ref_nms):
    def non_max_suppression(prediction, *args, **kwargs):
        p = prediction[0] if isinstance(prediction, (list, tuple)) else prediction
        rotated, end2end = kwargs.get("rotated", False), kwargs.get("end2end", False)
        if isinstance(p, torch.Tensor) and p.is_cuda and not rotated and not end2end and not kwargs.get("labels"):
            return my_ops.non_max_suppression(prediction, *args, **kwargs)
        return ref_nms(prediction, *args, **kwargs)
    non_max_suppression.__doc__ = ref_nms.__doc__
    return non_max_suppression


def _make_criterion(ref_cls):
    from .utils.loss import criterion_call

    class v8DetectionLoss(ref_cls):
        """The reference's v8DetectionLoss (utils/loss.py:163-275) with the CUDA criterion behind `__call__` for CUDA
        head maps (assigner + losses + gradients in six launches, csrc/det_loss.cu); CPU tensors run the reference."""

        def __call__(self, preds, batch):
            feats = preds[1] if isinstance(preds, tuple) else preds
            if feats[0].is_cuda and type(self.assigner).__name__ == "TaskAlignedAssigner":
                return criterion_call(self, preds, batch)
            return ref_cls.__call__(self, preds, batch)

    return v8DetectionLoss


def _make_ema(ref_cls):
    from .utils.torch_utils import _de_parallel, ema_update

    class ModelEMA(ref_cls):
        """The reference's ModelEMA (utils/torch_utils.py:495-530) with `update` as one CUDA launch (csrc/ema.cu),
        bit-identical to its Python loop; CPU / fp16 entries keep the reference arithmetic."""

        def update(self, model):
            if self.enabled:
                self.updates += 1
                d = self.decay(self.updates)
                self._specyolo_plan = ema_update(self.ema.state_dict, _de_parallel(model).state_dict(), d,
                                                 getattr(self, "_specyolo_plan", None))

    return ModelEMA


def _set(obj, name, value):
    _SAVED.append((obj, name, getattr(obj, name)))
    setattr(obj, name, value)


def install() -> dict:
    """Rebind the reference's blocks / NMS / interpreter to the B200 kernels (idempotent).  Returns {name: shim class}."""
    global _INSTALLED
    import ultralytics                                   # the REFERENCE package must be importable by the caller
    tasks = importlib.import_module("ultralytics.nn.tasks")
    mods = importlib.import_module("ultralytics.nn.modules")
    sub = {k: importlib.import_module(f"ultralytics.nn.modules.{k}") for k in _WHERE}
    uops = importlib.import_module("ultralytics.utils.ops")
    if _INSTALLED:
        return {n: getattr(tasks, n) for names in _WHERE.values() for n in names}
    ops._lib.load()                                      # fail loudly here, not at the first forward, if the .so is missing

    shims: Dict[str, type] = {}
    shims["Conv"] = _make_conv(sub["conv"].Conv)
    # DWConv derives from Conv in the reference: rebuild it on the shim Conv so it inherits the CUDA forward
    ref_dw = sub["conv"].DWConv
    # (MRO DWConv -> shim Conv -> reference DWConv -> reference Conv: the reference's __init__, the shim's forward)
    shims["DWConv"] = type("DWConv", (shims["Conv"], ref_dw), {"__doc__": ref_dw.__doc__})
    shims["DDWConv"] = _make_block(sub["conv"].DDWConv, M.DDWConv)
    shims["Fusion"] = _make_block(sub["conv"].Fusion, M.Fusion)
    shims["SobelSpatialAttention"] = _make_block(sub["conv"].SobelSpatialAttention, M.SobelSpatialAttention, extra=("stencil",))
    shims["ConvHCA"] = _make_block(sub["conv"].ConvHCA, M.ConvHCA)
    shims["MSCSpatialAttention"] = _make_block(sub["conv"].MSCSpatialAttention, M.MSCSpatialAttention, extra=("weights_f32",))
    shims["Bottleneck"] = _make_block(sub["block"].Bottleneck, M.Bottleneck)
    shims["C3"] = _make_block(sub["block"].C3, M.C3)
    shims["C2f"] = _make_block(sub["block"].C2f, M.C2f)
    shims["SPPF"] = _make_block(sub["block"].SPPF, M.SPPF)
    shims["PSABlock"] = _make_block(sub["block"].PSABlock, M.PSABlock)
    shims["C2PSA"] = _make_block(sub["block"].C2PSA, M.C2PSA)
    shims["Attention"] = _make_attention(sub["block"].Attention)
    shims["Detect"] = _make_detect(sub["head"].Detect)
    shims["Detect"]._specyolo_shim = True
    # C3k derives from C3 and C3k2 from C2f in the reference: subclass the reference class AND take the shim forward
    shims["C3k"] = _make_block(sub["block"].C3k, M.C3k)
    shims["C3k2"] = _make_block(sub["block"].C3k2, M.C3k2)
    shims["C3x"] = _make_block(sub["block"].C3x, M.C3x)
    shims["BottleNect"] = _make_block(sub["block"].BottleNect, M.BottleNect, extra=("weights_f32",))
    shims["C3k2GC"] = _make_block(sub["block"].C3k2GC, M.C3k2GC)
    for cls in shims.values():
        _as_reference_class(cls)

    # the blocks build their sub-blocks from their own module's globals; parse_model from tasks' globals
    for space in [tasks, mods, *sub.values()]:
        for name, cls in shims.items():
            if hasattr(space, name):
                _set(space, name, cls)
    _set(uops, "non_max_suppression", _nms_factory(uops.non_max_suppression))
    _set(tasks.BaseModel, "_predict_once", _predict_once_factory(tasks.BaseModel._predict_once))
    # training: DetectionModel.init_criterion (tasks.py:~395) resolves v8DetectionLoss from the tasks namespace
    shims["v8DetectionLoss"] = _make_criterion(tasks.v8DetectionLoss)
    trainer = importlib.import_module("ultralytics.engine.trainer")     # BaseTrainer._setup_train: self.ema = ModelEMA(self.model)
    shims["ModelEMA"] = _make_ema(trainer.ModelEMA)
    for name, spaces in (("v8DetectionLoss", (tasks, importlib.import_module("ultralytics.utils.loss"))),
                         ("ModelEMA", (trainer, importlib.import_module("ultralytics.utils.torch_utils")))):
        _as_reference_class(shims[name])
        for space in spaces:
            _set(space, name, shims[name])
    _INSTALLED = True
    return shims


def uninstall() -> None:
    global _INSTALLED
    while _SAVED:
        obj, name, old = _SAVED.pop()
        setattr(obj, name, old)
    _INSTALLED = False
