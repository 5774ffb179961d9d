"""Loading the reference's pickled checkpoints without the reference package.

`trainer.save_model` (ultralytics/engine/trainer.py:512-546) pickles whole `nn.Module` objects: `{"model": None,
"ema": deepcopy(ema).half(), "train_args": {...}, ...}`; `torch_safe_load` / `attempt_load_one_weight`
(ultralytics/nn/tasks.py:846-960) unpickle them, which needs every `ultralytics.*` class to be importable.  Here those
names resolve to inert stand-ins (an `nn.Module` subclass per name, created on demand) that exist only long enough to
hand over what the checkpoint holds — the model YAML dict, the `state_dict`, `names`, `args` — and the detector is
rebuilt from the YAML with this package's own modules (same sub-module names and `state_dict` keys as the reference,
so the load is strict) and repacked for the B200 kernels on first use.

Trust model: the unpickler resolves globals through an ALLOWLIST only — `ultralytics.*` (inert stand-ins), `torch.nn`
module classes, torch's tensor / storage rebuild helpers, `collections.OrderedDict`, numpy array / scalar
reconstructors, a few value types (`pathlib` paths, `datetime`) and data-only builtins.  Any other global
(`os.system`, `builtins.eval`, `subprocess.*`, ...) raises `pickle.UnpicklingError`, so a crafted `.pt` cannot name
a callable outside that list (tests/test_checkpoint.py::test_malicious_pickle_is_refused).
"""
from __future__ import annotations

import pickle
import types
from pathlib import Path
from typing import Dict, Tuple

import torch
import torch.nn as nn

_STUBS: Dict[Tuple[str, str], type] = {}


def _stand_in(module: str, name: str) -> type:
    key = (module, name)
    cls = _STUBS.get(key)
    if cls is None:
        # nn.Module subclass: pickle restores _parameters / _buffers / _modules through nn.Module.__setstate__, so
        # state_dict() works; plain objects (namespaces, loss holders) restore their __dict__ the same way
        cls = type(name, (nn.Module,), {"__module__": module, "forward": lambda self, *a, **k: (_ for _ in ()).throw(
            RuntimeError(f"{module}.{name} is a checkpoint stand-in and cannot run"))})
        _STUBS[key] = cls
    return cls


_SAFE_BUILTINS = {"set", "frozenset", "slice", "range", "complex", "dict", "list", "tuple", "int", "float", "bool",
                  "str", "bytes", "bytearray", "object"}
_SAFE_GLOBALS = {
    "collections": {"OrderedDict", "defaultdict", "deque"},
    "copyreg": {"_reconstructor"},
    "torch._utils": {"_rebuild_tensor", "_rebuild_tensor_v2", "_rebuild_parameter", "_rebuild_parameter_with_state",
                     "_rebuild_qtensor"},
    "torch.serialization": {"_get_layout"},
    "torch._tensor": {"_rebuild_from_type_v2"},
    "numpy.core.multiarray": {"scalar", "_reconstruct"},
    "numpy._core.multiarray": {"scalar", "_reconstruct"},
    "numpy": {"dtype", "ndarray", "float32", "float64", "int32", "int64", "bool_"},
    "pathlib": {"Path", "PosixPath", "PurePosixPath", "WindowsPath", "PureWindowsPath", "PurePath"},
    "datetime": {"datetime", "date", "timedelta"},
    "argparse": {"Namespace"},
    "types": {"SimpleNamespace"},
}


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module == "ultralytics" or module.startswith("ultralytics."):
            return _stand_in(module, name)
        if module in ("builtins", "__builtin__") and name in _SAFE_BUILTINS:
            return super().find_class("builtins", name)
        if name in _SAFE_GLOBALS.get(module, ()):
            return super().find_class(module, name)
        if module == "torch" and (name.endswith("Storage") or name in ("Size", "device", "dtype", "Tensor", "FloatTensor",
                                                                        "HalfTensor", "BFloat16Tensor", "LongTensor")):
            return super().find_class(module, name)
        if module.startswith("torch.nn.modules.") or module == "torch.nn.parameter":
            obj = super().find_class(module, name)
            if isinstance(obj, type) and (issubclass(obj, nn.Module) or module == "torch.nn.parameter"):
                return obj
        raise pickle.UnpicklingError(f"global '{module}.{name}' is not on the checkpoint allowlist "
                                     "(specyolo.nn.checkpoint refuses to resolve arbitrary callables)")


_pickle_module = types.ModuleType("specyolo_checkpoint_pickle")
_pickle_module.Unpickler = _Unpickler
_pickle_module.load = lambda f, **kw: _Unpickler(f, **kw).load()
_pickle_module.loads = lambda b, **kw: _Unpickler(__import__("io").BytesIO(b), **kw).load()
_pickle_module.dumps = pickle.dumps
_pickle_module.dump = pickle.dump
_pickle_module.__name__ = "pickle"          # torch.load only checks for the attributes above


def torch_safe_load(weight):
    """(ckpt dict, file) like ultralytics/nn/tasks.py:846-898: `ultralytics.*` objects inside are stand-ins."""
    file = str(weight)
    if not Path(file).is_file():
        raise FileNotFoundError(f"'{file}' does not exist (weights are never downloaded: no network on this path)")
    ckpt = torch.load(file, map_location="cpu", pickle_module=_pickle_module, weights_only=False)
    if not isinstance(ckpt, dict):
        # a YOLO instance saved with torch.save(model, f) (tasks.py:890-896)
        ckpt = {"model": ckpt.model}
    return ckpt, file


def attempt_load_one_weight(weight, device=None, inplace=True, fuse=False):
    """(model, ckpt) like ultralytics/nn/tasks.py:937-960, with `model` a `specyolo.DetectionModel` carrying the
    checkpoint's weights (EMA preferred, cast to fp32), `names`, `args`, `pt_path`, `task`."""
    from .tasks import DetectionModel

    ckpt, weight = torch_safe_load(weight)
    if isinstance(ckpt, dict) and "state_dict" in ckpt and "cfg" in ckpt:      # this package's own plain format
        model = DetectionModel(ckpt["cfg"], nc=ckpt.get("nc"))
        model.load_state_dict(ckpt["state_dict"])
        src_names = ckpt.get("names")
    else:
        src = ckpt.get("ema") or ckpt.get("model")
        if src is None or not hasattr(src, "yaml") or not isinstance(src.yaml, dict):
            raise TypeError(f"'{weight}' holds no detection model with a YAML description")
        cfg = {k: v for k, v in src.yaml.items()}
        if any(isinstance(layer[2], str) and layer[2] in ("Segment", "Pose", "OBB", "Classify", "WorldDetect")
               for layer in cfg.get("head", [])):
            raise NotImplementedError("specyolo implements the detect task only")
        model = DetectionModel(cfg, ch=cfg.get("ch", 3), nc=cfg.get("nc"))
        sd = {k: (v.float() if torch.is_floating_point(v) else v) for k, v in src.state_dict().items()}
        model.load_state_dict(sd, strict=True)
        src_names = getattr(src, "names", None)
    if src_names:
        model.names = dict(src_names) if isinstance(src_names, dict) else {i: n for i, n in enumerate(src_names)}
    args = ckpt.get("train_args") or {}
    model.args = dict(args) if isinstance(args, dict) else dict(vars(args))
    model.pt_path = weight
    model.task = "detect"
    model.inplace = inplace
    model = model.to(device) if device is not None else model
    model.eval()
    if fuse:
        model.fuse()
    return model, ckpt
