"""TEST INFRASTRUCTURE — imports the upstream reference package.

Two locations, in this order: `oracle/_ref/` (the unmodified reference installed by `make -C oracle`, git-ignored but
shipped to the GPU box with the snapshot) and `/root/reference` (build container only; nothing on the GPU box reads it).
It exists to (a) validate the restatements in oracle/*.py against the real reference, (b) generate the committed
fixtures in tests/golden/ (oracle/gen_golden.py), (c) time the reference as itself in bench.py's CPU / library-bar
legs and (d) host the drop-in test of specyolo.ultralytics_shim.

The reference imports three packages that are absent from this image and cannot be installed
(no network): matplotlib, thop, timm.  None of them is executed on the hot path, so they are
replaced by inert stubs before the import.
"""
from __future__ import annotations

import importlib.machinery
import os
import sys
import tempfile
import types

_VENDORED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REFERENCE_ROOT = os.environ.get("SPECYOLO_REFERENCE_ROOT") or (
    _VENDORED if os.path.isfile(os.path.join(_VENDORED, "ultralytics", "__init__.py")) else "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "ultralytics"))


def _lenient(factory):
    """module-level __getattr__ that never answers dunder lookups (inspect.getmodule probes __file__)."""
    def _ga(k):
        if k.startswith("__"):
            raise AttributeError(k)
        return factory()
    return _ga


def _stub(name: str, attrs: dict | None = None, package: bool = True) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None, is_package=package)
    if package:
        m.__path__ = []
    for k, v in (attrs or {}).items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def _install_stubs() -> None:
    import torch.nn as nn

    class _Anything:
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return self

        def __getattr__(self, k):
            if k.startswith("__"):
                raise AttributeError(k)
            return _Anything()

    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except ImportError:
            mpl = _stub("matplotlib", {"use": lambda *a, **k: None, "rc": lambda *a, **k: None,
                                       "rcParams": {}, "get_backend": lambda: "agg", "__version__": "0.0"})
            mpl.pyplot = _stub("matplotlib.pyplot", {"__getattr__": _lenient(_Anything)}, package=False)
            mpl.font_manager = _stub("matplotlib.font_manager", {"__getattr__": _lenient(_Anything)}, package=False)
            mpl.colors = _stub("matplotlib.colors", {"__getattr__": _lenient(_Anything)}, package=False)
    if "thop" not in sys.modules:
        try:
            import thop  # noqa: F401
        except ImportError:
            _stub("thop", {"profile": lambda *a, **k: (0.0, 0.0)}, package=False)
    if "timm" not in sys.modules:
        try:
            import timm  # noqa: F401
        except ImportError:
            # symbols the fork imports at module scope (conv.py:2278-2281, 2434); only used by blocks
            # (GlobalContext / Faster_Block) that the target configs never instantiate
            class DropPath(nn.Identity):
                def __init__(self, drop_prob=0.0, *a, **k):
                    super().__init__()

            def _noop(*a, **k):
                return None

            timm = _stub("timm")
            _stub("timm.layers", {"DropPath": DropPath, "trunc_normal_": _noop, "to_2tuple": lambda x: (x, x),
                                  "make_divisible": lambda v, d=8, *a, **k: int(v + d / 2) // d * d,
                                  "get_act_layer": lambda *a, **k: nn.ReLU, "__getattr__": _lenient(lambda: _Anything)})
            _stub("timm.models", {"__getattr__": _lenient(lambda: _Anything)})
            _stub("timm.models.layers", {"DropPath": DropPath, "trunc_normal_": _noop, "to_2tuple": lambda x: (x, x),
                                         "__getattr__": _lenient(lambda: _Anything)})
            for sub in ("create_act", "create_conv2d", "helpers", "mlp", "norm", "drop", "weight_init", "conv_bn_act"):
                _stub(f"timm.layers.{sub}", {"__getattr__": _lenient(lambda: _Anything), "DropPath": DropPath,
                                             "trunc_normal_": _noop}, package=False)
                _stub(f"timm.models.layers.{sub}", {"__getattr__": _lenient(lambda: _Anything)}, package=False)


_ref = None


def import_reference():
    """Returns the imported `ultralytics` package of the reference (CPU, offline)."""
    global _ref
    if _ref is not None:
        return _ref
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    os.environ.setdefault("YOLO_CONFIG_DIR", tempfile.mkdtemp(prefix="specyolo_ref_cfg_"))
    os.environ.setdefault("YOLO_OFFLINE", "1")
    os.environ.setdefault("YOLO_VERBOSE", "False")
    sys.dont_write_bytecode = True
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import ultralytics  # noqa: E402

    _ref = ultralytics
    return ultralytics
