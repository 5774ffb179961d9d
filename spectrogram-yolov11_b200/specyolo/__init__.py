"""specyolo — B200-native (sm_100a) Spectrogram-YOLOv11 inference hot path.

Python host layer over libspecyolo.so (C-ABI, include/specyolo.h).  Mirrors the reference's operator
surface: `YOLO(cfg).predict()`, `specyolo.nn.modules.{Conv, C3k2, SPPF, C2PSA, DDWConv, Fusion, Detect}`,
`specyolo.utils.ops.non_max_suppression`.  PyTorch is used for tensor plumbing only.
"""
from . import _lib, ops  # noqa: F401
from .engine import YOLO, Boxes, DetectionPredictor, DetectionValidator, Results  # noqa: F401
from .nn.tasks import DetectionModel  # noqa: F401

__version__ = "0.1.0"
__all__ = ["YOLO", "DetectionModel", "DetectionPredictor", "DetectionValidator", "Results", "Boxes", "ops"]
