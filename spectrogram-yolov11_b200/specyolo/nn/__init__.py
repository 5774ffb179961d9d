from .modules import (C2PSA, C3, SPPF, Attention, Bottleneck, C2f, C3k, C3k2, Concat, Conv, DDWConv, Detect, DFL,
                      DWConv, Fusion, GCT, PSABlock, WeightedSpatialAttention)
from .tasks import DetectionModel, parse_model, yaml_model_load

__all__ = ["Conv", "DWConv", "DDWConv", "Bottleneck", "C2f", "C3", "C3k", "C3k2", "SPPF", "Attention", "PSABlock",
           "C2PSA", "Concat", "Fusion", "GCT", "WeightedSpatialAttention", "DFL", "Detect", "DetectionModel",
           "parse_model", "yaml_model_load"]
