"""ctypes binding of libspecyolo.so (include/specyolo.h).

The library is the product: importing this module on a box without the built .so, or calling any
op without a CUDA device, raises — there is no CPU / PyTorch fallback path.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent.parent
_SO = _PKG / "libspecyolo.so"

OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED = 0, 1, 2, 3
ACT_NONE, ACT_SILU = 0, 1
DT_F32, DT_BF16, DT_U8 = 0, 1, 2
DECODE_SEG = 256


class SpecyoloError(RuntimeError):
    """Raised when a libspecyolo entry point returns a non-zero status."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"libspecyolo error {code}: {msg}")
        self.code = code


class ConvArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("Cin", C.c_int),
        ("x_pixstride", C.c_int), ("x_upshift", C.c_int),
        ("w_packed", C.c_void_p), ("bias", C.c_void_p),
        ("Cout", C.c_int), ("n_pad", C.c_int),
        ("kh", C.c_int), ("kw", C.c_int), ("stride", C.c_int), ("pad", C.c_int), ("dil", C.c_int),
        ("groups", C.c_int), ("act", C.c_int),
        ("y", C.c_void_p), ("Ho", C.c_int), ("Wo", C.c_int), ("y_pixstride", C.c_int), ("y_fp32", C.c_int),
        ("residual", C.c_void_p), ("r_pixstride", C.c_int), ("y_s2d", C.c_int),
    ]


class DwpwArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("C", C.c_int), ("x_pixstride", C.c_int),
        ("dw_w", C.c_void_p), ("dw_b", C.c_void_p), ("pw_packed", C.c_void_p), ("pw_bias", C.c_void_p),
        ("Cout", C.c_int), ("n_pad", C.c_int), ("y", C.c_void_p), ("y_pixstride", C.c_int),
        ("head_w", C.c_void_p), ("head_b", C.c_void_p), ("head_y", C.c_void_p), ("head_nc", C.c_int),
        ("head_pixstride", C.c_int),
    ]


class StemPairArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("B", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("w0", C.c_void_p), ("b0", C.c_void_p), ("c0", C.c_int),
        ("w1_packed", C.c_void_p), ("b1", C.c_void_p), ("Cout", C.c_int), ("n_pad", C.c_int),
        ("y", C.c_void_p), ("y_pixstride", C.c_int),
    ]


class BneckArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("C", C.c_int), ("x_pixstride", C.c_int),
        ("w1_packed", C.c_void_p), ("b1", C.c_void_p), ("Cmid", C.c_int), ("n_pad1", C.c_int),
        ("w2_packed", C.c_void_p), ("b2", C.c_void_p), ("Cout", C.c_int), ("n_pad2", C.c_int),
        ("add", C.c_int), ("y", C.c_void_p), ("y_pixstride", C.c_int),
    ]


class FusionArgs(C.Structure):
    _fields_ = [
        ("k", C.c_int), ("x", C.c_void_p * 3), ("pixstride", C.c_int * 3), ("upshift", C.c_int * 3),
        ("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("c", C.c_int),
        ("alpha", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p), ("gct_eps", C.c_float),
        ("sab_w", C.c_void_p), ("y", C.c_void_p), ("y_pixstride", C.c_int), ("ws", C.c_void_p),
    ]


class SpatialGateArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("x_pixstride", C.c_int), ("y", C.c_void_p), ("y_pixstride", C.c_int),
        ("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("C", C.c_int),
        ("w", C.c_float * 18), ("mm", C.c_void_p),
    ]


class MscGateArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("x_pixstride", C.c_int), ("y", C.c_void_p), ("y_pixstride", C.c_int),
        ("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("C", C.c_int),
        ("w_big", C.c_void_p), ("k_big", C.c_int), ("w_small", C.c_void_p), ("fc_w", C.c_void_p), ("fc_b", C.c_void_p),
        ("ws", C.c_void_p),
    ]


class BottleNectArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("x_pixstride", C.c_int), ("y", C.c_void_p), ("y_pixstride", C.c_int),
        ("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("C", C.c_int),
        ("in_w", C.c_void_p), ("in_b", C.c_void_p), ("fac_w", C.c_void_p), ("fac_b", C.c_void_p),
        ("sca_w", C.c_void_p), ("sca_b", C.c_void_p), ("dw1_w", C.c_void_p), ("dw1_b", C.c_void_p),
        ("dw2_w", C.c_void_p), ("dw2_b", C.c_void_p), ("alpha", C.c_void_p), ("beta", C.c_void_p),
        ("ws", C.c_void_p),
    ]


class DetLossArgs(C.Structure):
    _fields_ = [
        ("nl", C.c_int), ("h", C.c_int * 4), ("w", C.c_int * 4), ("stride", C.c_float * 4),
        ("B", C.c_int), ("nc", C.c_int), ("reg_max", C.c_int),
        ("pred_distri", C.c_void_p), ("pred_scores", C.c_void_p),
        ("M", C.c_int), ("gt_boxes", C.c_void_p), ("gt_labels", C.c_void_p), ("gt_count", C.c_void_p),
        ("topk", C.c_int), ("alpha", C.c_float), ("beta", C.c_float), ("tal_eps", C.c_float),
        ("gain_box", C.c_float), ("gain_cls", C.c_float), ("gain_dfl", C.c_float),
        ("out", C.c_void_p), ("grad_distri", C.c_void_p), ("grad_scores", C.c_void_p), ("ws", C.c_void_p),
    ]


class DecodeArgs(C.Structure):
    _fields_ = [
        ("nl", C.c_int), ("logits", C.c_void_p * 4), ("h", C.c_int * 4), ("w", C.c_int * 4),
        ("stride", C.c_float * 4), ("no_stride", C.c_int),
        ("B", C.c_int), ("nc", C.c_int), ("reg_max", C.c_int),
        ("y", C.c_void_p), ("conf_thres", C.c_float), ("cand", C.c_void_p), ("seg_count", C.c_void_p),
    ]


class NmsArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int), ("nc", C.c_int), ("A", C.c_int),
        ("prediction", C.c_void_p), ("cand", C.c_void_p), ("seg_count", C.c_void_p),
        ("conf_thres", C.c_float), ("iou_thres", C.c_double),
        ("agnostic", C.c_int), ("multi_label", C.c_int), ("max_det", C.c_int), ("max_nms", C.c_int),
        ("max_wh", C.c_float), ("classes", C.c_void_p), ("n_classes", C.c_int),
        ("out", C.c_void_p), ("out_count", C.c_void_p), ("keep_idx", C.c_void_p), ("n_cand", C.c_void_p),
        ("ws", C.c_void_p), ("clip_w", C.c_float), ("clip_h", C.c_float),
    ]


class StftArgs(C.Structure):
    _fields_ = [
        ("iq", C.c_void_p), ("B", C.c_int), ("L", C.c_int), ("nfft", C.c_int), ("hop", C.c_int),
        ("db_min", C.c_float), ("db_max", C.c_float), ("out_h", C.c_int), ("out_w", C.c_int),
        ("pad_value", C.c_float), ("out", C.c_void_p), ("out_fp32", C.c_int),
    ]


# name -> (restype, argtypes); also the list the CPU test-suite checks the .so exports against
SIGNATURES = {
    "specyolo_last_error": (C.c_char_p, []),
    "specyolo_version": (C.c_int, []),
    "specyolo_init": (C.c_int, []),
    "specyolo_launch_count": (C.c_uint64, []),
    "specyolo_reset_launch_count": (None, []),
    "specyolo_nchw_to_nhwc_bf16": (C.c_int, [C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int,
                                             C.c_void_p, C.c_int, C.c_void_p]),
    "specyolo_nhwc_bf16_to_nchw_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                                 C.c_void_p, C.c_void_p]),
    "specyolo_sobel_spatial_attention": (C.c_int, [C.POINTER(SpatialGateArgs), C.c_void_p]),
    "specyolo_bottlenect_ws_bytes": (C.c_size_t, [C.c_int] * 4),
    "specyolo_bottlenect": (C.c_int, [C.POINTER(BottleNectArgs), C.c_void_p]),
    "specyolo_msc_ws_bytes": (C.c_size_t, [C.c_int] * 4),
    "specyolo_msc_spatial_attention": (C.c_int, [C.POINTER(MscGateArgs), C.c_void_p]),
    "specyolo_det_loss_ws_bytes": (C.c_size_t, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int]),
    "specyolo_det_loss": (C.c_int, [C.POINTER(DetLossArgs), C.c_void_p]),
    "specyolo_ema_update": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p]),
    "specyolo_upsample2x": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "specyolo_fold_pack_conv": (C.c_int, [C.c_void_p] * 6 + [C.c_float] + [C.c_int] * 7 +
                                [C.c_void_p, C.c_void_p, C.c_void_p]),
    "specyolo_conv_merge": (C.c_int, [C.c_int] * 7),
    "specyolo_conv_npad": (C.c_int, [C.c_int, C.c_int]),
    "specyolo_conv2d_bias_act": (C.c_int, [C.POINTER(ConvArgs), C.c_void_p]),
    "specyolo_dwconv_pwconv": (C.c_int, [C.POINTER(DwpwArgs), C.c_void_p]),
    "specyolo_stem_pair_ok": (C.c_int, [C.c_int] * 5),
    "specyolo_stem_pair": (C.c_int, [C.POINTER(StemPairArgs), C.c_void_p]),
    "specyolo_bottleneck_ok": (C.c_int, [C.c_int] * 5),
    "specyolo_bottleneck": (C.c_int, [C.POINTER(BneckArgs), C.c_void_p]),
    "specyolo_stem_conv3x3s2": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                          C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "specyolo_stem_space_to_depth": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                               C.c_void_p]),
    "specyolo_sppf_pool": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "specyolo_fusion_ws_bytes": (C.c_size_t, [C.c_int] * 5),
    "specyolo_fusion_eschannel": (C.c_int, [C.POINTER(FusionArgs), C.c_void_p]),
    "specyolo_psa_attention": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "specyolo_detect_decode": (C.c_int, [C.POINTER(DecodeArgs), C.c_void_p]),
    "specyolo_nms_ws_bytes": (C.c_size_t, [C.c_int] * 4),
    "specyolo_nms": (C.c_int, [C.POINTER(NmsArgs), C.c_void_p]),
    "specyolo_scale_boxes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_float] * 5 + [C.c_void_p]),
    "specyolo_match_predictions": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                             C.POINTER(C.c_float), C.c_int, C.c_void_p, C.c_void_p]),
    "specyolo_jpeg_info": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "specyolo_jpeg_decode_bgr": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "specyolo_jpeg_decode_batch_bgr": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_void_p),
                                               C.POINTER(C.c_int), C.c_int, C.c_int, C.c_void_p]),
    "specyolo_letterbox_u8": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p] + [C.c_int] * 9 + [C.c_void_p]),
    "specyolo_iq_to_letterbox": (C.c_int, [C.POINTER(StftArgs), C.c_void_p]),
}

_lib = None
_inited_devices: set[int] = set()


def so_path() -> Path:
    return _SO


def load() -> C.CDLL:
    """dlopen libspecyolo.so (building is __graft_entry__.build()'s / build.py's job)."""
    global _lib
    if _lib is not None:
        return _lib
    if not _SO.exists():
        raise ImportError(
            f"{_SO} is missing: build it with `python {_PKG / 'build.py'}` (nvcc, sm_100a). "
            "specyolo has no CPU or PyTorch fallback."
        )
    lib = C.CDLL(os.fspath(_SO))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/.so mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code: int) -> None:
    if code != OK:
        raise SpecyoloError(code, load().specyolo_last_error().decode("utf-8", "replace"))


def init_device() -> None:
    """specyolo_init() once per device; requires a CUDA (sm_100) device."""
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("specyolo needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    dev = torch.cuda.current_device()
    if dev in _inited_devices:
        return
    check(load().specyolo_init())
    _inited_devices.add(dev)


def stream_ptr() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream
