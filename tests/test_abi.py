"""CPU-side checks of the drop-in boundary: the shared library builds, loads and exports exactly the C-ABI
declared in include/specyolo.h; the product path fails loudly without a GPU (no CPU fallback)."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "specyolo.h"


def _declared_symbols():
    txt = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(specyolo_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported(lib):
    syms = _declared_symbols()
    assert len(syms) >= 20
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, f"declared in specyolo.h but not exported: {missing}"


def test_binding_table_matches_header(lib):
    from specyolo import _lib

    assert sorted(_lib.SIGNATURES) == _declared_symbols()


def test_no_torch_types_in_abi():
    txt = HEADER.read_text()
    code = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)          # declarations only, comments stripped
    assert "at::" not in code and "torch" not in code.lower() and "std::" not in code
    assert 'extern "C"' in txt


def test_pure_host_entry_points(lib):
    assert lib.specyolo_version() >= 100
    assert lib.specyolo_conv_npad(128, 1) == 128
    assert lib.specyolo_conv_npad(2, 1) == 16
    assert lib.specyolo_conv_npad(128, 128) == 1
    assert lib.specyolo_conv_npad(7, 2) == -1
    assert lib.specyolo_conv_merge(128, 128, 8, 1, 1, 0, 1) == 4   # 16-channel groups fused to 64-channel K chunks
    assert lib.specyolo_conv_merge(256, 128, 8, 3, 2, 1, 1) == 2   # s2 without matching dilation: per-tap kernel
    assert lib.specyolo_conv_merge(128, 128, 8, 7, 2, 6, 2) == 1   # DDWConv geometry: halo kernel, group by group
    assert lib.specyolo_conv_merge(128, 128, 1, 3, 1, 1, 1) == 1
    assert lib.specyolo_conv_merge(128, 128, 128, 3, 1, 1, 1) == 64  # depthwise 3x3: block-diagonal GEMM on the halo kernel
    assert lib.specyolo_conv_merge(24, 24, 24, 3, 1, 1, 1) == 1      # odd channel counts: CUDA-core depthwise kernel
    assert lib.specyolo_fusion_ws_bytes(3, 2, 40, 40, 128) > 0
    assert lib.specyolo_nms_ws_bytes(2, 2, 8400, 0) > 0
    lib.specyolo_reset_launch_count()
    assert lib.specyolo_launch_count() == 0


def test_argument_errors_are_reported(lib):
    """Invalid arguments come back as status codes + message, before any CUDA call."""
    from specyolo import _lib

    a = _lib.NmsArgs()
    a.B, a.nc, a.A = 1, 2, 10
    a.conf_thres = 1.5
    rc = lib.specyolo_nms(ctypes.byref(a), None)
    assert rc == _lib.ERR_INVALID
    assert b"Invalid Confidence threshold" in lib.specyolo_last_error()
    assert lib.specyolo_conv2d_bias_act(None, None) == _lib.ERR_INVALID


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu(lib):
    import specyolo
    from specyolo import ops
    from specyolo.utils.ops import non_max_suppression

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.iq_to_letterbox(torch.zeros(1, 2048, dtype=torch.complex64))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        non_max_suppression(torch.zeros(1, 6, 10))
    m = specyolo.DetectionModel("yolo11s_fusion_sand3_new.yaml", nc=2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 64, 64))
    with pytest.raises(RuntimeError):
        specyolo.YOLO("yolo11s_fusion_sand3_new.yaml", nc=2).predict(torch.zeros(1, 3, 64, 64), device="cpu")


def test_product_never_imports_oracle():
    pkg = ROOT / "spectrogram-yolov11_b200"
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.h")) + list(pkg.rglob("*.cuh")):
        txt = f.read_text()
        assert "from oracle" not in txt and "import oracle" not in txt, f
