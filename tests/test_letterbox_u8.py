"""Image ingest: uint8 LetterBox.  CPU: the numpy restatement of LetterBox + cv2's 8-bit INTER_LINEAR against
outputs of the REAL reference LetterBox (tests/golden/letterbox_u8.npz).  GPU: the kernel against the same fixtures,
bit for bit, and predict() on images of arbitrary size."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
GOLD = ROOT / "tests" / "golden" / "letterbox_u8.npz"
CASES = [  # must mirror oracle/gen_golden.py: LB_CASES
    ((97, 131), (160, 160), {}), ((333, 517), (192, 256), {}), ((120, 90), (160, 160), dict(auto=True)),
    ((320, 256), (160, 160), {}), ((64, 96), (160, 160), dict(scaleup=False)), ((50, 300), (128, 160), dict(auto=True, stride=32)),
    ((160, 160), (160, 160), {}), ((200, 300), (256, 256), dict(center=False)),
]


def test_letterbox_oracle_matches_reference():
    from oracle import letterbox_ref

    z = np.load(GOLD)
    for k, ((h, w), ns, kw) in enumerate(CASES):
        img = z[f"in{k}"]
        assert img.shape == (h, w, 3)
        got = letterbox_ref.letterbox(img, new_shape=ns, **kw)
        assert got.shape == z[f"out{k}"].shape, (k, got.shape)
        assert np.array_equal(got, z[f"out{k}"]), f"case {k}"


@pytest.mark.gpu
def test_letterbox_kernel_bit_exact(lib):
    from specyolo.data import LetterBox

    z = np.load(GOLD)
    for k, (_, ns, kw) in enumerate(CASES):
        kw = dict(kw)
        lb = LetterBox(ns, **kw)
        img = z[f"in{k}"]
        got = lb(image=img)                                    # HWC BGR, like the reference
        assert isinstance(got, np.ndarray) and np.array_equal(got, z[f"out{k}"]), f"case {k}"
        net = lb.to_network_input(torch.from_numpy(img).cuda()[None])       # CHW RGB
        ref = torch.from_numpy(np.ascontiguousarray(z[f"out{k}"][..., ::-1].transpose(2, 0, 1)))
        assert torch.equal(net[0].cpu(), ref)


@pytest.mark.gpu
def test_predict_arbitrary_size_images(lib):
    """ndarray sources of any size: same detections as letterboxing on the host and predicting the tensor, boxes mapped back
    to original-image coordinates with the reference's scale_boxes arithmetic."""
    import specyolo
    from oracle import letterbox_ref
    from specyolo.nn.init import synth_images, synth_state_dict
    from specyolo.utils.ops import scale_boxes

    yolo = specyolo.YOLO("yolo11s_fusion_sand3_new.yaml", nc=2)
    yolo.load_state_dict(synth_state_dict(yolo.model, seed=0))
    yolo.to("cuda")
    base = (synth_images(1, 640, seed=3)[0].permute(1, 2, 0).numpy() * 255).astype(np.uint8)
    img = np.ascontiguousarray(base[:300, :437, ::-1])                     # 300 x 437 BGR
    res = yolo.predict([img], conf=0.1, imgsz=320)
    lb = letterbox_ref.letterbox(img, new_shape=(320, 320), auto=True, stride=32)       # one shape -> auto=True
    x = torch.from_numpy(np.ascontiguousarray(lb[..., ::-1].transpose(2, 0, 1)))[None].cuda()
    ref = yolo.predict(x, conf=0.1, imgsz=320)
    a, b = res[0].boxes.data, ref[0].boxes.data.clone()
    assert len(a) == len(b) and res[0].orig_shape == (300, 437)
    scale_boxes(tuple(x.shape[2:]), b[:, :4], (300, 437))
    assert torch.allclose(a, b, atol=1e-3)
    # two different shapes in one call -> auto=False, full imgsz squares
    img2 = np.ascontiguousarray(base[:200, :200, ::-1])
    r2 = yolo.predict([img, img2], conf=0.1, imgsz=320)
    assert len(r2) == 2 and r2[0].orig_shape == (300, 437) and r2[1].orig_shape == (200, 200)
