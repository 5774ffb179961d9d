"""SURVEY 8 f2, first slice — the detection criterion.  CPU: the oracle restatement (oracle/loss_ref.py) against outputs
of the REAL v8DetectionLoss (tests/golden/loss_cases.npz, oracle/gen_golden.py loss): loss items, total, autograd
gradients with respect to every head map, assigner targets."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
GOLD = ROOT / "tests" / "golden"
CASES = ["mixed", "dense80", "empty", "tiny_boxes"]


def load_case(name):
    from oracle.loss_ref import loss_case

    z = np.load(GOLD / "loss_cases.npz")
    meta = z[f"{name}.meta"]
    seed, B, H, W, nc, dense = (int(v) for v in meta[:6])
    feats, batch = loss_case(seed, B, H, W, nc, [int(v) for v in meta[6:]], bool(dense))
    batch["bboxes"] = torch.from_numpy(z[f"{name}.bboxes"])
    want = {k[len(name) + 1:]: z[k] for k in z.files if k.startswith(name + ".")}
    return feats, batch, nc, want


@pytest.mark.parametrize("name", CASES)
def test_loss_oracle_matches_reference(name):
    from oracle.loss_ref import detection_loss

    feats, batch, nc, want = load_case(name)
    feats = [f.requires_grad_(True) for f in feats]
    total, items, ex = detection_loss(feats, batch, (8.0, 16.0, 32.0), nc)
    total.backward()
    assert np.allclose(items.numpy(), want["items"], rtol=2e-5, atol=1e-6), (items, want["items"])
    assert np.allclose(float(total), float(want["total"]), rtol=2e-5)
    for i, f in enumerate(feats):
        g, w = f.grad.numpy(), want[f"grad{i}"]
        assert np.abs(g - w).max() <= 1e-5 * max(1.0, np.abs(w).max()), (i, np.abs(g - w).max())
    ts, tw = ex["target_scores"].numpy(), want["target_scores"]
    assert np.allclose(ts, tw, rtol=1e-4, atol=1e-6)
    scored = tw.sum(-1) > 0
    assert np.array_equal(ex["fg"].numpy()[scored], want["fg"][scored])
    # the fixture holds the targets after the reference's in-place `target_bboxes /= stride_tensor` (loss.py:266): grid units
    from oracle.loss_ref import make_anchors
    _, stride_t = make_anchors([tuple(f.shape[2:]) for f in feats], (8.0, 16.0, 32.0))
    assert np.allclose((ex["target_boxes"] / stride_t).numpy()[scored], want["target_boxes"][scored], atol=1e-4)
