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
    assert np.allclose(float(total.detach()), float(want["total"]), rtol=2e-5)
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


# ---------------------------------------------------------------------------------------------------------------------
# GPU: the CUDA criterion (csrc/det_loss.cu through specyolo.utils.loss.v8DetectionLoss) against the same fixtures of the
# REAL criterion, and against the oracle at a training-size batch.
# ---------------------------------------------------------------------------------------------------------------------
class _FakeDetect:
    def __init__(self, nc):
        self.nc, self.reg_max, self.stride = nc, 16, torch.tensor([8.0, 16.0, 32.0])


class _FakeModel:
    """What v8DetectionLoss.__init__ reads of a model (loss.py:166-186): model[-1] (Detect), args (hyp), a parameter."""

    def __init__(self, nc):
        self.model = [_FakeDetect(nc)]
        self.args = {"box": 7.5, "cls": 0.5, "dfl": 1.5}
        self._p = torch.zeros(1, device="cuda")

    def parameters(self):
        return iter([self._p])


def _cuda_loss(feats, batch, nc):
    from specyolo.utils.loss import v8DetectionLoss

    crit = v8DetectionLoss(_FakeModel(nc))
    f = [x.detach().cuda().requires_grad_(True) for x in feats]
    total, items = crit(f, batch)
    total.backward()
    return total.detach().cpu(), items.cpu(), [x.grad.cpu() for x in f], crit.last_aux.cpu()


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_loss_matches_reference_fixture(lib, name):
    feats, batch, nc, want = load_case(name)
    total, items, grads, aux = _cuda_loss(feats, batch, nc)
    assert np.allclose(items.numpy(), want["items"], rtol=1e-4, atol=1e-5), (items, want["items"])
    assert np.allclose(float(total), float(want["total"]), rtol=1e-4)
    for i, g in enumerate(grads):
        w = want[f"grad{i}"]
        assert np.abs(g.numpy() - w).max() <= 2e-5 * max(1.0, np.abs(w).max()), (i, np.abs(g.numpy() - w).max())
    assert int(aux[1]) == int((want["target_scores"].sum(-1) > 0).sum())        # positives with a non-zero target score


@pytest.mark.gpu
def test_cuda_loss_training_size_vs_oracle(lib):
    """B = 16 at 640^2 (8 400 anchors), 0..14 boxes per image, nc = 2: losses and gradients vs the oracle, run-to-run
    bit-identical results (no floating-point atomics), fp16 head maps (AMP) accepted."""
    from oracle.loss_ref import detection_loss, loss_case

    n_gt = [(7 * b) % 15 for b in range(16)]
    feats, batch = loss_case(11, 16, 640, 640, 2, n_gt)
    fo = [f.clone().requires_grad_(True) for f in feats]
    total_o, items_o, ex = detection_loss(fo, batch, (8.0, 16.0, 32.0), 2)
    total_o.backward()
    total, items, grads, aux = _cuda_loss(feats, batch, 2)
    assert np.allclose(items.numpy(), items_o.numpy(), rtol=1e-4, atol=1e-5), (items, items_o)
    for g, f in zip(grads, fo):
        w = f.grad.numpy()
        assert np.abs(g.numpy() - w).max() <= 2e-5 * max(1.0, np.abs(w).max())
    assert abs(float(aux[0]) - ex["tss"]) <= 1e-4 * ex["tss"]
    total2, items2, grads2, _ = _cuda_loss(feats, batch, 2)
    assert torch.equal(items, items2) and all(torch.equal(a, b) for a, b in zip(grads, grads2))
    th, ih, gh, _ = _cuda_loss([f.half() for f in feats], batch, 2)
    assert gh[0].dtype == torch.float16 and np.allclose(ih.numpy(), items_o.numpy(), rtol=2e-2)


@pytest.mark.gpu
def test_cuda_loss_packed_targets_and_graph_capture(lib):
    """`pack_batch_targets` (static target buffers padded to a fixed number of slots per image): the same losses and gradients
    as the plain batch (the padding changes the reduction layout: last-ulp differences only); with them the criterion makes
    no host round trip, so forward + backward capture into a CUDA graph whose replays are bit-identical to the eager call, and
    refilling the static buffers between replays switches the labels."""
    from oracle.loss_ref import loss_case
    from specyolo.utils.loss import pack_batch_targets, v8DetectionLoss

    B, nc = 8, 2
    feats, batch_a = loss_case(31, B, 320, 320, nc, [(5 * b) % 9 for b in range(B)])
    _, batch_b = loss_case(32, B, 320, 320, nc, [(3 * b + 1) % 7 for b in range(B)])
    crit = v8DetectionLoss(_FakeModel(nc))
    f = [x.detach().cuda().requires_grad_(True) for x in feats]

    def eager(batch):
        for x in f:
            x.grad = None
        total, items = crit(f, batch)
        total.backward()
        return items.cpu(), [x.grad.cpu().clone() for x in f]

    def close(got, want):
        return np.allclose(got[0].numpy(), want[0].numpy(), rtol=1e-5, atol=1e-6) and all(
            float((g - w).abs().max()) <= 1e-5 * max(1.0, float(w.abs().max())) for g, w in zip(got[1], want[1]))

    packed = pack_batch_targets(batch_a, B, (320, 320), "cuda", max_boxes=12)
    assert packed[0].shape == (B, 12, 4) and packed[1].shape == (B, 12) and packed[3] == 12
    static = dict(batch_a, packed_targets=packed)
    plain_a, plain_b = eager(batch_a), eager(batch_b)
    padded_a = eager(static)
    assert close(padded_a, plain_a)

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        eager(static)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    for x in f:
        x.grad = None
    with torch.cuda.graph(graph):
        total_g, items_g = crit(f, static)
        total_g.backward()
    graph.replay()
    assert torch.equal(items_g.cpu(), padded_a[0]) and all(torch.equal(x.grad.cpu(), g) for x, g in zip(f, padded_a[1]))
    nb = pack_batch_targets(batch_b, B, (320, 320), "cuda", max_boxes=12)
    for dst, src in zip(packed[:3], nb[:3]):
        dst.copy_(src)                                             # new labels into the same buffers
    graph.replay()
    assert close((items_g.cpu(), [x.grad.cpu() for x in f]), plain_b)
    assert not close((items_g.cpu(), [x.grad.cpu() for x in f]), plain_a)
