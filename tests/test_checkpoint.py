"""the reference's pickled checkpoints load without the reference package.

Fixture: tests/golden/ref_tiny_ckpt.pt, written by the REAL reference the way its trainer does
(oracle/gen_golden.py ckpt; engine/trainer.py:512-546) for a 0.4 M-parameter detector using the fork's modules."""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
GOLD = ROOT / "tests" / "golden"


def test_pickled_reference_checkpoint_loads_without_reference():
    # in a fresh interpreter where `ultralytics` cannot be imported at all
    code = r"""
import sys, importlib.abc
class Block(importlib.abc.MetaPathFinder):
    def find_spec(self, name, path, target=None):
        if name == "ultralytics" or name.startswith("ultralytics."):
            raise ImportError("the reference package is not available here")
sys.meta_path.insert(0, Block())
sys.path.insert(0, %r)
import numpy as np, torch
from specyolo.nn.tasks import attempt_load_one_weight, torch_safe_load
model, ckpt = attempt_load_one_weight(%r)
exp = np.load(%r)
sd = model.state_dict()
assert list(sd.keys()) == [str(k) for k in exp["keys"]], "state_dict keys / order differ"
for k, s, a, n in zip(exp["keys"], exp["sums"], exp["abs_sums"], exp["numel"]):
    v = sd[str(k)]
    assert v.numel() == int(n)
    if torch.is_floating_point(v):
        assert v.dtype == torch.float32
        assert abs(float(v.double().sum()) - float(s)) <= 1e-6 * max(1.0, float(a)), k
        assert abs(float(v.double().abs().sum()) - float(a)) <= 1e-6 * max(1.0, float(a)), k
assert model.names == {0: "wifi", 1: "bluetooth"}
assert model.yaml["nc"] == 2 and model.args["imgsz"] == 640 and model.args["data"] == "Spectrogram.yaml"
assert model.pt_path.endswith("ref_tiny_ckpt.pt") and model.task == "detect" and not model.training
assert ckpt["epoch"] == 7 and ckpt["updates"] == 123 and ckpt["model"] is None
assert [float(s) for s in model.stride] == [8.0, 16.0]
assert "ultralytics" not in sys.modules
print("OK", sum(p.numel() for p in model.parameters()))
""" % (str(ROOT / "spectrogram-yolov11_b200"), str(GOLD / "ref_tiny_ckpt.pt"), str(GOLD / "ref_tiny_ckpt_expect.npz"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.strip().endswith("OK 397588")


def test_missing_file_and_wrong_task(tmp_path):
    sys.path.insert(0, str(ROOT / "spectrogram-yolov11_b200"))
    from specyolo.nn.checkpoint import attempt_load_one_weight

    with pytest.raises(FileNotFoundError):
        attempt_load_one_weight(tmp_path / "nope.pt")
    torch.save({"ema": None, "model": None}, tmp_path / "empty.pt")
    with pytest.raises(TypeError):
        attempt_load_one_weight(tmp_path / "empty.pt")


def test_plain_state_dict_checkpoint_roundtrip(tmp_path):
    sys.path.insert(0, str(ROOT / "spectrogram-yolov11_b200"))
    import specyolo
    from specyolo.nn.checkpoint import attempt_load_one_weight

    m = specyolo.DetectionModel("yolo11n.yaml", nc=3)
    torch.save({"cfg": "yolo11n.yaml", "nc": 3, "state_dict": m.state_dict(), "names": ["a", "b", "c"]}, tmp_path / "w.pt")
    m2, _ = attempt_load_one_weight(tmp_path / "w.pt")
    assert m2.names == {0: "a", 1: "b", 2: "c"}
    for (k1, v1), (k2, v2) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)


@pytest.mark.gpu
def test_checkpoint_model_forward_matches_reference_output():
    """The loaded model runs on the B200 kernels and reproduces the reference's fp32 forward of the same weights."""
    sys.path.insert(0, str(ROOT / "spectrogram-yolov11_b200"))
    import specyolo

    exp = np.load(GOLD / "ref_tiny_ckpt_expect.npz")
    yolo = specyolo.YOLO(str(GOLD / "ref_tiny_ckpt.pt"))
    assert yolo.names == {0: "wifi", 1: "bluetooth"}
    yolo.to("cuda")
    x = torch.from_numpy(exp["x"]).cuda()
    y, _ = yolo.model(x)
    ref = torch.from_numpy(exp["y"])
    got = y.float().cpu()
    assert got.shape == ref.shape
    # boxes in pixels (64 x 96 input), scores in [0, 1]: bf16 operands against the fp32 reference
    assert float((got[:, :4] - ref[:, :4]).abs().max()) < 1.0
    assert float((got[:, 4:] - ref[:, 4:]).abs().max()) < 0.03


def test_malicious_pickle_is_refused(tmp_path):
    """a crafted .pt must not be able to name arbitrary callables (os.system, eval, ...)."""
    import pickle

    sys.path.insert(0, str(ROOT / "spectrogram-yolov11_b200"))
    from specyolo.nn.checkpoint import torch_safe_load

    marker = tmp_path / "pwned"

    class Evil:
        def __reduce__(self):
            import os
            return (os.system, (f"touch {marker}",))

    class Evil2:
        def __reduce__(self):
            return (eval, ("__import__('os').getcwd()",))

    for i, payload in enumerate((Evil(), Evil2())):
        f = tmp_path / f"evil{i}.pt"
        torch.save({"model": payload, "train_args": {}}, f)
        with pytest.raises(pickle.UnpicklingError):
            torch_safe_load(f)
    assert not marker.exists()
