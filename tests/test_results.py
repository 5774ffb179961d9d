"""Results.verbose / summary / to_json / save_txt and the normalised box views equal the REAL
reference's outputs for the same detections (tests/golden/results_egress.json, oracle/gen_golden.py results)."""
import json
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "spectrogram-yolov11_b200"))
CASES = json.loads((ROOT / "tests" / "golden" / "results_egress.json").read_text())


@pytest.mark.parametrize("tag", ["some", "none"])
def test_egress_formats_match_reference(tag, tmp_path):
    from specyolo.engine import Results

    c = CASES[tag]
    boxes = torch.tensor(c["boxes"], dtype=torch.float32).reshape(-1, 6)
    r = Results(tuple(c["orig_shape"]), boxes, {0: "wifi", 1: "bluetooth"}, path="x.jpg")
    assert r.verbose() == c["verbose"]
    assert r.summary() == c["summary"]
    assert r.summary(normalize=True, decimals=3) == c["summary_norm"]
    assert r.to_json() == c["to_json"] and r.tojson() == c["to_json"]
    assert r.to_csv() == c["to_csv"]
    r.save_txt(tmp_path / "sub" / "a.txt", save_conf=True)
    r.save_txt(tmp_path / "sub" / "b.txt", save_conf=False)
    if len(boxes):
        assert (tmp_path / "sub" / "a.txt").read_text() == c["txt_conf"]
        assert (tmp_path / "sub" / "b.txt").read_text() == c["txt"]
        r.save_txt(tmp_path / "sub" / "b.txt")                       # appended, not overwritten
        assert (tmp_path / "sub" / "b.txt").read_text() == c["txt"] * 2
        assert np.allclose(r.boxes.xywhn.numpy(), np.asarray(c["xywhn"]), rtol=0, atol=1e-7)
        assert np.allclose(r.boxes.xyxyn.numpy(), np.asarray(c["xyxyn"]), rtol=0, atol=1e-7)
    else:
        assert not (tmp_path / "sub").exists()
