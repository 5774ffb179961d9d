"""Host-side logic that runs without a GPU: YAML -> graph, state_dict compatibility, strides, shard maths,
synthetic-data determinism, the reference arm of bench.py."""
import json
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def test_model_graph_matches_survey():
    import specyolo

    m = specyolo.DetectionModel("yolo11s_fusion_sand3_new.yaml", nc=2)
    assert sum(p.numel() for p in m.parameters()) == 6_824_734           # README "6.8 M" 
    assert m.stride.tolist() == [8.0, 16.0, 32.0]
    assert [type(l).__name__ for l in m.model][-1] == "Detect" and len(m.model) == 29
    assert m.model[-1].f == [21, 24, 27] and m.model[17].f == [-1, 13, 14]
    n = specyolo.DetectionModel("yolo11n.yaml")
    assert sum(p.numel() for p in n.parameters()) == 2_624_080
    s = specyolo.DetectionModel("yolo11s.yaml")
    assert sum(p.numel() for p in s.parameters()) == 9_458_752
    with pytest.raises(FileNotFoundError):
        specyolo.DetectionModel("does_not_exist.yaml")


def test_scale_comes_from_filename():
    from specyolo.nn.tasks import guess_model_scale, yaml_model_load

    assert guess_model_scale("yolo11s_fusion_sand3_new.yaml") == "s"
    assert guess_model_scale("yolo11n.yaml") == "n"
    d = yaml_model_load("yolo11s_fusion_sand3_new.yaml")
    assert d["scale"] == "s" and len(d["backbone"]) == 11 and len(d["head"]) == 18


@pytest.mark.reference
def test_state_dict_keys_match_reference():
    from oracle.ref_loader import import_reference, reference_available

    if not reference_available():
        pytest.skip("needs /root/reference")
    import_reference()
    from ultralytics.nn.tasks import DetectionModel as Ref

    import specyolo

    for mine_cfg, ref_cfg, nc in [("yolo11s_fusion_sand3_new.yaml", "yolo11s_fusion_sand3_new.yaml", 2), ("yolo11n.yaml", "yolo11n.yaml", 80),
                                  ("yolo11s_fusion_sand3_new_convHCA.yaml", "yolo11s_fusion_sand3_new_convHCA.yaml", 2),
                                  ("yolo11s_fusion_sand3_new_OMN.yaml", "yolo11s_fusion_sand3_new_OMN.yaml", 2),
                                  ("yolo11s_fusion_sand3_new_GC.yaml", "yolo11s_fusion_sand3_new_GC.yaml", 2)]:
        r = Ref(f"/root/reference/ultralytics/cfg/models/11/{ref_cfg}", nc=nc, verbose=False).state_dict()
        m = specyolo.DetectionModel(mine_cfg, nc=nc).state_dict()
        assert list(r.keys()) == list(m.keys())
        assert all(r[k].shape == m[k].shape for k in r)


def test_synthetic_data_is_deterministic():
    import specyolo
    from specyolo.nn.init import synth_images, synth_iq, synth_state_dict

    m = specyolo.DetectionModel("yolo11s_fusion_sand3_new.yaml", nc=2)
    a, b = synth_state_dict(m, 3), synth_state_dict(m, 3)
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert not torch.equal(a["model.3.conv.weight"], synth_state_dict(m, 4)["model.3.conv.weight"])
    m.load_state_dict(a)                                                   # strict
    assert torch.equal(synth_images(2, 64, 1), synth_images(2, 64, 1))
    assert synth_images(1, 64, 0, torch.uint8).dtype == torch.uint8
    iq = synth_iq(1, 4096, 2)
    assert iq.dtype == torch.complex64 and torch.equal(iq, synth_iq(1, 4096, 2))


def test_shard_range():
    from specyolo.dist import shard_range

    assert [shard_range(256, r, 8) for r in range(8)] == [(32 * r, 32 * r + 32) for r in range(8)]
    parts = [shard_range(10, r, 4) for r in range(4)]
    assert parts == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert shard_range(3, 3, 4) == (3, 3)                                  # empty shard
    with pytest.raises(ValueError):
        shard_range(8, 4, 4)


def test_nms_wrapper_argument_contract():
    from specyolo.utils.ops import non_max_suppression, xywh2xyxy

    with pytest.raises(AssertionError, match="Invalid Confidence threshold"):
        non_max_suppression(torch.zeros(1, 6, 4), conf_thres=1.5)
    with pytest.raises(AssertionError, match="Invalid IoU"):
        non_max_suppression(torch.zeros(1, 6, 4), iou_thres=-0.1)
    # end2end outputs are plain filtering (ops.py:240-244), no kernel involved
    p = torch.tensor([[[0, 0, 1, 1, 0.9, 1.0], [0, 0, 1, 1, 0.1, 0.0]]])
    out = non_max_suppression(p, 0.25)
    assert out[0].shape == (1, 6)
    assert torch.equal(xywh2xyxy(torch.tensor([[10.0, 10.0, 4.0, 2.0]])), torch.tensor([[8.0, 9.0, 12.0, 11.0]]))


def test_bench_reference_arm_runs_on_cpu():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-batch", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    # "reference" when oracle/_ref (the unmodified reference, oracle/Makefile) is present, else the oracle port
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] in ("reference", "port")
    vendored = (ROOT / "oracle" / "_ref" / "ultralytics" / "__init__.py").is_file()
    assert line["cpu_baseline"]["kind"] == ("reference" if vendored else "port")
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["unit"] == "images/s"


def test_pack_batch_targets_pads_to_static_slots():
    """Host side of the graph-capturable criterion call: targets grouped by image, normalised xywh -> xyxy pixels, padded to
    a fixed number of slots per image (static buffers), counts unchanged."""
    from specyolo.nn.init import synth_det_batch
    from specyolo.utils.loss import pack_batch_targets, pack_targets

    batch = synth_det_batch(4, 320, nc=3, boxes_per_image=5, seed=1)
    keep = torch.tensor([True] * 20)
    keep[[2, 3, 11]] = False                                   # ragged: 3, 5, 4, 5 boxes
    batch = {k: v[keep] for k, v in batch.items()}
    b0, l0, c0, m0 = pack_targets(batch, 4, (320, 320), "cpu")
    assert m0 == 5 and c0.tolist() == [3, 5, 4, 5] and b0.shape == (4, 5, 4)
    b1, l1, c1, m1 = pack_batch_targets(batch, 4, (320, 320), "cpu", max_boxes=9)
    assert m1 == 9 and b1.shape == (4, 9, 4) and l1.shape == (4, 9) and torch.equal(c1, c0)
    assert torch.equal(b1[:, :5], b0) and torch.equal(l1[:, :5], l0) and float(b1[:, 5:].abs().sum()) == 0.0
    xywh = batch["bboxes"][batch["batch_idx"] == 1] * 320
    want = torch.cat((xywh[:, :2] - xywh[:, 2:] / 2, xywh[:, :2] + xywh[:, 2:] / 2), 1)
    assert torch.allclose(b1[1, :5], want)
    b2, _, _, m2 = pack_batch_targets(batch, 4, (320, 320), "cpu", max_boxes=3)      # never shrinks below the data
    assert m2 == 5 and torch.equal(b2, b0)
