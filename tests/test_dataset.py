"""dataset YAML -> validation batches.  Expected values (tests/golden/tiny_dataset_batches.npz) come from
the REAL reference's build_yolo_dataset + build_dataloader on tests/golden/tiny_dataset (oracle/gen_golden.py dataset)."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "spectrogram-yolov11_b200"))
DS = ROOT / "tests" / "golden" / "tiny_dataset"
EXP = ROOT / "tests" / "golden" / "tiny_dataset_batches.npz"


def _dataset(rect):
    from specyolo.data import build_yolo_dataset, check_det_dataset

    data = check_det_dataset(DS / "data.yaml")
    assert data["nc"] == 2 and data["names"] == {0: "wifi", 1: "bluetooth"} and Path(data["val"]).is_dir()
    return build_yolo_dataset({"imgsz": 96, "rect": rect}, data["val"], 3, data, mode="val", stride=32)


@pytest.mark.parametrize("rect", [False, True])
def test_host_side_matches_reference(rect):
    """Order of the images (aspect-ratio sort), batch shapes, pads, ratio_pad and the letterboxed labels."""
    exp = np.load(EXP)
    ds = _dataset(rect)
    nb = int(exp[f"r{int(rect)}_nb"])
    assert len(ds) == nb
    for bi in range(nb):
        p = f"r{int(rect)}_b{bi}_"
        idx = np.nonzero(ds.batch == bi)[0]
        geos = [ds.geometry(int(i)) for i in idx]
        assert [Path(ds.labels[int(i)]["im_file"]).name for i in idx] == [str(f) for f in exp[p + "files"]]
        assert [tuple(g["new_shape"]) for g in geos] == [tuple(exp[p + "img"].shape[2:])] * len(idx)
        assert np.array_equal(np.asarray([g["ori_shape"] for g in geos]), exp[p + "ori_shape"])
        assert np.array_equal(np.asarray([g["ratio_pad"][1] for g in geos]), exp[p + "pad"])
        assert np.allclose(np.asarray([g["ratio_pad"][0] for g in geos]), exp[p + "ratio"], rtol=0, atol=1e-15)
        cls, boxes, bidx = [], [], []
        for k, (i, g) in enumerate(zip(idx, geos)):
            c, b = ds.sample_labels(int(i), g)
            cls.append(c); boxes.append(b); bidx.append(np.full(len(c), k, np.float32))
        assert np.array_equal(np.concatenate(cls).reshape(exp[p + "cls"].shape), exp[p + "cls"])
        assert np.array_equal(np.concatenate(bidx), exp[p + "batch_idx"])
        assert np.allclose(np.concatenate(boxes), exp[p + "bboxes"], rtol=0, atol=2e-7)


def test_dataset_yaml_errors(tmp_path):
    from specyolo.data import check_det_dataset, img2label_paths

    with pytest.raises(FileNotFoundError):
        check_det_dataset(tmp_path / "none.yaml")
    (tmp_path / "bad.yaml").write_text("train: a\nnames: [x]\n")
    with pytest.raises(SyntaxError):
        check_det_dataset(tmp_path / "bad.yaml")
    (tmp_path / "bad2.yaml").write_text("train: a\nval: b\nnc: 3\nnames: [x]\n")
    with pytest.raises(SyntaxError):
        check_det_dataset(tmp_path / "bad2.yaml")
    assert img2label_paths(["/d/images/val/a.png", "/d/images/images/b.jpg"]) == ["/d/labels/val/a.txt", "/d/images/labels/b.txt"]


@pytest.mark.gpu
@pytest.mark.parametrize("rect", [False, True])
def test_batches_match_reference_pixels(rect):
    """Every collated batch — image tensor bit for bit (decode + cv2-exact resize + pad + RGB/CHW on the device), labels,
    bookkeeping — equals what the reference's dataloader yields."""
    exp = np.load(EXP)
    for bi, batch in enumerate(_dataset(rect)):
        p = f"r{int(rect)}_b{bi}_"
        assert batch["img"].is_cuda and batch["img"].dtype == torch.uint8
        assert np.array_equal(batch["img"].cpu().numpy(), exp[p + "img"])
        assert np.array_equal(batch["cls"].numpy(), exp[p + "cls"])
        assert np.allclose(batch["bboxes"].numpy(), exp[p + "bboxes"], rtol=0, atol=2e-7)
        assert np.array_equal(batch["batch_idx"].numpy(), exp[p + "batch_idx"])
        assert [Path(f).name for f in batch["im_file"]] == [str(f) for f in exp[p + "files"]]


@pytest.mark.gpu
def test_val_from_dataset_yaml():
    """YOLO.val(data=yaml) == the validator fed with the same batches by hand; metrics keys of the reference."""
    import specyolo
    from specyolo.nn.init import synth_state_dict

    yolo = specyolo.YOLO("yolo11s_fusion_sand3_new.yaml", nc=2)
    yolo.load_state_dict(synth_state_dict(yolo.model, seed=0))
    yolo.to("cuda")
    r1 = yolo.val(data=str(DS / "data.yaml"), imgsz=96, batch=3)
    r2 = yolo.val(data=list(_dataset(True)))
    assert set(r1) >= {"metrics/precision(B)", "metrics/recall(B)", "metrics/mAP50(B)", "metrics/mAP50-95(B)", "fitness"}
    assert r1 == r2
    r3 = yolo.val(data=str(DS / "data.yaml"), imgsz=96, batch=3, rect=False)
    assert set(r3) == set(r1)


@pytest.mark.gpu
def test_validator_matching_in_original_coordinates():
    """Crafted detections (jittered labels, flipped classes, boxes reaching into the padding) for the rect batches: the
    validator's scale-back + clip + matching reproduces the REAL reference's _prepare_batch / _prepare_pred /
    _process_batch results — the scaled boxes to 1e-4 px and the tp matrices exactly."""
    import specyolo
    from specyolo import ops
    from specyolo.engine import DetectionValidator

    exp = np.load(EXP)
    model = specyolo.DetectionModel("yolo11n.yaml", nc=2).to("cuda")          # only its device is used here
    v = DetectionValidator(model, {"plots": True})          # plots=True: the confusion matrix is accumulated as well
    k = 0
    for bi, batch in enumerate(_dataset(True)):
        preds = torch.from_numpy(exp[f"match_b{bi}_preds"]).cuda()
        cnt = torch.from_numpy(exp[f"match_b{bi}_cnt"]).cuda()
        B, _, H, W = batch["img"].shape
        out = preds.clone()
        v.match(out, cnt, batch, H, W)
        for si in range(B):
            n = int(cnt[si])
            want_tp, want_box = exp[f"match_b{bi}_tp{si}"], exp[f"match_b{bi}_predn{si}"]
            assert np.allclose(out[si, :n].cpu().numpy(), want_box, rtol=0, atol=1e-4)
            got_tp = v.stats["tp"][k]
            k += 1
            assert got_tp.shape == want_tp.shape and np.array_equal(got_tp, want_tp), (bi, si)
    assert any(exp[f"match_b{b}_tp{s}"].any() for b in range(3) for s in range(int(exp[f"match_b{b}_cnt"].shape[0])))
    assert np.array_equal(v.confusion_matrix.matrix, exp["confusion_matrix"])
