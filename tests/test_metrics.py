"""Validation path.  CPU: the numpy restatement of box_iou / match_predictions and the product's host-side
ap_per_class against fixtures produced by the REAL reference (tests/golden/metrics.npz, oracle/gen_golden.py).
GPU: the matching kernel against the oracle and the fixtures; YOLO.val() end to end."""
from pathlib import Path

import numpy as np
import pytest
import torch

GOLD = Path(__file__).resolve().parent / "golden" / "metrics.npz"
NCASE = 5


def test_match_oracle_matches_reference():
    from oracle import metrics_ref

    z = np.load(GOLD)
    for k in range(NCASE):
        gt, det = z[f"c{k}_gt"], z[f"c{k}_det"]
        iou = metrics_ref.box_iou(gt, det)
        assert iou.shape == z[f"c{k}_iou"].shape
        assert np.allclose(iou, z[f"c{k}_iou"], rtol=0, atol=1e-6)
        corr = metrics_ref.match_predictions(z[f"c{k}_dcls"], z[f"c{k}_gcls"], z[f"c{k}_iou"], z["iouv"].tolist())
        assert np.array_equal(corr, z[f"c{k}_correct"]), f"case {k}"


def test_ap_per_class_matches_reference():
    import sys
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "spectrogram-yolov11_b200"))
    from specyolo.utils.metrics import ap_per_class, results_dict

    z = np.load(GOLD)
    tpc, fpc, p, r, f1, ap, classes = ap_per_class(z["ap_tp"], z["ap_conf"], z["ap_pcls"], z["ap_tcls"])
    assert np.array_equal(classes, z["ap_classes"].astype(int))
    for got, name in ((tpc, "ap_tpc"), (fpc, "ap_fpc"), (p, "ap_p"), (r, "ap_r"), (f1, "ap_f1"), (ap, "ap_ap")):
        assert np.allclose(got, z[name], rtol=1e-9, atol=1e-12), name
    d = results_dict(z["ap_tp"], z["ap_conf"], z["ap_pcls"], z["ap_tcls"])
    assert abs(d["metrics/mAP50(B)"] - z["ap_ap"][:, 0].mean()) < 1e-12
    assert abs(d["metrics/mAP50-95(B)"] - z["ap_ap"].mean()) < 1e-12
    empty = results_dict(np.zeros((0, 10), bool), np.zeros(0), np.zeros(0), np.zeros(0))
    assert empty["metrics/mAP50(B)"] == 0.0


@pytest.mark.gpu
def test_match_kernel_vs_reference_fixture(lib):
    """All five fixture cases batched into one launch: correct[] must equal the reference's bit for bit."""
    from specyolo import ops

    z = np.load(GOLD)
    max_det = 300
    out = torch.zeros((NCASE, max_det, 6))
    cnt = torch.zeros(NCASE, dtype=torch.int32)
    labels, off = [], [0]
    for k in range(NCASE):
        det, n = z[f"c{k}_det"], z[f"c{k}_det"].shape[0]
        out[k, :n, :4] = torch.from_numpy(det)
        out[k, :n, 4] = torch.from_numpy(z[f"c{k}_conf"])
        out[k, :n, 5] = torch.from_numpy(z[f"c{k}_dcls"])
        cnt[k] = n
        labels.append(np.concatenate((z[f"c{k}_gcls"][:, None], z[f"c{k}_gt"]), 1).reshape(-1, 5))
        off.append(off[-1] + labels[-1].shape[0])
    lab = torch.from_numpy(np.concatenate(labels, 0).astype(np.float32))
    corr = ops.match_predictions(out.cuda(), cnt.cuda(), lab, torch.tensor(off, dtype=torch.int32),
                                 max(b - a for a, b in zip(off, off[1:])), z["iouv"].tolist()).cpu().numpy()
    for k in range(NCASE):
        n = int(cnt[k])
        assert np.array_equal(corr[k, :n], z[f"c{k}_correct"]), f"case {k}"
        assert not corr[k, n:].any()


@pytest.mark.gpu
def test_match_kernel_vs_oracle_random(lib):
    from oracle import metrics_ref
    from specyolo import ops

    rng = np.random.default_rng(5)
    B, max_det = 6, 128
    iouv = [0.5 + 0.05 * i for i in range(10)]
    out = torch.zeros((B, max_det, 6)); cnt = torch.zeros(B, dtype=torch.int32)
    labs, off, per = [], [0], []
    for b in range(B):
        nl, nd = int(rng.integers(0, 30)), int(rng.integers(0, max_det + 1))
        gxy = rng.uniform(60, 580, (nl, 2)); gwh = rng.uniform(10, 160, (nl, 2))
        gt = np.concatenate((gxy - gwh / 2, gxy + gwh / 2), 1).astype(np.float32)
        gcls = rng.integers(0, 3, nl).astype(np.float32)
        src = rng.integers(0, max(nl, 1), nd)
        det = (gt[src] + rng.normal(0, 5, (nd, 4))).astype(np.float32) if nl else rng.uniform(0, 640, (nd, 4)).astype(np.float32)
        dcls = (gcls[src] if nl else np.zeros(nd)).astype(np.float32)
        out[b, :nd, :4] = torch.from_numpy(det); out[b, :nd, 5] = torch.from_numpy(dcls)
        out[b, :nd, 4] = torch.linspace(0.9, 0.1, max(nd, 1))[:nd]
        cnt[b] = nd
        labs.append(np.concatenate((gcls[:, None], gt), 1).reshape(-1, 5)); off.append(off[-1] + nl); per.append((gt, gcls, det, dcls))
    lab = torch.from_numpy(np.concatenate(labs, 0).astype(np.float32))
    corr = ops.match_predictions(out.cuda(), cnt.cuda(), lab, torch.tensor(off, dtype=torch.int32),
                                 max(b - a for a, b in zip(off, off[1:])), iouv).cpu().numpy()
    for b, (gt, gcls, det, dcls) in enumerate(per):
        ref = metrics_ref.match_predictions(dcls, gcls, metrics_ref.box_iou(gt, det), iouv)
        assert np.array_equal(corr[b, : det.shape[0]], ref), f"image {b}"


@pytest.mark.gpu
def test_val_end_to_end(lib):
    """YOLO.val(): labels = the model's own confident detections -> mAP50 must be ~1; shuffled labels -> ~0."""
    import specyolo
    from specyolo.nn.init import synth_images, synth_state_dict

    yolo = specyolo.YOLO("yolo11s_fusion_sand3_new.yaml", nc=2)
    yolo.load_state_dict(synth_state_dict(yolo.model, seed=0))
    yolo.to("cuda")
    x = synth_images(4, 320, seed=5)
    res = yolo.predict(x.cuda(), conf=0.25, iou=0.7)
    cls, box, bidx = [], [], []
    for b, r in enumerate(res):
        d = r.boxes.data.cpu()
        xywh = torch.cat(((d[:, :2] + d[:, 2:4]) / 2, d[:, 2:4] - d[:, :2]), 1) / 320.0
        cls.append(d[:, 5]); box.append(xywh); bidx.append(torch.full((len(d),), b))
    batch = {"img": x, "cls": torch.cat(cls), "bboxes": torch.cat(box), "batch_idx": torch.cat(bidx)}
    assert len(batch["cls"]) > 0
    m = yolo.val(data=[batch])
    assert set(m) >= {"metrics/precision(B)", "metrics/recall(B)", "metrics/mAP50(B)", "metrics/mAP50-95(B)", "fitness"}
    assert m["metrics/mAP50(B)"] > 0.9 and m["metrics/recall(B)"] > 0.9, m
    bad = dict(batch, bboxes=batch["bboxes"].flip(0).clone() * 0.37)
    m2 = yolo.val(data=[bad])
    assert m2["metrics/mAP50-95(B)"] < 0.3 * m["metrics/mAP50-95(B)"] + 0.05, (m, m2)


def test_confusion_matrix_matches_reference():
    """ConfusionMatrix.process_batch / tp_fp on the crafted detections of the tiny dataset == the REAL reference's matrix
    (tests/golden/tiny_dataset_batches.npz: inputs and result written by oracle/gen_golden.py dataset)."""
    import sys
    from pathlib import Path

    import numpy as np

    root = Path(__file__).resolve().parent.parent
    sys.path.insert(0, str(root / "spectrogram-yolov11_b200"))
    from specyolo.utils.metrics import ConfusionMatrix, box_iou

    exp = np.load(root / "tests" / "golden" / "tiny_dataset_batches.npz")
    cm = ConfusionMatrix(nc=2, conf=0.001, iou_thres=0.45)
    assert cm.conf == 0.25
    for bi in range(3):
        for si in range(int(exp[f"match_b{bi}_cnt"].shape[0])):
            gt = exp[f"match_b{bi}_gt{si}"]
            cm.process_batch(exp[f"match_b{bi}_predn{si}"], gt[:, 1:], gt[:, 0])
    assert np.array_equal(cm.matrix, exp["confusion_matrix"])
    tp, fp = cm.tp_fp()
    assert np.array_equal(tp, exp["confusion_tp"]) and np.array_equal(fp, exp["confusion_fp"])
    # degenerate inputs (metrics.py:437-448)
    cm2 = ConfusionMatrix(nc=2)
    cm2.process_batch(None, np.zeros((2, 4), np.float32), np.array([0.0, 1.0]))
    cm2.process_batch(np.array([[0, 0, 5, 5, 0.9, 1], [0, 0, 5, 5, 0.1, 0]], np.float32), np.zeros((0, 4), np.float32), np.zeros(0))
    assert cm2.matrix.tolist() == [[0, 0, 0], [0, 0, 1], [1, 1, 0]]
    assert abs(float(box_iou(np.array([[0, 0, 2, 2]]), np.array([[1, 1, 3, 3]]))[0, 0]) - 1 / 7) < 1e-6


def test_detmetrics_matches_reference():
    """DetMetrics / Metric (mean_results, maps, class_result, fitness, map75, results_dict) vs the real reference's."""
    import sys
    from pathlib import Path

    import numpy as np

    root = Path(__file__).resolve().parent.parent
    sys.path.insert(0, str(root / "spectrogram-yolov11_b200"))
    from specyolo.utils.metrics import DetMetrics

    g = np.load(root / "tests" / "golden" / "metrics.npz")
    dm = DetMetrics(names={i: f"c{i}" for i in range(5)})
    assert dm.mean_results() == [0.0, 0.0, 0.0, 0.0] and dm.results_dict["fitness"] == 0.0      # before any data
    dm.process(g["ap_tp"], g["ap_conf"], g["ap_pcls"], g["ap_tcls"])
    assert np.allclose(np.asarray(dm.mean_results()), g["dm_mean"], rtol=0, atol=1e-9)
    assert np.allclose(dm.maps, g["dm_maps"], rtol=0, atol=1e-9) and dm.maps.shape == (5,)
    assert np.allclose(np.asarray(dm.class_result(2)), g["dm_class2"], rtol=0, atol=1e-9)
    assert np.array_equal(np.asarray(dm.ap_class_index), g["dm_index"])
    assert abs(float(dm.fitness) - float(g["dm_fitness"])) < 1e-9 and abs(float(dm.box.map75) - float(g["dm_map75"])) < 1e-9
    assert list(dm.results_dict) == ["metrics/precision(B)", "metrics/recall(B)", "metrics/mAP50(B)", "metrics/mAP50-95(B)", "fitness"]
    assert np.allclose(np.asarray(list(dm.results_dict.values())), g["dm_results"], rtol=0, atol=1e-9)
