"""TEST INFRASTRUCTURE — generates tests/golden/* by running the REAL reference (/root/reference, CPU).

Run in the build container only:   python oracle/gen_golden.py
The fixtures are small and committed; the GPU box (no /root/reference) checks the oracle restatements and
the CUDA kernels against them.

  model_<cfg>.npz   seeded input + reference DetectionModel outputs (dense prediction, raw head maps, per-layer
                    abs-mean / abs-max fingerprints) for weights = specyolo.nn.init.synth_state_dict(seed) loaded
                    into the reference model (this also pins state_dict key compatibility)
  nms_cases.npz     prediction tensors + outputs of the reference's ops.non_max_suppression
                    (torchvision CPU nms) for single-label, agnostic, multi-label, class-filter, max_det cases
  letterbox.npz     LetterBox(auto=False) geometry + cv2 INTER_LINEAR float resize of a float image
  letterbox_u8.npz  LetterBox (cv2.resize INTER_LINEAR uint8 + copyMakeBorder) on random images, several geometries
  metrics.npz       validation path: box_iou + DetectionValidator.match_predictions on random detections / labels,
                    ap_per_class on a pooled synthetic run
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "spectrogram-yolov11_b200"))

from oracle.ref_loader import REFERENCE_ROOT, import_reference  # noqa: E402

GOLD = ROOT / "tests" / "golden"

MODEL_CASES = [
    # fixture name, reference yaml, our cfg name, nc, input H, W, seed
    ("specyolo_s", "yolo11s_fusion_sand3_new.yaml", "yolo11s_fusion_sand3_new.yaml", 2, 96, 128, 0),
    ("yolo11n", "yolo11n.yaml", "yolo11n.yaml", 80, 64, 96, 1),
    ("specyolo_s_convhca", "yolo11s_fusion_sand3_new_convHCA.yaml", "yolo11s_fusion_sand3_new_convHCA.yaml", 2, 96, 128, 2),
    ("specyolo_s_omn", "yolo11s_fusion_sand3_new_OMN.yaml", "yolo11s_fusion_sand3_new_OMN.yaml", 2, 96, 128, 3),
    ("specyolo_s_gc", "yolo11s_fusion_sand3_new_GC.yaml", "yolo11s_fusion_sand3_new_GC.yaml", 2, 96, 128, 4),
]


def gen_models():
    from ultralytics.nn.tasks import DetectionModel as RefModel

    import specyolo
    from specyolo.nn.init import synth_images, synth_state_dict

    only = {a[len("model="):] for a in sys.argv[1:] if a.startswith("model=")}     # e.g. `models model=specyolo_s_convhca`
    for name, ref_yaml, cfg, nc, H, W, seed in MODEL_CASES:
        if only and name not in only:
            continue
        ref = RefModel(f"{REFERENCE_ROOT}/ultralytics/cfg/models/11/{ref_yaml}", nc=nc, verbose=False).eval()
        mine = specyolo.DetectionModel(cfg, nc=nc)
        sd = synth_state_dict(mine, seed=seed)
        missing, unexpected = ref.load_state_dict(sd, strict=True)
        assert not missing and not unexpected
        x = synth_images(2, 640, seed=seed)[:, :, :H, :W].contiguous()
        feats = {}
        hooks = [m.register_forward_hook(lambda mod, i, o, idx=idx: feats.__setitem__(idx, o))
                 for idx, m in enumerate(ref.model)]
        with torch.no_grad():
            y, raw = ref(x)
        for h in hooks:
            h.remove()
        fp = []
        for idx in range(len(ref.model) - 1):
            o = feats[idx]
            fp.append([float(o.abs().mean()), float(o.abs().max())])
        out = {"x": x.numpy(), "y": y.numpy(), "layer_fingerprint": np.asarray(fp, dtype=np.float64),
               "strides": ref.stride.numpy(), "nc": np.int64(nc), "seed": np.int64(seed)}
        for i, r in enumerate(raw):
            out[f"raw{i}"] = r.numpy()
        np.savez_compressed(GOLD / f"model_{name}.npz", **out)
        print(name, "y", tuple(y.shape), "max score", float(y[:, 4:].max()), "cands>0.25", int((y[:, 4:].amax(1) > 0.25).sum()))


def gen_nms():
    from ultralytics.utils import ops as ref_ops

    rng = np.random.default_rng(7)
    cases = {}

    def make(B, nc, A, scale=1.0):
        xy = rng.uniform(20, 620, (B, 2, A))
        wh = rng.uniform(8, 200, (B, 2, A))
        sc = rng.beta(0.5, 8.0, (B, nc, A)) * scale
        p = np.concatenate((xy, wh, sc), 1).astype(np.float32)
        p[:, :, 1::5] = p[:, :, 0:-1:5][:, :, : p[:, :, 1::5].shape[2]]   # duplicates -> score ties
        return p

    specs = [
        ("single", make(2, 2, 2000), dict(conf_thres=0.25, iou_thres=0.7)),
        ("lowconf", make(2, 2, 1500), dict(conf_thres=0.02, iou_thres=0.45)),
        ("agnostic", make(2, 5, 1000), dict(conf_thres=0.1, iou_thres=0.5, agnostic=True)),
        ("multilabel", make(2, 3, 800), dict(conf_thres=0.05, iou_thres=0.6, multi_label=True)),
        ("classes", make(2, 4, 1000), dict(conf_thres=0.1, iou_thres=0.7, classes=[1, 3])),
        ("maxdet", make(1, 1, 3000, 3.0).clip(0, 1), dict(conf_thres=0.05, iou_thres=0.9, max_det=50)),
        ("empty", make(2, 2, 300, 0.01), dict(conf_thres=0.9, iou_thres=0.7)),
    ]
    for name, pred, kw in specs:
        out = ref_ops.non_max_suppression(torch.from_numpy(pred.copy()), **kw)
        cases[f"{name}_pred"] = pred
        cases[f"{name}_kw"] = np.asarray(repr(kw))
        for b, o in enumerate(out):
            cases[f"{name}_out{b}"] = o.numpy().astype(np.float32)
        print("nms", name, [tuple(o.shape) for o in out])
    np.savez_compressed(GOLD / "nms_cases.npz", **cases)


def gen_letterbox():
    import cv2
    from ultralytics.data.augment import LetterBox

    rng = np.random.default_rng(3)
    shapes = [(1024, 4093), (1024, 253), (700, 900), (480, 640), (1024, 1024), (33, 1000)]
    geo = []
    for (h, w) in shapes:
        lb = LetterBox((640, 640), auto=False, scaleup=True, center=True)
        img = np.zeros((h, w, 3), dtype=np.uint8)
        img[...] = 255
        out = lb(image=img)
        ys, xs = np.nonzero(out[..., 0] == 255)
        geo.append([h, w, xs.min(), ys.min(), xs.max() - xs.min() + 1, ys.max() - ys.min() + 1])
    f = rng.random((1024, 509)).astype(np.float32)
    res = cv2.resize(f, (320, 640), interpolation=cv2.INTER_LINEAR)     # float path of cv2.INTER_LINEAR
    res2 = cv2.resize(f, (80, 160), interpolation=cv2.INTER_LINEAR)
    np.savez_compressed(GOLD / "letterbox.npz", geometry=np.asarray(geo, dtype=np.int64), src=f, up=res, down=res2)
    print("letterbox", geo)


LB_CASES = [  # (h, w), new_shape, kwargs
    ((97, 131), (160, 160), {}), ((333, 517), (192, 256), {}), ((120, 90), (160, 160), dict(auto=True)),
    ((320, 256), (160, 160), {}), ((64, 96), (160, 160), dict(scaleup=False)), ((50, 300), (128, 160), dict(auto=True, stride=32)),
    ((160, 160), (160, 160), {}), ((200, 300), (256, 256), dict(center=False)),
]


def gen_letterbox_u8():
    """Real LetterBox on random uint8 images (cv2.resize INTER_LINEAR + copyMakeBorder)."""
    from ultralytics.data.augment import LetterBox

    rng = np.random.default_rng(21)
    out = {}
    for k, ((h, w), ns, kw) in enumerate(LB_CASES):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        y = LetterBox(ns, **kw)(image=img)
        out[f"in{k}"] = img
        out[f"out{k}"] = y
        print("letterbox_u8", (h, w), ns, kw, "->", y.shape)
    np.savez_compressed(GOLD / "letterbox_u8.npz", **out)


def gen_metrics():
    """Validation path: real box_iou / DetectionValidator.match_predictions / ap_per_class on seeded random cases."""
    from types import SimpleNamespace

    from ultralytics.models.yolo.detect.val import DetectionValidator
    from ultralytics.utils.metrics import ap_per_class, box_iou

    rng = np.random.default_rng(11)
    iouv = torch.linspace(0.5, 0.95, 10)
    out = {"iouv": iouv.numpy()}
    for k, (nd, nl, nc) in enumerate([(300, 40, 2), (57, 9, 3), (5, 0, 2), (0, 4, 2), (200, 120, 80)]):
        gxy = rng.uniform(60, 580, (nl, 2)); gwh = rng.uniform(10, 160, (nl, 2))
        gt = np.concatenate((gxy - gwh / 2, gxy + gwh / 2), 1).astype(np.float32)
        gcls = rng.integers(0, nc, nl).astype(np.float32)
        # detections: jittered copies of labels (several per label) + random boxes, sorted by confidence
        src = rng.integers(0, max(nl, 1), nd)
        det = (gt[src] + rng.normal(0, 6, (nd, 4)).astype(np.float32)) if nl else rng.uniform(0, 640, (nd, 4)).astype(np.float32)
        rnd = rng.random(nd) < 0.3
        det[rnd] = np.sort(rng.uniform(0, 640, (int(rnd.sum()), 4)).astype(np.float32).reshape(-1, 2, 2), 1).reshape(-1, 4)
        dcls = np.where(rng.random(nd) < 0.8, gcls[src] if nl else 0, rng.integers(0, nc, nd)).astype(np.float32)
        conf = np.sort(rng.random(nd).astype(np.float32))[::-1].copy()
        iou = box_iou(torch.from_numpy(gt), torch.from_numpy(det))
        corr = DetectionValidator.match_predictions(SimpleNamespace(iouv=iouv), torch.from_numpy(dcls), torch.from_numpy(gcls), iou)
        out.update({f"c{k}_gt": gt, f"c{k}_gcls": gcls, f"c{k}_det": det, f"c{k}_dcls": dcls, f"c{k}_conf": conf,
                    f"c{k}_iou": iou.numpy(), f"c{k}_correct": corr.numpy()})
        print("metrics case", k, "correct@.5", int(corr[:, 0].sum()) if nd else 0)
    # ap_per_class on a pooled synthetic run
    n, nc = 4000, 5
    conf = rng.random(n).astype(np.float32)
    pcls = rng.integers(0, nc, n).astype(np.float32)
    base = rng.random(n) < (0.2 + 0.7 * conf)
    tp = np.stack([base & (rng.random(n) < 1.0 - 0.08 * j) for j in range(10)], 1)
    tcls = rng.integers(0, nc - 1, 1500).astype(np.float32)          # class nc-1 has predictions but no labels
    r = ap_per_class(tp, conf, pcls, tcls)
    out.update({"ap_tp": tp, "ap_conf": conf, "ap_pcls": pcls, "ap_tcls": tcls, "ap_tpc": r[0], "ap_fpc": r[1], "ap_p": r[2],
                "ap_r": r[3], "ap_f1": r[4], "ap_ap": r[5], "ap_classes": r[6]})
    # DetMetrics on the same pooled run: what validator.metrics reports (names for all 5 classes)
    from ultralytics.utils.metrics import DetMetrics

    dm = DetMetrics(names={i: f"c{i}" for i in range(nc)})
    dm.process(tp, conf, pcls, tcls)
    out.update({"dm_mean": np.asarray(dm.mean_results()), "dm_maps": dm.maps, "dm_fitness": np.float64(dm.fitness),
                "dm_class2": np.asarray(dm.class_result(2)), "dm_index": np.asarray(dm.ap_class_index),
                "dm_map75": np.float64(dm.box.map75), "dm_results": np.asarray(list(dm.results_dict.values()))})
    np.savez_compressed(GOLD / "metrics.npz", **out)
    print("ap_per_class mAP50", float(r[5][:, 0].mean()), "mAP", float(r[5].mean()))


TINY_CFG = {"nc": 2, "scales": {"n": [0.50, 0.25, 1024]}, "scale": "n",
            "backbone": [[-1, 1, "Conv", [64, 3, 2]], [-1, 1, "Conv", [128, 3, 2]], [-1, 2, "C3k2", [256, False, 0.25]],
                         [-1, 1, "Conv", [256, 3, 2]], [-1, 1, "DDWConv", [512, 3, 2, 2]], [-1, 1, "SPPF", [512, 5]],
                         [-1, 1, "C2PSA", [512]]],
            "head": [[[3, 6], 1, "Detect", ["nc"]]]}


def gen_ckpt():
    """A checkpoint written the way the reference's trainer writes it (engine/trainer.py:512-546: pickled EMA module in
    fp16 + train_args) for a 0.4 M-parameter detector that uses the fork's own modules, plus the expected tensors."""
    import io
    from copy import deepcopy
    from datetime import datetime

    from ultralytics import __version__ as ref_version
    from ultralytics.nn.tasks import DetectionModel as RefModel

    torch.manual_seed(11)
    ref = RefModel(deepcopy(TINY_CFG), nc=2, verbose=False)
    with torch.no_grad():
        for m in ref.modules():                      # non-trivial BN statistics, as after training
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.1)
                m.running_var.uniform_(0.5, 1.5)
                m.weight.uniform_(0.5, 1.5)
                m.bias.normal_(0, 0.1)
    ref.names = {0: "wifi", 1: "bluetooth"}
    ref.args = {"imgsz": 640, "batch": 16}
    ema = deepcopy(ref).half()
    ckpt = {"epoch": 7, "best_fitness": 0.5, "model": None, "ema": ema, "updates": 123, "optimizer": None,
            "train_args": {"task": "detect", "mode": "train", "model": "yolo11s_fusion_sand3_new.yaml",
                           "data": "Spectrogram.yaml", "imgsz": 640, "batch": 16, "epochs": 300},
            "train_metrics": {"metrics/mAP50(B)": 0.9, "fitness": 0.5}, "train_results": None,
            "date": datetime(2024, 1, 1).isoformat(), "version": ref_version, "license": "AGPL-3.0", "docs": ""}
    buf = io.BytesIO()
    torch.save(ckpt, buf)
    (GOLD / "ref_tiny_ckpt.pt").write_bytes(buf.getvalue())
    sd = ema.float().eval().state_dict()
    x = torch.rand((1, 3, 64, 96), generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        y, raw = ema.float().eval()(x)
    out = {"keys": np.asarray(list(sd.keys())), "sums": np.asarray([float(v.double().sum()) for v in sd.values()]),
           "abs_sums": np.asarray([float(v.double().abs().sum()) for v in sd.values()]),
           "numel": np.asarray([v.numel() for v in sd.values()], dtype=np.int64), "x": x.numpy(), "y": y.numpy()}
    np.savez_compressed(GOLD / "ref_tiny_ckpt_expect.npz", **out)
    print("ckpt bytes", len(buf.getvalue()), "tensors", len(sd), "y", tuple(y.shape))


def gen_dataset():
    """tests/golden/tiny_dataset (7 PNGs of assorted aspect ratios, YOLO txt labels incl. a missing file, an empty file
    and a duplicated row) and the batches the REAL reference's validation dataloader makes of it (square and rect)."""
    import os
    import shutil

    import cv2
    from ultralytics.cfg import get_cfg
    from ultralytics.data.build import build_dataloader, build_yolo_dataset

    root = GOLD / "tiny_dataset"
    shutil.rmtree(root, ignore_errors=True)
    (root / "images" / "val").mkdir(parents=True)
    (root / "labels" / "val").mkdir(parents=True)
    g = np.random.default_rng(5)
    for k, (h, w) in enumerate([(60, 100), (100, 60), (80, 80), (50, 120), (120, 48), (64, 96), (90, 70)]):
        img = cv2.GaussianBlur(g.integers(0, 255, (h, w, 3), dtype=np.uint8), (5, 5), 0)
        cv2.imwrite(str(root / "images" / "val" / f"im{k}.png"), img)
        n = int(g.integers(0, 4))
        if k == 3:
            continue                                   # no label file: background image
        rows = []
        for _ in range(n):
            cx, cy = g.uniform(0.2, 0.8, 2)
            bw, bh = g.uniform(0.05, 0.3, 2)
            rows.append(f"{int(g.integers(0, 2))} {cx:.6f} {cy:.6f} {bw:.6f} {bh:.6f}")
        if k == 5 and rows:
            rows.append(rows[0])                       # duplicate row
        (root / "labels" / "val" / f"im{k}.txt").write_text("\n".join(rows) + ("\n" if rows else ""))
    (root / "data.yaml").write_text("path: .\ntrain: images/val\nval: images/val\nnames:\n  0: wifi\n  1: bluetooth\n")
    data = {"path": str(root), "val": str(root / "images" / "val"), "train": str(root / "images" / "val"),
            "names": {0: "wifi", 1: "bluetooth"}, "nc": 2}      # what check_det_dataset returns (it also wants fonts)
    out = {}
    for rect in (False, True):
        cfg = get_cfg(overrides=dict(imgsz=96, task="detect", rect=rect, workers=0))
        ds = build_yolo_dataset(cfg, data["val"], 3, data, mode="val", stride=32)
        bi = -1
        for bi, b in enumerate(build_dataloader(ds, 3, 0, shuffle=False, rank=-1)):
            p = f"r{int(rect)}_b{bi}_"
            out[p + "img"] = b["img"].numpy()
            out[p + "cls"] = b["cls"].numpy()
            out[p + "bboxes"] = b["bboxes"].numpy()
            out[p + "batch_idx"] = b["batch_idx"].numpy()
            out[p + "files"] = np.asarray([os.path.basename(f) for f in b["im_file"]])
            out[p + "ori_shape"] = np.asarray(b["ori_shape"])
            out[p + "ratio"] = np.asarray([rp[0] for rp in b["ratio_pad"]], dtype=np.float64)
            out[p + "pad"] = np.asarray([rp[1] for rp in b["ratio_pad"]])
        out[f"r{int(rect)}_nb"] = np.int64(bi + 1)
    # crafted detections per image (jittered labels, some flipped classes, boxes reaching into the padding / beyond the
    # image) through the REAL validator's _prepare_batch / _prepare_pred / _process_batch -> the expected tp matrices
    import functools
    import types

    from ultralytics.engine.validator import BaseValidator
    from ultralytics.models.yolo.detect.val import DetectionValidator as RefVal

    ns = types.SimpleNamespace(device="cpu", iouv=torch.linspace(0.5, 0.95, 10), niou=10)
    ns.match_predictions = functools.partial(BaseValidator.match_predictions, ns)
    from ultralytics.utils.metrics import ConfusionMatrix as RefCM

    cm = RefCM(nc=2, conf=0.001, iou_thres=0.45)         # conf 0.001 -> 0.25 inside, as the validator constructs it
    rng = np.random.default_rng(17)
    cfg = get_cfg(overrides=dict(imgsz=96, task="detect", rect=True, workers=0))
    ds = build_yolo_dataset(cfg, data["val"], 3, data, mode="val", stride=32)
    for bi, b in enumerate(build_dataloader(ds, 3, 0, shuffle=False, rank=-1)):
        B, _, H, W = b["img"].shape
        max_n = 12
        preds = np.zeros((B, max_n, 6), np.float32)
        cnts = np.zeros(B, np.int32)
        for si in range(B):
            m = (b["batch_idx"] == si).numpy()
            bb = b["bboxes"].numpy()[m] * np.array([W, H, W, H], np.float32)
            cl = b["cls"].numpy().reshape(-1)[m]
            rows = []
            for (cx, cy, w, h), c in zip(bb, cl):
                j = rng.normal(0, 1.5, 4)
                rows.append([cx - w / 2 + j[0], cy - h / 2 + j[1], cx + w / 2 + j[2], cy + h / 2 + j[3], rng.uniform(0.3, 0.9),
                             c if rng.uniform() > 0.25 else 1 - c])
            for _ in range(3):                          # boxes that stick out of the content area / the tensor
                x1, y1 = rng.uniform(-10, W - 20), rng.uniform(-10, H - 20)
                rows.append([x1, y1, x1 + rng.uniform(10, 60), y1 + rng.uniform(10, 60), rng.uniform(0.05, 0.5), int(rng.integers(0, 2))])
            rows.sort(key=lambda r: -r[4])              # NMS output order: confidence descending
            preds[si, :len(rows)] = np.asarray(rows, np.float32)
            cnts[si] = len(rows)
            pb = RefVal._prepare_batch(ns, si, b)
            predn = RefVal._prepare_pred(ns, torch.from_numpy(preds[si, :len(rows)]), pb)
            tp = RefVal._process_batch(ns, predn, pb["bbox"], pb["cls"]) if len(pb["cls"]) else torch.zeros((len(rows), 10), dtype=torch.bool)
            out[f"match_b{bi}_tp{si}"] = tp.numpy()
            out[f"match_b{bi}_predn{si}"] = predn.numpy()
            out[f"match_b{bi}_gt{si}"] = torch.cat((pb["cls"].reshape(-1, 1).float(), pb["bbox"].reshape(-1, 4).float()), 1).numpy()
            cm.process_batch(predn, pb["bbox"], pb["cls"])
        out[f"match_b{bi}_preds"] = preds
        out[f"match_b{bi}_cnt"] = cnts
    out["confusion_matrix"] = cm.matrix
    out["confusion_tp"], out["confusion_fp"] = cm.tp_fp()
    for c in (root / "labels").glob("*.cache"):
        c.unlink()                                     # the reference caches the label scan next to the labels
    np.savez_compressed(GOLD / "tiny_dataset_batches.npz", **out)
    print("dataset batches", len(out))


def gen_results():
    """Egress formats of the REAL reference's Results (verbose / summary / to_json / save_txt) for a fixed detection set."""
    import json
    import tempfile

    from ultralytics.engine.results import Results as RefResults

    rng = np.random.default_rng(23)
    n = 7
    xy = rng.uniform(5, 300, (n, 2))
    wh = rng.uniform(8, 120, (n, 2))
    boxes = np.concatenate([xy, xy + wh, rng.uniform(0.26, 0.97, (n, 1)), rng.integers(0, 2, (n, 1))], 1).astype(np.float32)
    boxes = boxes[np.argsort(-boxes[:, 4])]
    names = {0: "wifi", 1: "bluetooth"}
    cases = {}
    for tag, b in (("some", boxes), ("none", boxes[:0])):
        r = RefResults(np.zeros((360, 480, 3), np.uint8), path="x.jpg", names=names, boxes=torch.from_numpy(b))
        with tempfile.TemporaryDirectory() as td:
            r.save_txt(f"{td}/a.txt", save_conf=True)
            r.save_txt(f"{td}/b.txt", save_conf=False)
            ta = open(f"{td}/a.txt").read() if len(b) else ""
            tb = open(f"{td}/b.txt").read() if len(b) else ""
        cases[tag] = {"boxes": b.tolist(), "orig_shape": [360, 480], "verbose": r.verbose(), "summary": r.summary(),
                      "summary_norm": r.summary(normalize=True, decimals=3), "to_json": r.to_json(), "to_csv": r.to_csv(),
                      "txt_conf": ta, "txt": tb,
                      "xywhn": r.boxes.xywhn.tolist(), "xyxyn": r.boxes.xyxyn.tolist()}
    (GOLD / "results_egress.json").write_text(json.dumps(cases, indent=1))
    print("results egress", {k: len(v["boxes"]) for k, v in cases.items()})


from oracle.loss_ref import loss_case  # noqa: E402


LOSS_CASES = [  # name, seed, B, H, W, nc, boxes per image, piled up
    ("mixed", 0, 3, 160, 192, 2, (5, 0, 9), False),
    ("dense80", 1, 2, 128, 128, 80, (12, 7), True),
    ("empty", 2, 2, 96, 96, 2, (0, 0), False),
    ("tiny_boxes", 3, 2, 128, 160, 3, (6, 6), False),
]


def gen_loss():
    """loss_cases.npz: the REAL v8DetectionLoss (utils/loss.py:166-275) on seeded head maps: loss items, total, autograd
    gradients with respect to every head map, and the assigner's targets."""
    from types import SimpleNamespace

    from ultralytics.nn.tasks import DetectionModel as RefModel
    from ultralytics.utils.loss import v8DetectionLoss

    out = {}
    for name, seed, B, H, W, nc, n_gt, dense in LOSS_CASES:
        model = RefModel(f"{REFERENCE_ROOT}/ultralytics/cfg/models/11/yolo11n.yaml", nc=nc, verbose=False)
        model.args = SimpleNamespace(box=7.5, cls=0.5, dfl=1.5)
        crit = v8DetectionLoss(model)
        feats, batch = loss_case(seed, B, H, W, nc, n_gt, dense)
        if name == "tiny_boxes":                      # boxes narrower than a stride-8 cell: fewer than topk anchors inside
            batch["bboxes"][::2, 2:] = torch.tensor([0.03, 0.45])
        feats = [f.requires_grad_(True) for f in feats]
        captured = {}
        orig = crit.assigner.forward

        def spy(*a, **k):
            r = orig(*a, **k)
            captured["r"] = r
            return r

        crit.assigner.forward = spy
        total, items = crit(feats, batch)
        total.backward()
        _, tb, ts, fg, gi = captured["r"]
        out[f"{name}.meta"] = np.asarray([seed, B, H, W, nc, int(dense)] + list(n_gt), dtype=np.int64)
        out[f"{name}.bboxes"] = batch["bboxes"].numpy()
        out[f"{name}.total"] = total.detach().numpy()
        out[f"{name}.items"] = items.numpy()
        for i, f in enumerate(feats):
            out[f"{name}.grad{i}"] = f.grad.numpy()
        out[f"{name}.target_scores"] = ts.numpy().astype(np.float32)
        out[f"{name}.target_boxes"] = tb.numpy().astype(np.float32)      # grid units: the criterion divides the assigner's output by the stride in place (loss.py:266)
        out[f"{name}.fg"] = fg.numpy()
        print(name, "items", items.tolist(), "fg", int(fg.sum()), "scored", int((ts.sum(-1) > 0).sum()))
    np.savez_compressed(GOLD / "loss_cases.npz", **out)


if __name__ == "__main__":
    import_reference()
    GOLD.mkdir(parents=True, exist_ok=True)
    which = {a for a in sys.argv[1:] if "=" not in a} or {"models", "nms", "letterbox", "metrics", "letterbox_u8"}
    if "models" in which:
        gen_models()
    if "nms" in which:
        gen_nms()
    if "letterbox" in which:
        gen_letterbox()
    if "metrics" in which:
        gen_metrics()
    if "letterbox_u8" in which:
        gen_letterbox_u8()
    if "ckpt" in which:
        gen_ckpt()
    if "dataset" in which:
        gen_dataset()
    if "results" in which:
        gen_results()
    if "loss" in which:
        gen_loss()
