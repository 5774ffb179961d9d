"""Parity at the BASELINE configurations' own sizes, at DETECTION level ("next" item 1).

* C2: the exact input bench.py times — synth_images(64, 640, seed=0, uint8), seed-0 weights — through
  `YOLO.predict()` (uint8 -> fused stem kernel -> ... -> fused decode -> NMS kernel, CUDA-graph replay) against the
  CPU oracle (fp32 forward + numpy NMS restatement) and, where oracle/_ref travelled, against the REAL reference's
  `YOLO(...).predict()` on the same images.
* C3: 32 bursts x 2^20 complex64 samples through `YOLO.predict_iq()` against float64 STFT spec -> fp32 forward -> NMS.
* The GPU NMS kernel on every case of tests/golden/nms_cases.npz (outputs of the real reference function).
* torchvision.ops.nms on CUDA tensors as a second bit-exactness oracle, at-threshold pairs included.

Tolerance (stated here, measured numbers written to profiles/parity_*.json by the tests).  The reference's own
half-precision precedent is identical count + atol 0.5 on [x1,y1,x2,y2,conf,cls] for a TRAINED fp16 model
(utils/checks.py:691-699).  With the synthetic (random, untrained) weights of the benchmark the DFL distributions are
broad, the box is a 16-bin expectation times the stride (up to 32 px per bin), and bf16 storage between ~60 chained convs
puts ~1-3 % relative error on the head logits: an IDEAL bf16 pipeline (the fp32 oracle with weights and every conv
output rounded to bf16, fp32 accumulation) already sits 4.6 px mean / 40 px worst from fp32 on the candidates of this
input.  The bound asserted is therefore relative to that measured floor — the CUDA path must be no further from fp32 than
the ideal bf16 pipeline (x 1.25 sampling slack) — plus absolute caps: mean <= 0.2 DFL bins, max <= 1.5 bins, scores <= 0.08;
at detection level >= 98 % of the oracle's detections with conf > 0.30 are found (same class, IoU >= 0.5), >= 75 % of them as
the very same anchor.  NMS itself is exact (tests below and tests/test_gpu_kernels.py).
"""
import ast
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import yaml

from _parity import clip_boxes_np, compare_detections, record

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
CFG = ROOT / "spectrogram-yolov11_b200" / "specyolo" / "cfg"
GOLD = ROOT / "tests" / "golden"
CONF, IOU, MAX_DET = 0.25, 0.7, 300


def _yolo(seed=0, cls_bias=None):
    import specyolo
    from specyolo.nn.init import synth_state_dict

    yolo = specyolo.YOLO("yolo11s_fusion_sand3_new.yaml", nc=2)
    sd = synth_state_dict(yolo.model, seed=seed, cls_bias=cls_bias)
    yolo.load_state_dict(sd)
    yolo.to("cuda")
    return yolo, sd


def _graph():
    from oracle import yolo_ref

    return yolo_ref.parse_graph(yaml.safe_load((CFG / "yolo11_fusion_sand3_new.yaml").read_text()), "s", 2)


def _oracle_dense(sd, x_float, chunk=8, bf16_storage=False):
    """fp32 oracle forward (restatement of the reference, pinned by tests/golden/model_*.npz): dense y [B, 6, A].
    bf16_storage=True: the same forward with weights and every Conv output rounded to bf16 (fp32 accumulation) — the
    numerically ideal bf16 pipeline, i.e. the error floor any bf16 implementation has against fp32."""
    from oracle import yolo_ref

    R = yolo_ref.Ref(sd, bf16_storage=bf16_storage)
    g = _graph()
    out = []
    with torch.no_grad():
        for i in range(0, x_float.shape[0], chunk):
            out.append(yolo_ref.forward(g, R, x_float[i:i + chunk])[0])
    return torch.cat(out).numpy()


def _nms_with_anchors(y: np.ndarray, hw):
    """numpy NMS restatement on a dense prediction -> (clipped detections per image, anchor index of every detection)."""
    from oracle import nms_ref

    dets, idx = nms_ref.non_max_suppression(y, CONF, IOU, max_det=MAX_DET, return_indices=True)
    anchors = [np.nonzero(y[b, 4:].max(0) > np.float32(CONF))[0][idx[b]] for b in range(y.shape[0])]
    return [clip_boxes_np(d, *hw) for d in dets], anchors


def _anchor_level(y_ref: np.ndarray, y: np.ndarray, strides=(8, 16, 32), size=640) -> dict:
    """Errors of a dense prediction against the fp32 oracle on the oracle's candidates (max score > conf): box error in
    pixels and in DFL bins (pixels / stride: the decode multiplies a 16-bin expectation by the stride), score error."""
    cand = y_ref[:, 4:].max(1) > CONF
    eb = np.abs(y[:, :4] - y_ref[:, :4]).max(1)
    es = np.abs(y[:, 4:] - y_ref[:, 4:]).max(1)
    st = np.concatenate([np.full((size // s) ** 2, float(s)) for s in strides])[None]
    bins = eb / st
    q = lambda a, p: float(np.quantile(a, p)) if a.size else 0.0   # noqa: E731
    out = {"candidates": int(cand.sum()), "box_px_mean": float(eb[cand].mean()), "box_px_p50": q(eb[cand], 0.5),
           "box_px_p99": q(eb[cand], 0.99), "box_px_max": float(eb[cand].max()),
           "box_bins_mean": float(bins[cand].mean()), "box_bins_p99": q(bins[cand], 0.99), "box_bins_max": float(bins[cand].max()),
           "score_mean": float(es[cand].mean()), "score_p99": q(es[cand], 0.99), "score_max": float(es[cand].max())}
    o = 0
    for s in strides:
        n = (size // s) ** 2
        c = cand[:, o:o + n]
        if c.any():
            e = eb[:, o:o + n][c]
            out[f"stride{s}"] = {"candidates": int(c.sum()), "box_px_mean": float(e.mean()), "box_px_p99": q(e, 0.99),
                                 "box_px_max": float(e.max())}
        o += n
    return out


def _same_anchor(ref_dets, ref_anchors, got_dets, got_anchors, floor=0.30) -> dict:
    """Oracle detections (conf > floor) whose ANCHOR the CUDA path's NMS kept too, and the box / score difference on
    exactly those pairs (no IoU matching involved: the same anchor's prediction through both pipelines)."""
    n = hit = 0
    db, ds = [], []
    for r, ra, g, ga in zip(ref_dets, ref_anchors, got_dets, got_anchors):
        pos = {int(a): k for k, a in enumerate(ga)}
        for k, a in enumerate(ra):
            if r[k, 4] <= floor:
                continue
            n += 1
            j = pos.get(int(a))
            if j is not None and g[j, 5] == r[k, 5]:
                hit += 1
                db.append(np.abs(g[j, :4] - r[k, :4]).max())
                ds.append(abs(g[j, 4] - r[k, 4]))
    db, ds = np.asarray(db), np.asarray(ds)
    return {"ref_detections_above_floor": n, "same_anchor_kept": hit, "same_anchor_rate": hit / max(n, 1),
            "dbox_px_mean": float(db.mean()) if len(db) else 0.0, "dbox_px_p99": float(np.quantile(db, 0.99)) if len(db) else 0.0,
            "dbox_px_max": float(db.max()) if len(db) else 0.0, "frac_dbox_le_0p5": float((db <= 0.5).mean()) if len(db) else 1.0,
            "dscore_max": float(ds.max()) if len(ds) else 0.0}


def _cuda_dense_and_detections(yolo, x_dev):
    """Dense prediction of the CUDA path + its NMS kernel with kept anchors, and the check that the API the bench times
    (fused decode + NMS inside the captured graph) returns exactly these detections."""
    from specyolo.utils.ops import non_max_suppression

    y, _ = yolo.model(x_dev)
    dets, idx = non_max_suppression(y, CONF, IOU, max_det=MAX_DET, return_idxs=True)
    yh = y.cpu().numpy()
    anchors = [np.nonzero(yh[b, 4:].max(0) > np.float32(CONF))[0][idx[b].cpu().numpy()] for b in range(yh.shape[0])]
    H, W = x_dev.shape[2:]
    return yh, [clip_boxes_np(d.cpu().numpy(), H, W) for d in dets], anchors


def _check(stats):
    a, f, d, s = stats["anchor_level"], stats["bf16_floor_anchor_level"], stats["detection_level"], stats["same_anchor"]
    # (1) per-anchor: the CUDA path is no further from fp32 than the ideal bf16-storage pipeline is (25 % slack for
    #     sampling: the floor is measured on a subset), and bounded absolutely in DFL bins / score
    assert a["box_px_mean"] <= 1.25 * f["box_px_mean"] + 0.05, (a, f)
    assert a["box_px_p99"] <= 1.25 * f["box_px_p99"] + 0.25, (a, f)
    assert a["score_p99"] <= 1.25 * f["score_p99"] + 0.005, (a, f)
    assert a["box_bins_mean"] <= 0.2 and a["box_bins_max"] <= 1.5 and a["score_max"] <= 0.08, a
    # (2) detections: >= 98 % of the oracle's confident detections are found (same class, IoU >= 0.5) ...
    assert d["ref_detections_above_floor"] > d["images"], "too few confident detections to call this a test"
    assert d["matched_rate"] >= 0.98, d
    # (3) ... most of them as the very same anchor; on those the box moves by bf16 noise only
    assert s["same_anchor_rate"] >= 0.75, s
    assert s["dscore_max"] <= 0.08, s


def test_c2_bench_input_detections_vs_oracle(lib):
    """BASELINE configs[1] exactly as bench.py runs it (B=64, 640^2, uint8, seed 0) against the fp32 oracle, per anchor
    and per detection; the bf16 error floor is measured beside it (oracle with bf16 storage, first 16 images)."""
    from specyolo.nn.init import synth_images

    yolo, sd = _yolo(0)
    x_u8 = synth_images(64, 640, seed=0, dtype=torch.uint8)
    res = yolo.predict(x_u8.cuda(), conf=CONF, iou=IOU, max_det=MAX_DET)         # graph-captured fused path
    y_cuda, got, got_anchors = _cuda_dense_and_detections(yolo, x_u8.cuda())
    for r, g in zip(res, got):                                                   # fused decode+NMS == dense + NMS kernel
        assert np.array_equal(r.boxes.data.cpu().numpy(), g)
    xf = x_u8.float() / 255                                                      # predictor.py:133-135
    y_ref = _oracle_dense(sd, xf)
    ref, ref_anchors = _nms_with_anchors(y_ref, (640, 640))
    y_floor = _oracle_dense(sd, xf[:16], bf16_storage=True)
    stats = {"config": "C2: spectrogram-yolov11-s, B=64, 640^2, uint8, seed 0 (the bench input)",
             "anchor_level": _anchor_level(y_ref, y_cuda),
             "bf16_floor_anchor_level": _anchor_level(y_ref[:16], y_floor),
             "detection_level": compare_detections(ref, got),
             "same_anchor": _same_anchor(ref, ref_anchors, got, got_anchors)}
    record("c2_b64_640_vs_oracle", stats)
    _check(stats)
    # the same batch through the streaming API (what `e2e` times): identical to the one-shot call
    streamed = list(yolo.predict([x_u8.pin_memory()], stream=True, conf=CONF, iou=IOU, max_det=MAX_DET))[0]
    for a, b in zip(res, streamed):
        assert torch.equal(a.boxes.data.cpu(), b.boxes.data.cpu())


def test_c2_bench_input_detections_vs_real_reference(lib):
    """Same input against the REAL reference's public API on the CPU (oracle/_ref, installed by oracle/Makefile)."""
    sys.path.insert(0, str(ROOT))
    from oracle import ref_loader

    if not ref_loader.reference_available():
        pytest.skip("oracle/_ref not present on this box")
    ultralytics = ref_loader.import_reference()
    from specyolo.nn.init import synth_images

    yolo, sd = _yolo(0)
    n = 16                                                                        # bounded: the CPU reference runs ~10-30 images/s
    x_u8 = synth_images(64, 640, seed=0, dtype=torch.uint8)[:n]
    res = yolo.predict(x_u8.cuda(), conf=CONF, iou=IOU, max_det=MAX_DET)
    got = [r.boxes.data.cpu().numpy() for r in res]
    cfg = Path(ultralytics.__file__).parent / "cfg" / "models" / "11" / "yolo11s_fusion_sand3_new.yaml"
    from ultralytics.nn.tasks import DetectionModel as RefModel

    ref_model = RefModel(str(cfg), nc=2, verbose=False)
    ref_model.load_state_dict(sd, strict=True)
    ref_yolo = ultralytics.YOLO(str(cfg), task="detect")
    ref_yolo.model = ref_model.eval()
    ref_res = ref_yolo.predict(x_u8.float() / 255, device="cpu", conf=CONF, iou=IOU, max_det=MAX_DET, verbose=False)
    ref = [r.boxes.data.cpu().numpy() for r in ref_res]
    # the oracle port IS the reference to 1e-2 px / 1e-5 (so the port-based numbers above carry over) ...
    port, _ = _nms_with_anchors(_oracle_dense(sd, x_u8.float() / 255), (640, 640))
    pin = compare_detections(ref, port)
    assert pin["matched_rate"] == 1.0 and pin["max_dbox_px"] < 0.02 and pin["max_dscore"] < 1e-4 and \
        pin["images_with_count_diff"] == 0, pin
    # ... and the CUDA path finds the real reference's detections
    stats = {"oracle_port_vs_real_reference": pin, "cuda_vs_real_reference": compare_detections(ref, got)}
    record("c2_b16_640_vs_real_reference", stats)
    assert stats["cuda_vs_real_reference"]["matched_rate"] >= 0.98, stats


def test_c3_iq_bursts_detections_vs_oracle(lib):
    """BASELINE configs[2] per-GPU share: 32 bursts x 2^20 samples -> STFT kernel -> detector -> NMS (one captured graph),
    vs the float64 STFT spec -> fp32 oracle forward -> numpy NMS.  STFT parity itself is UNPINNED: the reference has no
    IQ code; the spectrogram tensor is compared with the float64 spec at the bf16 output resolution."""
    from oracle import stft_ref
    from specyolo import ops
    from specyolo.nn.init import EMISSION_DB_RANGE, IQ_CLS_BIAS, synth_iq_emissions

    yolo, sd = _yolo(0, cls_bias=IQ_CLS_BIAS)                                     # what bench.py's c3 leg runs
    iq = synth_iq_emissions(32, 1 << 20, seed=3)
    lo, hi = EMISSION_DB_RANGE
    res = yolo.predict_iq(iq, db_min=lo, db_max=hi, conf=CONF, iou=IOU, max_det=MAX_DET, imgsz=640)
    img = ops.iq_to_letterbox(iq.cuda(), db_min=lo, db_max=hi, out_hw=(640, 640))  # the kernel's own image (bf16)
    y_cuda, got, got_anchors = _cuda_dense_and_detections(yolo, img)
    for r, g in zip(res, got):                                                   # graph(STFT + detector + NMS) == staged
        assert np.array_equal(r.boxes.data.cpu().numpy(), g)
    spec = stft_ref.iq_to_letterbox(iq.numpy(), db_min=lo, db_max=hi, out_hw=(640, 640))
    err_img = float(np.abs(img.float().cpu().numpy() - spec).max())
    assert err_img <= 8e-3, err_img                                               # bf16 output: half an ulp near 1.0 is 3.9e-3
    xf = torch.from_numpy(spec).float()
    y_ref = _oracle_dense(sd, xf)
    ref, ref_anchors = _nms_with_anchors(y_ref, (640, 640))
    y_floor = _oracle_dense(sd, xf[:8], bf16_storage=True)
    stats = {"config": "C3: 32 bursts x 2^20 complex64 -> STFT 1024/256 -> letterbox 640^2 -> spectrogram-yolov11-s -> NMS",
             "spectrogram_max_abs_err": err_img,
             "anchor_level": _anchor_level(y_ref, y_cuda),
             "bf16_floor_anchor_level": _anchor_level(y_ref[:8], y_floor),
             "detection_level": compare_detections(ref, got),
             "same_anchor": _same_anchor(ref, ref_anchors, got, got_anchors)}
    record("c3_b32_iq_vs_oracle", stats)
    _check(stats)


def test_gpu_nms_on_reference_fixtures(lib):
    """specyolo.utils.ops.non_max_suppression (nms_kernel) on every case of nms_cases.npz — the stored outputs are those
    of the REAL reference's non_max_suppression (oracle/gen_golden.py nms): bit-exact rows, same order."""
    from specyolo.utils.ops import non_max_suppression

    z = np.load(GOLD / "nms_cases.npz")
    names = sorted({k[:-5] for k in z.files if k.endswith("_pred")})
    assert len(names) >= 7
    for name in names:
        pred = z[f"{name}_pred"]
        kw = ast.literal_eval(str(z[f"{name}_kw"]))
        out = non_max_suppression(torch.from_numpy(pred.copy()).cuda(), **kw)
        for b, o in enumerate(out):
            exp = z[f"{name}_out{b}"]
            assert tuple(o.shape) == exp.shape, (name, b, tuple(o.shape), exp.shape)
            assert np.array_equal(o.cpu().numpy(), exp), (name, b)


def test_nms_vs_torchvision_cuda(lib):
    """torchvision.ops.nms on CUDA tensors (what the reference runs at utils/ops.py:312 on a GPU) as the oracle for the
    keep indices: random candidates with score ties, plus pairs sitting EXACTLY at the IoU threshold.  SURVEY 8 a10-6:
    torchvision's CPU kernel compares the fp32 quotient with a double threshold, the CUDA kernel takes a float threshold;
    the test records which way each at-threshold pair falls in both and requires agreement with the CUDA kernel wherever
    CPU and CUDA agree, and with the CPU kernel (the readable spec) on every case."""
    import torchvision

    from specyolo import ops

    g = torch.Generator().manual_seed(11)
    report = {}
    for case, (n, thr) in {"random_0.7": (1500, 0.7), "random_0.45": (900, 0.45), "dense_0.5": (1024, 0.5)}.items():
        xy = torch.rand((n, 2), generator=g) * 500 + 20
        wh = torch.rand((n, 2), generator=g) * 150 + 8
        sc = torch.rand((n,), generator=g) * 0.7 + 0.3
        sc[1::7] = sc[0:-1:7][: sc[1::7].numel()]                               # ties
        pred = torch.zeros((1, 5, n))
        pred[0, :2], pred[0, 2:4], pred[0, 4] = xy.t(), wh.t(), sc
        # single class, agnostic: the kernel sees exactly these boxes (xywh -> xyxy is the reference's own arithmetic,
        # so rebuild the oracle's boxes the same way: ops.py:257-260)
        b2 = torch.cat((pred[0, :2].t() - pred[0, 2:4].t() / 2, pred[0, :2].t() + pred[0, 2:4].t() / 2), 1)
        keep_cuda = torchvision.ops.nms(b2.cuda(), sc.cuda(), thr).cpu()
        keep_cpu = torchvision.ops.nms(b2, sc, thr)
        out, cnt, keep, ncand = ops.nms(prediction=pred.cuda().contiguous(), B=1, nc=1, A=n, conf_thres=0.25,
                                        iou_thres=thr, agnostic=True, max_det=300)
        k = keep[0, : int(cnt[0])].cpu().long()
        same = torch.equal(keep_cuda, keep_cpu)
        report[case] = {"n": n, "kept_cuda": int(keep_cuda.numel()), "cpu_equals_cuda": bool(same)}
        assert torch.equal(k, keep_cpu[:300]), case
        if same:
            assert torch.equal(k, keep_cuda[:300]), case
    # at-threshold pairs: box B overlaps box A with IoU exactly 1/3, 1/2, 3/5 (representable geometry)
    at = {}
    for name, (shift, thr) in {"iou_1_3": (50.0, 1.0 / 3.0), "iou_1_2": (100.0 / 3.0, 0.5), "iou_3_5": (25.0, 0.6)}.items():
        # two 100 x 100 boxes shifted by s along x: inter = (100 - s) * 100, union = (100 + s) * 100
        a = torch.tensor([[100.0, 100.0, 200.0, 200.0], [100.0 + shift, 100.0, 200.0 + shift, 200.0]])
        s = torch.tensor([0.9, 0.8])
        kc = torchvision.ops.nms(a, s, thr).tolist()
        kg = torchvision.ops.nms(a.cuda(), s.cuda(), thr).cpu().tolist()
        pred = torch.zeros((1, 5, 2))
        pred[0, 0] = (a[:, 0] + a[:, 2]) / 2; pred[0, 1] = (a[:, 1] + a[:, 3]) / 2
        pred[0, 2] = a[:, 2] - a[:, 0]; pred[0, 3] = a[:, 3] - a[:, 1]; pred[0, 4] = s
        out, cnt, keep, _ = ops.nms(prediction=pred.cuda().contiguous(), B=1, nc=1, A=2, conf_thres=0.25, iou_thres=thr,
                                    agnostic=True, max_det=300)
        km = keep[0, : int(cnt[0])].cpu().tolist()
        # "cuda" semantics (specyolo.utils.ops.TORCHVISION_NMS_SEMANTICS): the threshold rounded to float first
        thr_f = float(torch.tensor(thr, dtype=torch.float32).item())
        _, cnt2, keep2, _ = ops.nms(prediction=pred.cuda().contiguous(), B=1, nc=1, A=2, conf_thres=0.25, iou_thres=thr_f,
                                    agnostic=True, max_det=300)
        km2 = keep2[0, : int(cnt2[0])].cpu().tolist()
        at[name] = {"torchvision_cpu": kc, "torchvision_cuda": kg, "specyolo": km, "specyolo_cuda_semantics": km2}
        assert km == kc, (name, at[name])                    # default: the CPU kernel, the spec the golden vectors pin
        assert km2 == kg, (name, at[name])                   # "cuda": what the reference returns on a GPU
    report["at_threshold"] = at
    record("nms_vs_torchvision_cuda", report)
