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
]


def gen_models():
    from ultralytics.nn.tasks import DetectionModel as RefModel

    import specyolo
    from specyolo.nn.init import synth_images, synth_state_dict

    for name, ref_yaml, cfg, nc, H, W, seed in MODEL_CASES:
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


if __name__ == "__main__":
    import_reference()
    GOLD.mkdir(parents=True, exist_ok=True)
    gen_models()
    gen_nms()
    gen_letterbox()
