"""TEST INFRASTRUCTURE — lets the UNMODIFIED reference train its own detector for a few hundred steps on synthetic
spectrogram images, so that parity can be stated on weights with trained statistics (peaked DFL distributions, calibrated
class scores) instead of random ones:

    python oracle/train_fixture.py --out tests/golden/_trained/spec_s.pt [--device 0] [--epochs 30] [--images 256]

The reference's trainer (`ultralytics/engine/trainer.py`, `models/yolo/detect/train.py`) runs as itself from `oracle/_ref`
(or /root/reference in the build container); nothing of specyolo is involved.  Outputs: the checkpoint the trainer wrote
(`weights/last.pt`, EMA weights in fp16 after `strip_optimizer`, trainer.py:672-687); `held_out()` regenerates held-out
images of the same generator for the parity test (`tests/test_parity_trained.py`).

Synthetic "spectrograms": 640 x 640 uint8, noise floor + 3..9 emissions of two classes (0: narrowband, long in time;
1: wideband burst, short in time) with soft edges — the geometry of the dataset YAMLs the reference ships
(`ultralytics/cfg/datasets/*.yaml` of the spectrogram sets: two classes, pre-rendered images).
"""
from __future__ import annotations

import argparse
import os
import shutil
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def synth_image(rng: np.random.Generator, size: int = 640):
    img = rng.normal(58.0, 9.0, (size, size)).astype(np.float32)
    img += 6.0 * np.sin(np.arange(size, dtype=np.float32) * rng.uniform(0.01, 0.03))[None, :]
    labels = []
    for _ in range(int(rng.integers(3, 10))):
        cls = int(rng.integers(0, 2))
        if cls == 0:
            h, w = int(rng.integers(6, 20)), int(rng.integers(90, 420))
        else:
            h, w = int(rng.integers(50, 220)), int(rng.integers(14, 70))
        h, w = max(2, h * size // 640), max(2, w * size // 640)
        y0, x0 = int(rng.integers(0, size - h)), int(rng.integers(0, size - w))
        amp = float(rng.uniform(60.0, 150.0))
        yy = np.minimum(np.arange(h), np.arange(h)[::-1]).astype(np.float32)
        xx = np.minimum(np.arange(w), np.arange(w)[::-1]).astype(np.float32)
        soft = np.minimum(1.0, (yy[:, None] + 1.0) / 2.0) * np.minimum(1.0, (xx[None, :] + 1.0) / 2.0)
        img[y0:y0 + h, x0:x0 + w] += amp * soft * rng.uniform(0.8, 1.0, (h, w)).astype(np.float32)
        labels.append((cls, (x0 + w / 2) / size, (y0 + h / 2) / size, w / size, h / size))
    img = np.clip(img, 0, 255).astype(np.uint8)
    return np.repeat(img[:, :, None], 3, axis=2), labels


def held_out(n: int = 64, size: int = 640, seed: int = 12345):
    """Held-out images of the same generator (never written to disk: the parity test regenerates them):
    uint8 [n, 3, size, size] and the ground-truth rows [n, nmax, 5] = (cls, x, y, w, h) normalised, -1 padded."""
    rng = np.random.default_rng(seed)
    imgs, boxes = [], []
    for _ in range(n):
        img, labels = synth_image(rng, size)
        imgs.append(np.ascontiguousarray(img.transpose(2, 0, 1)))
        boxes.append(np.asarray(labels, np.float32).reshape(-1, 5))
    nmax = max(len(b) for b in boxes)
    gt = np.full((len(boxes), nmax, 5), -1.0, np.float32)
    for i, b in enumerate(boxes):
        gt[i, :len(b)] = b
    return np.stack(imgs), gt


def write_dataset(root: str, n_train: int, n_val: int, seed: int = 0) -> str:
    import cv2
    rng = np.random.default_rng(seed)
    for split, n in (("train", n_train), ("val", n_val)):
        os.makedirs(os.path.join(root, "images", split), exist_ok=True)
        os.makedirs(os.path.join(root, "labels", split), exist_ok=True)
        for i in range(n):
            img, labels = synth_image(rng)
            cv2.imwrite(os.path.join(root, "images", split, f"{i:05d}.png"), img)
            with open(os.path.join(root, "labels", split, f"{i:05d}.txt"), "w") as f:
                for c, x, y, w, h in labels:
                    f.write(f"{c} {x:.6f} {y:.6f} {w:.6f} {h:.6f}\n")
    yaml_path = os.path.join(root, "data.yaml")
    with open(yaml_path, "w") as f:
        f.write(f"path: {root}\ntrain: images/train\nval: images/val\nnames:\n  0: narrowband\n  1: burst\n")
    return yaml_path


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--cfg", default="yolo11s_fusion_sand3_new.yaml")
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--epochs", type=int, default=30)
    ap.add_argument("--images", type=int, default=256)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--imgsz", type=int, default=640)
    ap.add_argument("--shim", action="store_true",
                    help="bind specyolo's CUDA criterion + EMA update into the reference trainer (specyolo.ultralytics_shim)")
    a = ap.parse_args()

    import ref_loader
    ref_loader.import_reference()
    from ultralytics import YOLO
    import ultralytics.data.utils as du
    import ultralytics.utils.checks as ck
    # host-side conveniences that need the network / matplotlib (absent here): the label font and the AMP self-test
    ck.check_font = du.check_font = lambda *a, **k: None

    if a.shim:
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "spectrogram-yolov11_b200"))
        from specyolo import ultralytics_shim
        ultralytics_shim.install()

    work = tempfile.mkdtemp(prefix="specyolo_trainfix_")
    data = write_dataset(os.path.join(work, "data"), a.images, 16)
    model = YOLO(a.cfg)
    import time
    stamps = []
    model.add_callback("on_train_batch_end", lambda trainer: stamps.append(time.perf_counter()))
    device = a.device if a.device == "cpu" else int(a.device)
    model.train(data=data, epochs=a.epochs, imgsz=a.imgsz, batch=a.batch, device=device, workers=0, amp=False,
                plots=False, val=False, project=os.path.join(work, "runs"), name="fix", exist_ok=True, pretrained=False,
                optimizer="SGD", lr0=0.01, warmup_epochs=1.0, close_mosaic=0, mosaic=0.0, fliplr=0.0, hsv_h=0.0,
                hsv_s=0.0, hsv_v=0.0, translate=0.0, scale=0.0, erasing=0.0, seed=0, deterministic=False, verbose=False)
    last = os.path.join(work, "runs", "fix", "weights", "last.pt")
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    shutil.copyfile(last, a.out)

    print(f"checkpoint {a.out} ({os.path.getsize(a.out) / 1e6:.1f} MB)")
    if len(stamps) > 12:
        dt = np.diff(np.asarray(stamps))[10:]          # skip the first iterations (cuDNN autotune, allocator warm-up)
        print(f"train step ({'shim: CUDA criterion + EMA' if a.shim else 'stock reference'}): median {np.median(dt) * 1e3:.1f} ms, "
              f"mean {dt.mean() * 1e3:.1f} ms over {len(dt)} iterations, batch {a.batch}, imgsz {a.imgsz}")
    shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
