"""Validation-mode dataset ingest: dataset YAML -> image + label files -> the reference's collated batches.

Mirrors, for `mode="val"` (no augmentation), what `DetectionValidator.get_dataloader` builds in the reference:
`check_det_dataset` (data/utils.py:300-394, local paths only), `YOLODataset.get_img_files / get_labels`
(data/base.py:104-135, data/dataset.py:73-172, `verify_image_label` data/utils.py:97-165), rectangular batch shapes
(`BaseDataset.set_rectangle`, data/base.py:261-284: images sorted by aspect ratio, one shape per batch, pad 0.5),
`load_image(rect_mode=True)` (data/base.py:151-187: long side resized to imgsz, up or down), `LetterBox(scaleup=False,
center=True)` with its label update (data/augment.py:1535-1615), `Format(xywh, normalize)` (data/augment.py:1965-2095) and
`YOLODataset.collate_fn` (data/dataset.py:232-248).  Host work = file lists, label text, a few floats per box; pixels are
decoded into device memory (specyolo.data.loaders) and resized + padded + BGR->RGB + HWC->CHW by ONE launch of
`specyolo_letterbox_u8` per image (bit-exact with cv2.resize INTER_LINEAR + copyMakeBorder).
"""
from __future__ import annotations

import glob
import math
import os
from pathlib import Path
from typing import Dict, List, Optional

import numpy as np
import torch

from .. import ops
from .loaders import IMG_FORMATS, imread_device


def check_det_dataset(dataset) -> dict:
    """Dataset YAML -> dict with absolute `train` / `val` / `test` paths, `names` (dict) and `nc` (data/utils.py:300-394;
    nothing is downloaded)."""
    import yaml

    file = Path(str(dataset))
    if not file.is_file():
        raise FileNotFoundError(f"dataset '{dataset}' not found")
    data = yaml.safe_load(file.read_text())
    for k in ("train", "val"):
        if k not in data:
            raise SyntaxError(f"{dataset} '{k}:' key missing. 'train' and 'val' are required in all data YAMLs.")
    if "names" not in data and "nc" not in data:
        raise SyntaxError(f"{dataset} key missing. either 'names' or 'nc' are required in all data YAMLs.")
    if "names" in data and "nc" in data and len(data["names"]) != data["nc"]:
        raise SyntaxError(f"{dataset} 'names' length {len(data['names'])} and 'nc: {data['nc']}' must match.")
    if "names" not in data:
        data["names"] = [f"class_{i}" for i in range(data["nc"])]
    else:
        data["nc"] = len(data["names"])
    if isinstance(data["names"], (list, tuple)):
        data["names"] = dict(enumerate(data["names"]))
    path = Path(data.get("path") or file.parent)
    if not path.is_absolute():
        path = (file.parent / path).resolve()
    data["path"] = path
    for k in ("train", "val", "test"):
        if data.get(k):
            if isinstance(data[k], str):
                x = (path / data[k]).resolve()
                data[k] = str(x)
            else:
                data[k] = [str((path / x).resolve()) for x in data[k]]
    data["yaml_file"] = str(file)
    return data


def img2label_paths(img_paths: List[str]) -> List[str]:
    sa, sb = f"{os.sep}images{os.sep}", f"{os.sep}labels{os.sep}"
    return [sb.join(x.rsplit(sa, 1)).rsplit(".", 1)[0] + ".txt" for x in img_paths]


def _image_hw(path: str):
    from PIL import Image

    with Image.open(path) as im:
        w, h = im.size
        fmt = (im.format or "").lower()
        if im.format == "JPEG":
            try:
                if exif := im.getexif():
                    if exif.get(274, None) in {6, 8}:
                        w, h = h, w
            except Exception:
                pass
    return (h, w), fmt


def verify_image_label(im_file: str, lb_file: str, num_cls: int):
    """(labels [n,5] float32 normalised `cls cx cy w h`, (h, w)) or None for a corrupt pair (data/utils.py:97-165)."""
    try:
        shape, fmt = _image_hw(im_file)
        assert (shape[0] > 9) and (shape[1] > 9), f"image size {shape} <10 pixels"
        assert fmt in IMG_FORMATS | {"jpeg"}, f"invalid image format {fmt}"
        if os.path.isfile(lb_file):
            rows = [x.split() for x in Path(lb_file).read_text().strip().splitlines() if len(x)]
            lb = np.array(rows, dtype=np.float32)
            if nl := len(lb):
                assert lb.shape[1] == 5, f"labels require 5 columns, {lb.shape[1]} columns detected"
                assert lb[:, 1:].max() <= 1, "non-normalized or out of bounds coordinates"
                assert lb.min() >= 0, "negative label values"
                assert lb[:, 0].max() < num_cls, f"Label class {int(lb[:, 0].max())} exceeds dataset class count {num_cls}"
                _, i = np.unique(lb, axis=0, return_index=True)
                if len(i) < nl:
                    lb = lb[i]                       # duplicates removed (rows come back in np.unique order, as upstream)
            else:
                lb = np.zeros((0, 5), dtype=np.float32)
        else:
            lb = np.zeros((0, 5), dtype=np.float32)
        return lb[:, :5].reshape(-1, 5), shape
    except Exception:
        return None


class YOLODataset:
    """Validation batches of a detection dataset: iterate to get the reference's collated dicts with `img` on the device."""

    def __init__(self, img_path, imgsz: int = 640, batch_size: int = 16, rect: bool = False, stride: int = 32,
                 pad: float = 0.5, data: Optional[dict] = None, single_cls: bool = False, classes=None, device="cuda"):
        self.imgsz, self.batch_size, self.rect, self.stride, self.pad = imgsz, batch_size, rect, stride, pad
        self.data = data or {}
        self.device = device
        self.im_files = self.get_img_files(img_path)
        nc = len(self.data.get("names", {})) or int(self.data.get("nc", 1 << 30))
        labels, files = [], []
        for f, lf in zip(self.im_files, img2label_paths(self.im_files)):
            r = verify_image_label(f, lf, nc)
            if r is None:
                continue                                  # corrupt image / label: skipped with a warning upstream
            lb, shape = r
            if classes is not None:
                lb = lb[np.isin(lb[:, 0], np.asarray(classes, dtype=np.float32))]
            if single_cls:
                lb[:, 0] = 0
            labels.append({"im_file": f, "shape": shape, "cls": lb[:, 0:1], "bboxes": lb[:, 1:]})
            files.append(f)
        self.labels, self.im_files = labels, files
        self.ni = len(self.labels)
        if self.ni == 0:
            raise FileNotFoundError(f"No valid images found in {img_path}")
        self.batch = np.floor(np.arange(self.ni) / self.batch_size).astype(int)
        self.batch_shapes = None
        if self.rect:
            self.set_rectangle()

    @staticmethod
    def get_img_files(img_path) -> List[str]:
        f: List[str] = []
        for p in img_path if isinstance(img_path, list) else [img_path]:
            p = Path(p)
            if p.is_dir():
                f += glob.glob(str(p / "**" / "*.*"), recursive=True)
            elif p.is_file():                           # a text file of paths, relative entries anchored at its parent
                parent = str(p.parent) + os.sep
                f += [x.replace("./", parent) if x.startswith("./") else x for x in p.read_text().strip().splitlines()]
            else:
                raise FileNotFoundError(f"{p} does not exist")
        im_files = sorted(x.replace("/", os.sep) for x in f if x.split(".")[-1].lower() in IMG_FORMATS | {"heic"})
        if not im_files:
            raise FileNotFoundError(f"No images found in {img_path}")
        return im_files

    def set_rectangle(self):
        bi = self.batch
        nb = bi[-1] + 1
        s = np.array([x["shape"] for x in self.labels], dtype=np.float64)       # hw
        ar = s[:, 0] / s[:, 1]
        irect = ar.argsort()
        self.im_files = [self.im_files[i] for i in irect]
        self.labels = [self.labels[i] for i in irect]
        ar = ar[irect]
        shapes = [[1, 1]] * nb
        for i in range(nb):
            ari = ar[bi == i]
            mini, maxi = ari.min(), ari.max()
            if maxi < 1:
                shapes[i] = [maxi, 1]
            elif mini > 1:
                shapes[i] = [1, 1 / mini]
        self.batch_shapes = np.ceil(np.array(shapes) * self.imgsz / self.stride + self.pad).astype(int) * self.stride

    def __len__(self):
        return int(self.batch[-1]) + 1

    def geometry(self, index: int) -> dict:
        """Host-side geometry of sample `index`: resized size, final shape, pads, ratio_pad (no pixels touched)."""
        lab = self.labels[index]
        h0, w0 = lab["shape"]
        r = self.imgsz / max(h0, w0)                      # load_image(rect_mode=True)
        if r != 1:
            w, h = min(math.ceil(w0 * r), self.imgsz), min(math.ceil(h0 * r), self.imgsz)
        else:
            w, h = w0, h0
        new_shape = tuple(int(v) for v in self.batch_shapes[self.batch[index]]) if self.rect else (self.imgsz, self.imgsz)
        r2 = min(new_shape[0] / h, new_shape[1] / w, 1.0)                        # LetterBox(scaleup=False)
        new_unpad = int(round(w * r2)), int(round(h * r2))
        dw, dh = (new_shape[1] - new_unpad[0]) / 2, (new_shape[0] - new_unpad[1]) / 2
        top, left = int(round(dh - 0.1)), int(round(dw - 0.1))
        return dict(ori_shape=(h0, w0), resized=(h, w), new_shape=new_shape, new_unpad=new_unpad, r2=r2, left=left, top=top,
                    ratio_pad=((h / h0, w / w0), (left, top)))

    def sample_labels(self, index: int, g: dict):
        """Normalised xywh of the sample's boxes in the letterboxed frame, float32 arithmetic in the reference's order."""
        lab = self.labels[index]
        b = lab["bboxes"].astype(np.float32).copy()
        h, w = g["resized"]
        y = np.empty_like(b)                              # xywh -> xyxy (utils/ops.py xywh2xyxy)
        dw_, dh_ = b[:, 2] / 2, b[:, 3] / 2
        y[:, 0], y[:, 1], y[:, 2], y[:, 3] = b[:, 0] - dw_, b[:, 1] - dh_, b[:, 0] + dw_, b[:, 1] + dh_
        y[:, [0, 2]] *= w                                 # denormalize
        y[:, [1, 3]] *= h
        y[:, [0, 2]] *= g["r2"]                           # scale(ratio)
        y[:, [1, 3]] *= g["r2"]
        y[:, [0, 2]] += g["left"]                         # add_padding
        y[:, [1, 3]] += g["top"]
        o = np.empty_like(y)                              # xyxy -> xywh
        o[:, 0], o[:, 1] = (y[:, 0] + y[:, 2]) / 2, (y[:, 1] + y[:, 3]) / 2
        o[:, 2], o[:, 3] = y[:, 2] - y[:, 0], y[:, 3] - y[:, 1]
        H, W = g["new_shape"]
        o[:, [0, 2]] /= W                                 # normalize by the final image size
        o[:, [1, 3]] /= H
        return lab["cls"].astype(np.float32), o

    def load_sample_image(self, index: int, g: dict, out: torch.Tensor):
        """Decode + resize + pad + BGR->RGB + HWC->CHW of one image into out [1,3,H,W] (uint8, device)."""
        src = imread_device(self.labels[index]["im_file"], self.device)
        if g["r2"] != 1.0:
            raise NotImplementedError("a second resize inside LetterBox does not occur for rect / square val batches")
        h, w = g["resized"]
        H, W = g["new_shape"]
        ops.letterbox_u8(src[None].contiguous(), (w, h, g["left"], g["top"], H, W), swap_rb=True, chw=True, pad_value=114,
                         out=out)

    def __iter__(self):
        for bi in range(len(self)):
            idx = np.nonzero(self.batch == bi)[0]
            geos = [self.geometry(int(i)) for i in idx]
            H, W = geos[0]["new_shape"]
            img = torch.empty((len(idx), 3, H, W), dtype=torch.uint8, device=self.device)
            cls, boxes, bidx = [], [], []
            for k, (i, g) in enumerate(zip(idx, geos)):
                self.load_sample_image(int(i), g, img[k:k + 1])
                c, b = self.sample_labels(int(i), g)
                cls.append(c)
                boxes.append(b)
                bidx.append(np.full((len(c),), k, dtype=np.float32))
            cls_t = torch.from_numpy(np.concatenate(cls, 0))
            if cls_t.numel() == 0:
                cls_t = cls_t.reshape(0)                  # Format yields torch.zeros(0) for a label-free image
            yield {"img": img, "cls": cls_t,
                   "bboxes": torch.from_numpy(np.concatenate(boxes, 0)), "batch_idx": torch.from_numpy(np.concatenate(bidx, 0)),
                   "im_file": [self.labels[int(i)]["im_file"] for i in idx], "ori_shape": [g["ori_shape"] for g in geos],
                   "resized_shape": [g["new_shape"] for g in geos], "ratio_pad": [g["ratio_pad"] for g in geos]}


def build_yolo_dataset(cfg: dict, img_path, batch, data, mode="val", rect=False, stride=32) -> YOLODataset:
    """data/build.py:96-115 for the validation mode."""
    if mode != "val":
        raise NotImplementedError("specyolo builds validation datasets only (training is outside this package)")
    return YOLODataset(img_path, imgsz=int(cfg.get("imgsz", 640)), batch_size=batch, rect=bool(cfg.get("rect", False) or rect),
                       stride=int(stride), pad=0.5, data=data, single_cls=bool(cfg.get("single_cls", False)),
                       classes=cfg.get("classes"))
