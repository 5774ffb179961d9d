"""Image-file ingest: `LoadImagesAndVideos` for image sources, decoded straight into device memory.

Mirrors ultralytics/data/loaders.py:284-448 (constructor arguments, file discovery rules, images before videos,
`(paths, imgs, info)` batches that never mix the two, `vid_stride`, `__len__`).  JPEG files are decoded by
nvJPEG into HBM (`specyolo_jpeg_decode_bgr`: HWC BGR uint8, the layout of `cv2.imread`) — the host only reads the file
bytes; the pixels never exist in host memory.  Other formats (PNG, BMP, TIFF, WebP ...) have no GPU decoder in this image:
they are decoded on the host with OpenCV exactly as the reference does (`cv2.imdecode`, loaders.py:406 / patches.py:20)
and uploaded.  Video files are demuxed and decoded on the host by `cv2.VideoCapture` as in the reference
(loaders.py:388-412, 440-446; this image has no NVDEC binding) — grab `vid_stride` frames, retrieve the last — and each
retrieved frame is uploaded as one more [H, W, 3] BGR uint8 device tensor; live streams (`LoadStreams`) are not rebuilt.
Every image of a batch is then letterboxed by `specyolo_letterbox_u8` (bit-exact with cv2's resize).
"""
from __future__ import annotations

import contextlib
import ctypes as C
import glob
import math
import os
from pathlib import Path
from typing import List, Tuple

import numpy as np
import torch

from .. import _lib

IMG_FORMATS = {"bmp", "dng", "jpeg", "jpg", "mpo", "png", "tif", "tiff", "webp", "pfm"}   # data/utils.py:38 (no HEIC)
VID_FORMATS = {"asf", "avi", "gif", "m4v", "mkv", "mov", "mp4", "mpeg", "mpg", "ts", "wmv", "webm"}


_POOL = None


def _pool():
    """Host threads of the decode stage (file read + entropy decoding), shared by every loader."""
    global _POOL
    if _POOL is None:
        from concurrent.futures import ThreadPoolExecutor

        _POOL = ThreadPoolExecutor(max_workers=max(1, min(16, (os.cpu_count() or 4))), thread_name_prefix="specyolo-ingest")
    return _POOL


def decode_jpeg(data: bytes, device="cuda") -> torch.Tensor:
    """JPEG bytes -> [H, W, 3] uint8 BGR CUDA tensor (nvJPEG, asynchronous on the current stream)."""
    _lib.init_device()
    lib = _lib.load()
    buf = (C.c_char * len(data)).from_buffer_copy(data)
    h, w, c = C.c_int(), C.c_int(), C.c_int()
    _lib.check(lib.specyolo_jpeg_info(C.addressof(buf), len(data), C.byref(h), C.byref(w), C.byref(c)))
    out = torch.empty((h.value, w.value, 3), dtype=torch.uint8, device=device)
    _lib.check(lib.specyolo_jpeg_decode_bgr(C.addressof(buf), len(data), out.data_ptr(), h.value, w.value, _lib.stream_ptr()))
    # nvjpegDecode has consumed the host bytes (Huffman stage) when it returns; `buf` may go away
    return out


def decode_jpeg_batch(blobs: List[bytes], backend: int = 2, device="cuda") -> List[torch.Tensor]:
    """n JPEG byte strings -> n [H, W, 3] uint8 BGR CUDA tensors in ONE nvjpegDecodeBatched call (backend 2: GPU-assisted
    Huffman, 3: hardware engine).  Raises RuntimeError when the back end rejects the batch (callers fall back to decode_jpeg)."""
    _lib.init_device()
    lib = _lib.load()
    n = len(blobs)
    bufs = [(C.c_char * len(b)).from_buffer_copy(b) for b in blobs]
    outs, widths = [], []
    for buf, b in zip(bufs, blobs):
        h, w, c = C.c_int(), C.c_int(), C.c_int()
        _lib.check(lib.specyolo_jpeg_info(C.addressof(buf), len(b), C.byref(h), C.byref(w), C.byref(c)))
        outs.append(torch.empty((h.value, w.value, 3), dtype=torch.uint8, device=device))
        widths.append(w.value)
    data = (C.c_void_p * n)(*[C.addressof(b) for b in bufs])
    sizes = (C.c_size_t * n)(*[len(b) for b in blobs])
    dst = (C.c_void_p * n)(*[o.data_ptr() for o in outs])
    ws = (C.c_int * n)(*widths)
    _lib.check(lib.specyolo_jpeg_decode_batch_bgr(data, sizes, dst, ws, n, int(backend), _lib.stream_ptr()))
    torch.cuda.current_stream().synchronize()          # the bitstreams (`bufs`) must outlive the decode
    return outs


def jpeg_exif_orientation(data: bytes) -> int:
    """EXIF orientation tag (0x0112) of a JPEG byte string, 1 when absent / unreadable.  Walks the marker segments up to
    the first scan; reads the TIFF header of the APP1 "Exif" segment (either byte order) and IFD0 only, like OpenCV's
    ExifReader does for `cv2.imread` (which rotates / flips the decoded pixels unless IMREAD_IGNORE_ORIENTATION)."""
    import struct

    n = len(data)
    i = 2
    while i + 4 <= n and data[i] == 0xFF:
        marker = data[i + 1]
        if marker in (0xD8, 0x01) or 0xD0 <= marker <= 0xD7:          # standalone markers
            i += 2
            continue
        if marker == 0xDA or marker == 0xD9:                            # start of scan / end of image: no EXIF ahead
            break
        seglen = struct.unpack(">H", data[i + 2:i + 4])[0]
        if marker == 0xE1 and data[i + 4:i + 10] == b"Exif\x00\x00":
            t = i + 10                                                  # TIFF header
            bo = {b"II": "<", b"MM": ">"}.get(data[t:t + 2])
            if bo is None or t + 8 > n:
                return 1
            ifd = t + struct.unpack(bo + "I", data[t + 4:t + 8])[0]
            if ifd + 2 > n:
                return 1
            for k in range(struct.unpack(bo + "H", data[ifd:ifd + 2])[0]):
                e = ifd + 2 + 12 * k
                if e + 12 > n:
                    break
                tag, typ = struct.unpack(bo + "HH", data[e:e + 4])
                if tag == 0x0112:
                    v = struct.unpack(bo + "H", data[e + 8:e + 10])[0] if typ == 3 else \
                        struct.unpack(bo + "I", data[e + 8:e + 12])[0]
                    return int(v) if 1 <= v <= 8 else 1
            return 1
        i += 2 + seglen
    return 1


def imread_device(path: str, device="cuda") -> torch.Tensor:
    """An image file -> [H, W, 3] uint8 BGR CUDA tensor; None-like failure raises (the reference warns and skips)."""
    data = Path(path).read_bytes()
    # nvJPEG ignores EXIF; cv2.imread (the reference, loaders.py:406) applies the orientation tag.  Files that carry a
    # non-trivial orientation go through the reference's own decoder so pixels, `exif_size` (dataset._image_hw) and
    # labels stay in one frame.
    if data[:3] == b"\xff\xd8\xff" and jpeg_exif_orientation(data) == 1:   # JPEG magic, whatever the suffix says
        try:
            return decode_jpeg(data, device)
        except RuntimeError:
            pass                                           # progressive / CMYK / arithmetic-coded streams: host decoder
    import cv2                                             # the reference's own decoder for everything else

    im = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
    if im is None:
        raise ValueError(f"Image Read Error {path}")
    return torch.from_numpy(np.ascontiguousarray(im)).to(device, non_blocking=True)


class LoadImagesAndVideos:
    """`for paths, imgs, info in LoadImagesAndVideos(path, batch)`: imgs = list of [H,W,3] uint8 BGR CUDA tensors."""

    def __init__(self, path, batch: int = 1, vid_stride: int = 1, device="cuda", rank: int = 0, world: int = 1):
        parent = None
        if isinstance(path, str) and Path(path).suffix == ".txt":      # *.txt with one source per line
            parent = Path(path).parent
            path = Path(path).read_text().splitlines()
        files: List[str] = []
        for p in sorted(path) if isinstance(path, (list, tuple)) else [path]:
            a = str(Path(p).absolute())                                # loaders.py:330-343
            if "*" in a:
                files.extend(sorted(glob.glob(a, recursive=True)))
            elif os.path.isdir(a):
                files.extend(sorted(glob.glob(os.path.join(a, "*.*"))))
            elif os.path.isfile(a):
                files.append(a)
            elif parent and (parent / p).is_file():
                files.append(str((parent / p).absolute()))
            else:
                raise FileNotFoundError(f"{p} does not exist")
        images = [f for f in files if f.split(".")[-1].lower() in IMG_FORMATS]
        videos = [f for f in files if f.split(".")[-1].lower() in VID_FORMATS]
        if world > 1:       # multi-GPU: files are independent, every rank takes a contiguous share, no collective
            from ..dist import shard_range

            b, e = shard_range(len(images), rank, world)
            images = images[b:e]
            b, e = shard_range(len(videos), rank, world)
            videos = videos[b:e]
        self.files = images + videos                                   # loaders.py:347-352: images first
        self.ni = len(images)
        self.nf = self.ni + len(videos)
        self.video_flag = [False] * self.ni + [True] * len(videos)
        self.mode = "video" if self.ni == 0 and videos else "image"
        self.vid_stride = vid_stride
        self.bs = batch
        self.device = device
        self.count = 0
        self.cap = None
        self.frame = self.frames = self.fps = 0
        if self.nf == 0 and world == 1:
            raise FileNotFoundError(f"No images or videos found in {path}. Supported formats are: images: {IMG_FORMATS} "
                                    f"videos: {VID_FORMATS}")
        if videos:
            self._new_video(videos[0])                                 # loaders.py:358-359: fails early on an unreadable file

    def __iter__(self):
        self.count = 0
        if self.cap is not None and self.frame:                        # a second pass starts every video from its first frame
            self.cap.release()
            self.cap = None
        return self

    def _new_video(self, path: str) -> None:
        """loaders.py:440-446."""
        import cv2

        self.frame = 0
        self.cap = cv2.VideoCapture(path)
        self.fps = int(self.cap.get(cv2.CAP_PROP_FPS))
        if not self.cap.isOpened():
            raise FileNotFoundError(f"Failed to open video {path}")
        self.frames = int(self.cap.get(cv2.CAP_PROP_FRAME_COUNT) / self.vid_stride)

    def _next_frame(self, paths, imgs, info) -> None:
        """One step of the reference's video branch (loaders.py:388-412): grab vid_stride frames, keep the last."""
        path = self.files[self.count]
        self.mode = "video"
        if self.cap is None or not self.cap.isOpened():
            self._new_video(path)
        ok = False
        for _ in range(self.vid_stride):
            ok = self.cap.grab()
            if not ok:
                break
        if ok:
            ok, im0 = self.cap.retrieve()
            if ok:
                self.frame += 1
                paths.append(path)
                imgs.append(torch.from_numpy(np.ascontiguousarray(im0)).to(self.device, non_blocking=True))
                info.append(f"video {self.count + 1}/{self.nf} (frame {self.frame}/{self.frames}) {path}: ")
                if self.frame == self.frames:                          # the container's frame count is reached
                    self.count += 1
                    self.cap.release()
        else:                                                          # end of stream (or a failed grab): next file
            self.count += 1
            self.cap.release()
            if self.count < self.nf:
                self._new_video(self.files[self.count])

    def __next__(self) -> Tuple[List[str], List[torch.Tensor], List[str]]:
        paths, imgs, info = [], [], []
        while len(imgs) < self.bs:
            if self.count >= self.nf:
                if imgs:
                    return paths, imgs, info                           # last partial batch
                raise StopIteration
            if self.video_flag[self.count]:
                self._next_frame(paths, imgs, info)
                continue
            # the next image files of the batch, decoded concurrently: the host part of a JPEG decode (file read + Huffman)
            # runs on the calling thread and releases the GIL, the library keeps one nvJPEG state per in-flight decode
            self.mode = "image"
            chunk = self.files[self.count:min(self.ni, self.count + (self.bs - len(imgs)))]
            on_gpu = torch.device(self.device).type == "cuda"      # device="cpu": file-discovery / ordering tests only
            stream = torch.cuda.current_stream() if on_gpu else None

            def load(path):
                with (torch.cuda.stream(stream) if on_gpu else contextlib.nullcontext()):
                    try:
                        return imread_device(path, self.device)
                    except ValueError:
                        return None                                    # loaders.py:438-439: warn and move on

            decoded = list(_pool().map(load, chunk)) if len(chunk) > 1 else [load(chunk[0])]
            for k, (path, im0) in enumerate(zip(chunk, decoded)):
                if im0 is not None:
                    paths.append(path)
                    imgs.append(im0)
                    info.append(f"image {self.count + k + 1}/{self.nf} {path}: ")
            self.count += len(chunk)
            if self.count >= self.ni and imgs:                         # loaders.py:431-432: images and frames never share a batch
                break
        return paths, imgs, info

    def __len__(self):
        return math.ceil(self.nf / self.bs)
