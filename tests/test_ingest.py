"""SURVEY §8 f3 (second half): image-file ingest — LoadImagesAndVideos for image sources with JPEG decode on the GPU."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "spectrogram-yolov11_b200"))
cv2 = pytest.importorskip("cv2")


def _spectrogram_like(h, w, seed):
    g = np.random.default_rng(seed)
    img = np.clip(g.normal(64, 12, (h, w)), 0, 255)
    for _ in range(5):
        y0, x0 = int(g.integers(0, h - 20)), int(g.integers(0, w - 40))
        img[y0:y0 + int(g.integers(8, 20)), x0:x0 + int(g.integers(20, 40))] += g.uniform(90, 150)
    img = cv2.GaussianBlur(np.clip(img, 0, 255).astype(np.uint8), (5, 5), 0)
    return cv2.applyColorMap(img, cv2.COLORMAP_VIRIDIS)        # HWC BGR, smooth colours like a rendered spectrogram


def _write_files(d, synth=False):
    files = {}
    for k, (h, w, ext) in enumerate([(240, 320, "jpg"), (240, 320, "png"), (200, 280, "jpeg"), (320, 256, "bmp"), (192, 320, "jpg")]):
        if synth:       # crops of the images the synthetic weights were calibrated on: they produce detections
            from specyolo.nn.init import synth_images
            im = (synth_images(1, 320, seed=60 + k)[0].permute(1, 2, 0).numpy() * 255).round().astype(np.uint8)[:h, :w, ::-1]
            im = np.ascontiguousarray(im)
        else:
            im = _spectrogram_like(h, w, 50 + k)
        p = d / f"img_{k}.{ext}"
        assert cv2.imwrite(str(p), im)
        files[str(p)] = im
    (d / "notes.md").write_text("not an image")
    return files


def test_loader_file_discovery(tmp_path):
    from specyolo.data import LoadImagesAndVideos

    files = _write_files(tmp_path)
    ld = LoadImagesAndVideos(str(tmp_path), batch=2)
    assert ld.nf == 5 and len(ld) == 3 and ld.files == sorted(files)           # the .md file is ignored
    assert LoadImagesAndVideos(str(tmp_path / "*.jpg"), batch=4).nf == 2
    shards = [LoadImagesAndVideos(str(tmp_path), batch=2, rank=r, world=3).files for r in range(3)]       # multi-GPU split
    assert sum(shards, []) == sorted(files) and [len(x) for x in shards] == [2, 2, 1]
    (tmp_path / "list.txt").write_text("img_1.png\n" + str(tmp_path / "img_0.jpg") + "\n")
    assert LoadImagesAndVideos(str(tmp_path / "list.txt")).nf == 2
    with pytest.raises(FileNotFoundError):
        LoadImagesAndVideos(str(tmp_path / "missing.jpg"))
    (tmp_path / "clip.mp4").write_bytes(b"0")
    with pytest.raises(FileNotFoundError):                                     # loaders.py:444-445: "Failed to open video"
        LoadImagesAndVideos(str(tmp_path / "clip.mp4"))
    (tmp_path / "clip.mp4").unlink()
    empty = tmp_path / "empty"
    empty.mkdir()
    with pytest.raises(FileNotFoundError):
        LoadImagesAndVideos(str(empty))


def _write_video(path, n, hw=(48, 64), seed=0):
    h, w = hw
    vw = cv2.VideoWriter(str(path), cv2.VideoWriter_fourcc(*"MJPG"), 10, (w, h))
    assert vw.isOpened()
    for i in range(n):
        vw.write(_spectrogram_like(h, w, seed * 100 + i))
    vw.release()


@pytest.mark.parametrize("vid_stride,batch", [(1, 4), (2, 3), (3, 2)])
def test_loader_videos_match_reference(tmp_path, vid_stride, batch):
    """Video sources (loaders.py:388-412, 440-446): images first, then every vid_stride-th frame of every video, batches
    that never mix the two — the same (paths, frames, info) sequence as the REAL reference's loader on the same files."""
    from oracle import ref_loader

    if not ref_loader.reference_available():
        pytest.skip("reference package not available")
    ref_loader.import_reference()
    from ultralytics.data.loaders import LoadImagesAndVideos as RefLoader

    from specyolo.data import LoadImagesAndVideos

    for k in range(3):
        assert cv2.imwrite(str(tmp_path / f"img_{k}.png"), _spectrogram_like(40 + 8 * k, 56, k))
    _write_video(tmp_path / "a_clip.avi", 7, seed=1)
    _write_video(tmp_path / "b_clip.avi", 5, (32, 48), seed=2)
    mine = LoadImagesAndVideos(str(tmp_path), batch=batch, vid_stride=vid_stride, device="cpu")
    ref = RefLoader(str(tmp_path), batch=batch, vid_stride=vid_stride)
    assert mine.files == ref.files and mine.nf == ref.nf and mine.ni == ref.ni and mine.video_flag == ref.video_flag
    assert len(mine) == len(ref)
    got, want = list(mine), list(ref)
    assert len(got) == len(want)
    for (p1, i1, s1), (p2, i2, s2) in zip(got, want):
        assert p1 == p2 and s1 == s2
        assert len(i1) == len(i2)
        for a, b in zip(i1, i2):
            assert a.dtype == torch.uint8 and np.array_equal(a.numpy(), b)
    n_frames = sum(len(p) for p, _, _ in got) - 3
    assert n_frames == 7 // vid_stride + 5 // vid_stride
    assert [len(p) for p, _, _ in list(mine)] == [len(p) for p, _, _ in got]      # a second pass restarts the videos
    shards = [LoadImagesAndVideos(str(tmp_path), batch=2, rank=r, world=2, device="cpu").files for r in range(2)]
    assert sorted(sum(shards, [])) == sorted(mine.files) and all(any(f.endswith(".avi") for f in s) for s in shards)


@pytest.mark.gpu
def test_predict_from_video(tmp_path):
    """YOLO.predict(directory with a video, vid_stride=2): one Results per image and per kept frame, equal to predict on
    the same frames handed over as ndarrays."""
    import specyolo
    from specyolo.nn.init import synth_state_dict

    assert cv2.imwrite(str(tmp_path / "img_0.png"), _spectrogram_like(160, 200, 3))
    _write_video(tmp_path / "clip.avi", 6, (120, 160), seed=4)
    frames = [cv2.imread(str(tmp_path / "img_0.png"))]
    cap = cv2.VideoCapture(str(tmp_path / "clip.avi"))
    k = 0
    while True:
        ok, im = cap.read()
        if not ok:
            break
        k += 1
        if k % 2 == 0:
            frames.append(im)
    yolo = specyolo.YOLO("yolo11s_fusion_sand3_new.yaml", nc=2)
    yolo.load_state_dict(synth_state_dict(yolo.model, seed=0))
    yolo.to("cuda")
    res = yolo.predict(str(tmp_path), conf=0.25, iou=0.7, imgsz=160, batch=2, vid_stride=2)
    assert len(res) == len(frames) == 4
    assert res[0].path.endswith("img_0.png") and all(r.path.endswith("clip.avi") for r in res[1:])
    for r, im in zip(res, frames):
        assert tuple(r.orig_shape) == im.shape[:2]
        want = yolo.predict([im], conf=0.25, iou=0.7, imgsz=160)[0]
        assert torch.equal(r.boxes.data, want.boxes.data)


@pytest.mark.gpu
def test_gpu_decode_matches_cv2_imread(tmp_path):
    """PNG / BMP: bit-exact (same decoder as the reference, on the host).  JPEG cannot be bit-exact across decoders:
    with 4:4:4 sampling nvJPEG and libjpeg-turbo differ only by IDCT / colour-conversion rounding (mean |d| < 0.75, all
    samples within 4 grey levels); with the default 4:2:0 files the chroma up-sampling filters differ as well (libjpeg's
    "fancy" triangle filter vs nvJPEG's), so the bound is statistical: mean |d| < 2.5, 99 % of the samples within 12."""
    from specyolo.data import LoadImagesAndVideos, imread_device

    files = _write_files(tmp_path)
    for p in sorted(files):
        ref = cv2.imread(p)
        got = imread_device(p).cpu().numpy()
        assert got.shape == ref.shape and got.dtype == np.uint8
        d = np.abs(got.astype(np.int32) - ref.astype(np.int32))
        if p.endswith(("png", "bmp")):
            assert d.max() == 0
        else:
            assert d.mean() < 2.5 and (d <= 12).mean() > 0.99, (p, d.mean(), d.max())
    p444 = tmp_path / "full_chroma.jpg"
    cv2.imwrite(str(p444), _spectrogram_like(240, 320, 7), [cv2.IMWRITE_JPEG_QUALITY, 95, cv2.IMWRITE_JPEG_SAMPLING_FACTOR,
                                                            cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444])
    d = np.abs(imread_device(str(p444)).cpu().numpy().astype(np.int32) - cv2.imread(str(p444)).astype(np.int32))
    assert d.mean() < 0.75 and d.max() <= 4, (d.mean(), d.max())
    p444.unlink()
    seen = []
    for paths, imgs, info in LoadImagesAndVideos(str(tmp_path), batch=2):
        assert len(paths) == len(imgs) == len(info) <= 2
        assert all(i.is_cuda and i.dtype == torch.uint8 and i.shape[2] == 3 for i in imgs)
        seen += paths
    assert seen == sorted(files)
    gray = tmp_path / "gray.jpg"
    cv2.imwrite(str(gray), _spectrogram_like(64, 96, 1)[:, :, 0])
    g = imread_device(str(gray)).cpu().numpy()
    assert g.shape == (64, 96, 3) and np.abs(g.astype(int) - cv2.imread(str(gray)).astype(int)).max() <= 2


@pytest.mark.gpu
def test_predict_from_files(tmp_path):
    """YOLO.predict(directory) == predict(list of the same decoded pixels as ndarrays), with path / orig_shape attached,
    batch by batch through the streaming loop; for the lossless files that is predict(cv2.imread(...)) exactly."""
    import specyolo
    from specyolo.data import imread_device
    from specyolo.nn.init import synth_state_dict

    files = _write_files(tmp_path, synth=True)
    yolo = specyolo.YOLO("yolo11s_fusion_sand3_new.yaml", nc=2)
    yolo.load_state_dict(synth_state_dict(yolo.model, seed=0))
    yolo.to("cuda")
    res = yolo.predict(str(tmp_path), conf=0.25, iou=0.7, imgsz=320, batch=2)
    assert [r.path for r in res] == sorted(files)
    streamed = [r for b in yolo.predict(str(tmp_path), stream=True, conf=0.25, iou=0.7, imgsz=320, batch=2) for r in b]
    assert [r.path for r in streamed] == sorted(files)
    paths = sorted(files)
    for k in range(0, len(paths), 2):              # the loader's batches: LetterBox(auto) depends on the batch's shapes
        pix = [cv2.imread(p) if p.endswith(("png", "bmp")) else imread_device(p).cpu().numpy() for p in paths[k:k + 2]]
        ref = yolo.predict(pix, conf=0.25, iou=0.7, imgsz=320)
        for p, a, b, c in zip(paths[k:k + 2], ref, res[k:k + 2], streamed[k:k + 2]):
            assert b.orig_shape == a.orig_shape == cv2.imread(p).shape[:2]
            assert torch.equal(b.boxes.data, c.boxes.data)
            assert torch.equal(a.boxes.data, b.boxes.data)
    assert sum(len(r) for r in res) > 0


def _jpeg_with_orientation(path, im_bgr, orientation):
    from PIL import Image

    pil = Image.fromarray(np.ascontiguousarray(im_bgr[:, :, ::-1]))
    ex = Image.Exif()
    if orientation:
        ex[274] = orientation
    pil.save(str(path), quality=95, subsampling=0, exif=ex.tobytes() if orientation else b"")


def test_exif_orientation_parser(tmp_path):
    """the host-side EXIF walk finds the orientation tag cv2.imread honours (both byte orders via PIL/handmade)."""
    from specyolo.data.loaders import jpeg_exif_orientation

    im = _spectrogram_like(48, 80, 3)
    for o in (None, 1, 2, 3, 6, 8):
        p = tmp_path / f"o{o}.jpg"
        _jpeg_with_orientation(p, im, o)
        data = p.read_bytes()
        assert jpeg_exif_orientation(data) == (o or 1)
        ref = cv2.imread(str(p))
        assert ref.shape[:2] == ((80, 48) if o in (6, 8) else (48, 80))
    # big-endian TIFF header, hand-made APP1 segment
    import struct
    tiff = b"MM\x00\x2a" + struct.pack(">I", 8) + struct.pack(">H", 1) + struct.pack(">HHI", 0x0112, 3, 1) + \
        struct.pack(">H", 6) + b"\x00\x00" + struct.pack(">I", 0)
    app1 = b"Exif\x00\x00" + tiff
    base = (tmp_path / "oNone.jpg").read_bytes()
    data = base[:2] + b"\xff\xe1" + struct.pack(">H", len(app1) + 2) + app1 + base[2:]
    assert jpeg_exif_orientation(data) == 6
    assert jpeg_exif_orientation(b"\xff\xd8\xff\xd9") == 1 and jpeg_exif_orientation(b"") == 1


@pytest.mark.gpu
def test_exif_rotated_jpeg_matches_cv2(tmp_path):
    """A JPEG tagged orientation=6 must come back in cv2.imread's (rotated) frame, bit for bit (host decoder route)."""
    from specyolo.data.loaders import imread_device

    im = _spectrogram_like(96, 160, 9)
    for o in (3, 6, 8):
        p = tmp_path / f"rot{o}.jpg"
        _jpeg_with_orientation(p, im, o)
        ref = cv2.imread(str(p))
        got = imread_device(str(p)).cpu().numpy()
        assert got.shape == ref.shape and np.array_equal(got, ref), o
