"""File ingest rate: N JPEG files (HxW) -> nvJPEG decode -> letterbox -> detector (predict from a directory), vs decode alone.
python tools/one_ingest.py N H W [batch]"""
import sys, time, tempfile
from pathlib import Path
import numpy as np, cv2, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "spectrogram-yolov11_b200"))
import specyolo
from specyolo.data import imread_device
from specyolo.nn.init import synth_images, synth_state_dict
N, H, W = map(int, sys.argv[1:4]); batch = int(sys.argv[4]) if len(sys.argv) > 4 else 64
d = Path(tempfile.mkdtemp())
imgs = (synth_images(8, max(H, W), seed=1).permute(0, 2, 3, 1).numpy() * 255).round().astype(np.uint8)[:, :H, :W, ::-1]
for k in range(N):
    cv2.imwrite(str(d / f"{k:05d}.jpg"), np.ascontiguousarray(imgs[k % 8]), [cv2.IMWRITE_JPEG_QUALITY, 90])
files = sorted(str(p) for p in d.glob("*.jpg"))
size = sum(Path(f).stat().st_size for f in files) / N
for _ in range(2):
    for f in files[:16]: imread_device(f)
torch.cuda.synchronize(); t0 = time.perf_counter()
for f in files: imread_device(f)
torch.cuda.synchronize(); t_gpu = time.perf_counter() - t0
from specyolo.data import LoadImagesAndVideos
for _ in LoadImagesAndVideos(str(d), batch=batch): pass
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in LoadImagesAndVideos(str(d), batch=batch): pass
torch.cuda.synchronize(); t_pool = time.perf_counter() - t0
t0 = time.perf_counter()
for f in files: cv2.imread(f)
t_cpu = time.perf_counter() - t0
from specyolo.data import decode_jpeg_batch
blobs = [Path(f).read_bytes() for f in files]
for be in (2, 3):
    try:
        for _ in range(2): decode_jpeg_batch(blobs[:batch], be)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for k in range(0, N, batch): outs = decode_jpeg_batch(blobs[k:k + batch], be)
        torch.cuda.synchronize(); tb = time.perf_counter() - t0
        dd = np.abs(outs[0].cpu().numpy().astype(int) - cv2.imread(files[(N - batch) if N >= batch else 0]).astype(int))
        print(f"nvjpegDecodeBatched backend {be}: {N/tb:.0f} img/s (bytes already in memory), mean |d| vs cv2 {dd.mean():.2f}")
    except Exception as e:
        print(f"nvjpegDecodeBatched backend {be}: not available ({str(e)[:120]})")
yolo = specyolo.YOLO("yolo11s_fusion_sand3_new.yaml", nc=2); yolo.load_state_dict(synth_state_dict(yolo.model, seed=0)); yolo.to("cuda")
yolo.predict(str(d), batch=batch, imgsz=max(H, W))          # graph capture
torch.cuda.synchronize(); t0 = time.perf_counter()
n = sum(len(b) for b in yolo.predict(str(d), batch=batch, imgsz=max(H, W), stream=True))
torch.cuda.synchronize(); t_e2e = time.perf_counter() - t0
print(f"{N} JPEGs {H}x{W} (~{size/1e3:.0f} kB each): nvJPEG decode into HBM {N/t_gpu:.0f} img/s from one host thread, {N/t_pool:.0f} img/s through the loader's thread pool, cv2.imread (1 thread) {N/t_cpu:.0f} img/s, "
      f"files -> boxes (predict from the directory, batch {batch}) {n/t_e2e:.0f} img/s")
