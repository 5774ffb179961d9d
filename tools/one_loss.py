"""Times the CUDA criterion (forward + backward into the head maps) against the reference's own v8DetectionLoss on the
same GPU: python tools/one_loss.py [B] [nc] [boxes per image]"""
import sys, time
from pathlib import Path
from types import SimpleNamespace
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "spectrogram-yolov11_b200")); sys.path.insert(0, str(ROOT / "tests"))
from oracle.loss_ref import loss_case
from specyolo.utils.loss import v8DetectionLoss
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nc = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ng = int(sys.argv[3]) if len(sys.argv) > 3 else 8
feats, batch = loss_case(5, B, 640, 640, nc, [ng] * B)
from test_loss_oracle import _FakeModel
crit = v8DetectionLoss(_FakeModel(nc))
f = [x.cuda().requires_grad_(True) for x in feats]
def step(c):
    for x in f: x.grad = None
    total, items = c(f, batch)
    total.backward()
    return items
def timeit(c, n=20):
    for _ in range(3): step(c)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): it = step(c)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, it
ms, it = timeit(crit)
print(f"specyolo criterion  B={B} nc={nc} {ng} boxes/image: {ms:.3f} ms per forward+backward (host wall), items {it.tolist()}")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
from specyolo.utils.loss import _DetLossFn, pack_targets
try:
    from oracle import ref_loader
    ref_loader.import_reference()
    from ultralytics.nn.tasks import DetectionModel as RefModel
    from ultralytics.utils.loss import v8DetectionLoss as RefLoss
    m = RefModel(str(Path(ref_loader.REFERENCE_ROOT) / "ultralytics/cfg/models/11/yolo11n.yaml"), nc=nc, verbose=False).cuda()
    m.args = SimpleNamespace(box=7.5, cls=0.5, dfl=1.5)
    rc = RefLoss(m)
    ms_r, it_r = timeit(rc, 10)
    print(f"reference criterion (PyTorch eager, same GPU): {ms_r:.3f} ms per forward+backward, items {it_r.tolist()}  -> {ms_r / ms:.1f}x")
except Exception as e:  # noqa
    print("reference criterion unavailable:", repr(e)[:200])
