"""Which kernel does CUPTI not see?  yolo11s: per-op launch counts (recorder) vs kernel names in one profiled replay."""
import sys, json, tempfile, os
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "spectrogram-yolov11_b200"))
import specyolo
from specyolo import ops, _lib
from specyolo.nn.init import synth_images, synth_state_dict
from specyolo.utils import kprof
from torch.profiler import ProfilerActivity, profile
y = specyolo.YOLO("yolo11s.yaml", nc=80); y.load_state_dict(synth_state_dict(y.model, seed=3)); y.to("cuda"); y.fuse()
x = synth_images(2, 640, seed=0, dtype=torch.uint8).cuda()
fn = lambda: y.model.detect_fused(x)
os.environ["SPECYOLO_NO_PDL"] = "1"; ops.CONCURRENT = False
for _ in range(2): fn()
lib = _lib.load()
with kprof.OpRecorder() as rec:
    n0 = lib.specyolo_launch_count(); fn(); n = lib.specyolo_launch_count() - n0
print("launch_count", n, "recorded", sum(c["kernels"] for c in rec.calls))
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    fn(); torch.cuda.synchronize()
p = os.path.join(tempfile.mkdtemp(), "t.json"); prof.export_chrome_trace(p)
ev = json.load(open(p))["traceEvents"]
ks = sorted((e for e in ev if e.get("cat") == "kernel"), key=lambda e: e["ts"])
mine = [e for e in ks if "specyolo" in e["name"]]
print("kernels in trace", len(ks), "specyolo", len(mine))
i = 0
for c in rec.calls:
    names = [e["name"].split("(")[0][-60:] for e in mine[i:i + c["kernels"]]]
    print(c["kernels"], c["op"], c["label"][:60], "|", names)
    i += c["kernels"]
print("others:", sorted({e["name"][:80] for e in ks if "specyolo" not in e["name"]}))
