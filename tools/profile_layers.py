"""Per-launch IN-GRAPH timing of one step (specyolo/utils/kprof.py: single-stream, PDL-off capture, CUPTI durations)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "spectrogram-yolov11_b200"))
import specyolo
from specyolo.nn.init import synth_images, synth_state_dict
from specyolo.utils import kprof

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = sys.argv[2] if len(sys.argv) > 2 else "yolo11s_fusion_sand3_new.yaml"
nc = 2 if "fusion" in cfg else 80
imgsz = int(sys.argv[3]) if len(sys.argv) > 3 else 640
yolo = specyolo.YOLO(cfg, nc=nc)
yolo.load_state_dict(synth_state_dict(yolo.model, seed=0)); yolo.to("cuda"); yolo.fuse()
x = synth_images(B, imgsz, seed=0, dtype=torch.uint8).cuda()
prof = kprof.profile_graph(lambda: yolo.model.detect_fused(x))
tot = sum(c["us"] for c in prof["calls"])
print(f"B={B} cfg={cfg} source={prof['source']} note={prof['note']} kernels {prof['kernels_per_step']} "
      f"sum {tot / 1e3:.3f} ms serial graph {prof['serial_ms']:.3f} ms over {len(prof['calls'])} calls")
for i, c in enumerate(prof["calls"]):
    t = c["us"] * 1e-6
    print(f"{i:3d} {c['us']:8.1f} us  {c['label']:78s} {c['flops'] / t / 1e12 if c['flops'] else 0:7.1f} TF/s "
          f"{c['bytes'] / t / 1e9 if c['bytes'] else 0:7.0f} GB/s")
