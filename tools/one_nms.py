"""The bench's decode + NMS alone on the bench input (SPECYOLO_NMS_DBG=1 prints per-phase cycles)."""
import sys, os
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "spectrogram-yolov11_b200"))
import specyolo
from specyolo.nn.init import synth_images, synth_state_dict
yolo = specyolo.YOLO("yolo11s_fusion_sand3_new.yaml", nc=2)
yolo.load_state_dict(synth_state_dict(yolo.model, seed=0)); yolo.to("cuda")
x = synth_images(64, 640, seed=0, dtype=torch.uint8).cuda()
dbg = os.environ.pop("SPECYOLO_NMS_DBG", None)
for _ in range(2):
    out, cnt = yolo.model.detect_fused(x)
torch.cuda.synchronize()
if dbg:
    os.environ["SPECYOLO_NMS_DBG"] = "1"
out, cnt = yolo.model.detect_fused(x)
torch.cuda.synchronize()
print("detections", int(cnt.sum()), "max per image", int(cnt.max()))
