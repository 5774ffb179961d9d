"""Small end-to-end pass for compute-sanitizer: IQ -> STFT -> detector (eager, no graph) -> NMS, imgsz 128, batch 2."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "spectrogram-yolov11_b200"))
import specyolo
from specyolo.nn.init import synth_images, synth_iq, synth_state_dict
for cfg, nc in (("yolo11s_fusion_sand3_new.yaml", 2), ("yolo11n.yaml", 80)):
    yolo = specyolo.YOLO(cfg, nc=nc)
    yolo.load_state_dict(synth_state_dict(yolo.model, seed=0)); yolo.to("cuda")
    x = synth_images(2, 128, seed=1, dtype=torch.uint8).cuda()
    r = yolo.predict(x, conf=0.05, use_graph=False)
    torch.cuda.synchronize()
    print(cfg, [len(a) for a in r])
iq = synth_iq(1, 1 << 16, seed=2).cuda()
img = specyolo.ops.iq_to_letterbox(iq, out_hw=(128, 128))
torch.cuda.synchronize()
print("ok", float(img.float().mean()))
