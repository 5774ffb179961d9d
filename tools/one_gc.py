"""Runs BottleNect / FGM (the *_GC block) repeatedly: python tools/one_gc.py [B] [C] [H] [iters] (timing, ncu target)"""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "spectrogram-yolov11_b200"))
from specyolo import ops
from specyolo.nn.modules import BottleNect
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64; C = int(sys.argv[2]) if len(sys.argv) > 2 else 32
H = int(sys.argv[3]) if len(sys.argv) > 3 else 160; iters = int(sys.argv[4]) if len(sys.argv) > 4 else 10
m = BottleNect(C).cuda()
x = ops.new_act(B, C, H, H, "cuda").normal_()
o = ops.new_act(B, C, H, H, "cuda")
for _ in range(3): m(x, out=o)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters): m(x, out=o)
e1.record(); torch.cuda.synchronize()
print(f"BottleNect/FGM B={B} c={C} {H}x{H}: {e0.elapsed_time(e1) / iters * 1e3:.1f} us")
