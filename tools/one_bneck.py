"""Runs one fused-Bottleneck shape repeatedly (timing / ncu / SPECYOLO_BP_DBG=1 phase timers):
python tools/one_bneck.py C Cmid H W B [iters] [fused=1]"""
import sys, math, os
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "spectrogram-yolov11_b200"))
from specyolo import ops
C, Cm, H, W, B = map(int, sys.argv[1:6])
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 20
fused = int(sys.argv[7]) if len(sys.argv) > 7 else 1
gen = torch.Generator().manual_seed(0)
w1 = torch.randn((Cm, C, 3, 3), generator=gen) * math.sqrt(2.0 / (9 * C))
w2 = torch.randn((C, Cm, 3, 3), generator=gen) * math.sqrt(2.0 / (9 * Cm))
pc1 = ops.fold_pack(w1.cuda(), torch.zeros(Cm).cuda(), None, 0.0, 1, 1, 1, 1, True)
pc2 = ops.fold_pack(w2.cuda(), torch.zeros(C).cuda(), None, 0.0, 1, 1, 1, 1, True)
# as in C3k2: x = channels [C, 2C) of a 3C-channel concat buffer, y = channels [2C, 3C)
bufs = [ops.new_act(B, 3 * C, H, W, "cuda").normal_() for _ in range(3)]
def run(i):
    b = bufs[i % 3]
    x, y = b[:, C:2 * C], b[:, 2 * C:]
    if fused:
        ops.bottleneck(x, pc1, pc2, True, out=y)
    else:
        ops.conv2d(ops.conv2d(x, pc1), pc2, out=y, residual=x)
dbg = os.environ.pop("SPECYOLO_BP_DBG", None)
for i in range(3):
    run(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda._sleep(int(2e8))
e0.record()
for i in range(iters):
    run(i)
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / iters * 1e-3
fl = 2.0 * B * H * W * 9 * (C * Cm + Cm * C)
by = 2.0 * 2 * B * C * H * W
print(f"bottleneck {C}->{Cm}->{C} {H}x{W} B{B} fused={fused}: {t*1e6:.1f} us  {fl/t/1e12:.1f} TF/s  {by/t/1e9:.0f} GB/s")
if dbg and fused:
    os.environ["SPECYOLO_BP_DBG"] = "1"
    run(0); torch.cuda.synchronize()
