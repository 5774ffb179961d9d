"""Runs one conv shape repeatedly (for ncu / timing): python tools/one_conv.py cin cout k s H W B [iters]"""
import sys, math
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "spectrogram-yolov11_b200"))
from specyolo import ops
cin, cout, k, s, H, W, B = map(int, sys.argv[1:8])
iters = int(sys.argv[8]) if len(sys.argv) > 8 else 20
g = int(sys.argv[9]) if len(sys.argv) > 9 else 1
d = int(sys.argv[10]) if len(sys.argv) > 10 else 1
res = int(sys.argv[11]) if len(sys.argv) > 11 else 0          # 1: residual add
cbuf = int(sys.argv[12]) if len(sys.argv) > 12 else 0         # >0: output is the last `cout` channels of a cbuf-channel buffer
gen = torch.Generator().manual_seed(0)
w = torch.randn((cout, cin // g, k, k), generator=gen) * math.sqrt(2.0 / (cin // g * k * k))
pc = ops.fold_pack(w.cuda(), torch.zeros(cout).cuda(), None, 0.0, s, (d * (k - 1) + 1) // 2, d, g, True)
xs = [ops.new_act(B, cin, H, W, "cuda").normal_() for _ in range(3)]   # rotate inputs (> L2 for the big shapes)
Ho, Wo = pc.out_hw(H, W)
bufs = [ops.new_act(B, cbuf or cout, Ho, Wo, "cuda") for _ in range(3)]
outs = [b[:, -cout:] for b in bufs]
rs = [ops.new_act(B, cout, Ho, Wo, "cuda").normal_() if res else None for _ in range(3)]
for i in range(3):
    ops.conv2d(xs[i % 3], pc, out=outs[i % 3], residual=rs[i % 3])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda._sleep(int(2e8))
e0.record()
for i in range(iters):
    ops.conv2d(xs[i % 3], pc, out=outs[i % 3], residual=rs[i % 3])
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / iters * 1e-3
fl = 2.0 * B * Ho * Wo * cout * (cin // g) * k * k
by = 2.0 * (B * cin * H * W + B * cout * Ho * Wo)
print(f"conv {cin}->{cout} k{k} s{s} g{g} d{d} {H}x{W} B{B} res{res} cbuf{cbuf}: {t*1e6:.1f} us  {fl/t/1e12:.1f} TF/s  {by/t/1e9:.0f} GB/s")
