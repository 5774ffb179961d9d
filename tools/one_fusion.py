"""Times Fusion('ESChannel'): python tools/one_fusion.py B H W c k up0 [iters]   (up0=1: first input is a x2 upsample)"""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "spectrogram-yolov11_b200"))
from specyolo import ops
B, H, W, c, k, up0 = map(int, sys.argv[1:7]); iters = int(sys.argv[7]) if len(sys.argv) > 7 else 20
ups = [up0] + [0] * (k - 1)
xs = [ops.new_act(B, c, H >> u, W >> u, "cuda").normal_() for u in ups]
alpha = torch.ones(k * c, device="cuda"); gamma = torch.randn(k * c, device="cuda") * 0.1; beta = torch.randn(k * c, device="cuda") * 0.1
sab = torch.randn(18, device="cuda") * 0.2
out = ops.new_act(B, c, H, W, "cuda")
for _ in range(3): ops.fusion_eschannel(xs, ups, alpha, gamma, beta, 1e-5, sab, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters): ops.fusion_eschannel(xs, ups, alpha, gamma, beta, 1e-5, sab, out=out)
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / iters * 1e-3
by = 2.0 * (2 * sum(x.numel() for x in xs) + out.numel())
print(f"fusion B{B} {H}x{W} c{c} k{k} up{up0}: {t*1e6:.1f} us  {by/t/1e9:.0f} GB/s")
