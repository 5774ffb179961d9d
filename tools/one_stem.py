"""Times the stem pair (space-to-depth + 2x2 conv with blocked output, then the 3x3/s2 conv as a 2x2 conv over the blocked
tensor): python tools/one_stem.py B H W [iters]"""
import sys, math
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "spectrogram-yolov11_b200"))
from specyolo import ops
B, H, W = map(int, sys.argv[1:4]); iters = int(sys.argv[4]) if len(sys.argv) > 4 else 20
gen = torch.Generator().manual_seed(0)
x = (torch.rand((B, 3, H, W), generator=gen) * 255).to(torch.uint8).cuda()
w0 = torch.randn((32, 3, 3, 3), generator=gen) * 0.3; w1 = torch.randn((64, 32, 3, 3), generator=gen) * 0.08
pc0 = ops.fold_pack(w0.cuda(), torch.zeros(32).cuda(), None, 0.0, 2, 1, 1, 1, True)
pc1 = ops.pack_from_blocked(w1.cuda(), torch.zeros(64).cuda(), None, 0.0, True)
def ev(): return torch.cuda.Event(enable_timing=True)
for _ in range(60):      # long warm-up: the clocks ramp over tens of milliseconds
    xb = ops.stem_conv(x, pc0, blocked_out=True); y = ops.conv2d(xb, pc1)
torch.cuda.synchronize()
e = [ev() for _ in range(4)]
ts = [0.0, 0.0, 0.0]
for _ in range(iters):
    e[0].record(); s = ops.stem_space_to_depth(x); e[1].record()
    xb = ops.conv2d(s, pc0.s2d["u8"], blocked_out=True); e[2].record()
    y = ops.conv2d(xb, pc1); e[3].record()
    torch.cuda.synchronize()
    for i in range(3): ts[i] += e[i].elapsed_time(e[i + 1]) / iters
print(f"stem B{B} {H}x{W}: s2d {ts[0]*1e3:.1f} us, conv0 (16->32 k2, blocked out) {ts[1]*1e3:.1f} us, conv1 (128->64 k2) {ts[2]*1e3:.1f} us")
if ops.stem_pair_ok(x, pc0, pc1):
    for _ in range(30): yf = ops.stem_pair(x, pc0, pc1)
    torch.cuda.synchronize()
    a, b = ev(), ev()
    a.record()
    for _ in range(iters): yf = ops.stem_pair(x, pc0, pc1)
    b.record(); torch.cuda.synchronize()
    print(f"   fused stem pair: {a.elapsed_time(b)/iters*1e3:.1f} us   max |fused - layered| = {(yf.float()-y.float()).abs().max().item():.4f}")
