import sys
from pathlib import Path; R = Path(__file__).resolve().parent.parent; sys.path.insert(0, str(R)); sys.path.insert(0, str(R / "spectrogram-yolov11_b200"))
import torch, torch.nn.functional as F
from specyolo import ops
B, H, W = 1, int(sys.argv[1]), int(sys.argv[2])
gen = torch.Generator().manual_seed(33)
u8 = (torch.rand((B, 3, H, W), generator=gen) * 255).round().to(torch.uint8)
w0 = torch.randn((32, 3, 3, 3), generator=gen) * 0.3; w1 = torch.randn((64, 32, 3, 3), generator=gen) * 0.08
b0 = torch.randn(32, generator=gen) * 0.1; b1 = torch.randn(64, generator=gen) * 0.1
pc0 = ops.fold_pack(w0.cuda(), b0.cuda(), None, 0.0, 2, 1, 1, 1, True)
pc1 = ops.pack_from_blocked(w1.cuda(), b1.cuda(), None, 0.0, True)
y = ops.stem_pair(u8.cuda(), pc0, pc1)
torch.cuda.synchronize()
bf = lambda t: t.to(torch.bfloat16).float()
y0 = bf(F.silu(F.conv2d(u8.float(), bf(w0 / 255.0), b0, 2, 1)))
ref = F.silu(F.conv2d(y0, bf(w1), b1, 2, 1))
d = (y.float().cpu() - ref).abs()
print("max err", d.max().item(), "ref max", ref.abs().max().item())
print("err per row", d.amax(dim=(0, 1, 3)))
print("err per col", d.amax(dim=(0, 1, 2)))
