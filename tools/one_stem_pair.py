import sys
from pathlib import Path; R = Path(__file__).resolve().parent.parent; sys.path.insert(0, str(R)); sys.path.insert(0, str(R / "spectrogram-yolov11_b200"))
import torch
from specyolo import ops
gen = torch.Generator().manual_seed(0)
x = (torch.rand((64, 3, 640, 640), generator=gen) * 255).to(torch.uint8).cuda()
w0 = torch.randn((32, 3, 3, 3), generator=gen) * 0.3; w1 = torch.randn((64, 32, 3, 3), generator=gen) * 0.08
pc0 = ops.fold_pack(w0.cuda(), torch.zeros(32).cuda(), None, 0.0, 2, 1, 1, 1, True)
pc1 = ops.pack_from_blocked(w1.cuda(), torch.zeros(64).cuda(), None, 0.0, True)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 5): ops.stem_pair(x, pc0, pc1)
torch.cuda.synchronize()
