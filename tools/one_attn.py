"""Times the PSA attention kernel: python tools/one_attn.py B H W heads [iters]"""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "spectrogram-yolov11_b200"))
from specyolo import ops
B, H, W, heads = map(int, sys.argv[1:5]); iters = int(sys.argv[5]) if len(sys.argv) > 5 else 20
qkv = ops.new_act(B, heads * 128, H, W, "cuda").normal_()
pe_w = torch.randn(heads * 64, 9, device="cuda") * 0.2; pe_b = torch.randn(heads * 64, device="cuda") * 0.1
out = ops.new_act(B, heads * 64, H, W, "cuda")
for _ in range(3): ops.psa_attention(qkv, heads, 32, 64, 32 ** -0.5, pe_w, pe_b, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters): ops.psa_attention(qkv, heads, 32, 64, 32 ** -0.5, pe_w, pe_b, out=out)
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / iters * 1e-3
N = H * W
print(f"attention B{B} N{N} heads{heads}: {t*1e6:.1f} us  {2.0*B*heads*N*N*96/t/1e12:.1f} TF/s")
