"""Times IQ -> letterboxed spectrogram: python tools/one_stft.py B [log2 L] [iters]"""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "spectrogram-yolov11_b200"))
from specyolo import ops
from specyolo.nn.init import synth_iq
B = int(sys.argv[1]); lg = int(sys.argv[2]) if len(sys.argv) > 2 else 20; iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20
iq = synth_iq(2, 1 << lg, seed=1).cuda()
iq = iq.repeat((B + 1) // 2, 1)[:B].contiguous()
out = torch.empty((B, 3, 640, 640), device="cuda", dtype=torch.bfloat16)
for _ in range(3): ops.iq_to_letterbox(iq, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters): ops.iq_to_letterbox(iq, out=out)
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / iters * 1e-3
alg = B * ((1 << lg) * 8 + 3 * 640 * 640 * 2)
print(f"stft B{B} L=2^{lg}: {t*1e6:.1f} us  {B/t:.0f} bursts/s  algorithmic (whole burst + image) {alg/t/1e9:.0f} GB/s")
