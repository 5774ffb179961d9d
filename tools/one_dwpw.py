"""Times the fused DWConv3x3 + Conv1x1 kernel against the two separate launches: python tools/one_dwpw.py C cout H W B [iters]"""
import sys, math
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "spectrogram-yolov11_b200"))
from specyolo import ops
C, cout, H, W, B = map(int, sys.argv[1:6]); iters = int(sys.argv[6]) if len(sys.argv) > 6 else 20
gen = torch.Generator().manual_seed(0)
wd = torch.randn((C, 1, 3, 3), generator=gen) * 0.3; bd = torch.zeros(C)
wp = torch.randn((cout, C, 1, 1), generator=gen) * math.sqrt(2.0 / C)
pw = ops.fold_pack(wp.cuda(), torch.zeros(cout).cuda(), None, 0.0, 1, 0, 1, 1, True)
pd = ops.fold_pack(wd.cuda(), bd.cuda(), None, 0.0, 1, 1, 1, C, True)
dw_w = wd.view(C, 9).t().contiguous().cuda(); dw_b = bd.cuda()
xs = [ops.new_act(B, C, H, W, "cuda").normal_() for _ in range(3)]
outs = [ops.new_act(B, cout, H, W, "cuda") for _ in range(3)]
def timeit(f):
    for i in range(3): f(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): f(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
t_f = timeit(lambda i: ops.dwconv_pwconv(xs[i % 3], dw_w, dw_b, pw, out=outs[i % 3]))
t_s = timeit(lambda i: ops.conv2d(ops.conv2d(xs[i % 3], pd), pw, out=outs[i % 3]))
by = 2.0 * B * H * W * (C + cout)
print(f"dwpw C{C}->{cout} {H}x{W} B{B}: fused {t_f:.1f} us ({by/t_f/1e3:.0f} GB/s)   separate {t_s:.1f} us")
