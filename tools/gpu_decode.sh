#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_kernels.py -k "detect_decode" -x > gpurun_out/t_decode.log 2>&1; echo "decode tests exit $?"; tail -6 gpurun_out/t_decode.log
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_model.py -x > gpurun_out/t_model.log 2>&1; echo "model tests exit $?"; tail -4 gpurun_out/t_model.log
timeout 600 python - <<'PY'
import sys, torch
sys.path.insert(0, "spectrogram-yolov11_b200")
from specyolo import ops
def timeit(f, n=30):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for (B, nc, S) in ((128, 80, 1280), (64, 2, 640), (64, 80, 640)):
    hw = [(S // s, S // s) for s in (8, 16, 32)]
    no_stride = (64 + nc + 3) // 4 * 4
    bufs = []
    for (h, w) in hw:
        t = torch.randn((B, h * w, no_stride), device="cuda") * 1.5
        t[..., 64:] -= 7.0
        bufs.append(t)
    us = timeit(lambda: ops.detect_decode(bufs, hw, [8.0, 16.0, 32.0], nc, want_dense=False, conf_thres=0.25))
    by = sum(t.numel() for t in bufs) * 4
    print(f"decode B={B} nc={nc} {S}^2: {us:.1f} us  (class logits {B * sum(h * w for h, w in hw) * nc * 4 / us / 1e3:.0f} GB/s, whole rows {by / us / 1e3:.0f} GB/s)")
PY
