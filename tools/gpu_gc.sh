#!/bin/bash
# *_GC sibling config: BottleNect / FGM kernels vs the oracle, the whole model vs the real reference's fixture, timing
mkdir -p gpurun_out
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_kernels.py -k "bottlenect" -x > gpurun_out/t_gc.log 2>&1; echo "gc tests exit $?"; tail -25 gpurun_out/t_gc.log
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_model.py -k "golden" -x > gpurun_out/t_golden.log 2>&1; echo "golden exit $?"; tail -15 gpurun_out/t_golden.log
timeout 600 python - <<'PY'
import sys, torch
sys.path.insert(0, "spectrogram-yolov11_b200")
from specyolo import ops
from specyolo.nn.modules import BottleNect, MSCSpatialAttention
def timeit(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
m = BottleNect(32).cuda()
x = ops.new_act(64, 32, 160, 160, "cuda").normal_()
o = ops.new_act(64, 32, 160, 160, "cuda")
print(f"BottleNect/FGM B=64 c=32 160x160: {timeit(lambda: m(x, out=o)):.1f} us")
m2 = MSCSpatialAttention(64).cuda()
x2 = ops.new_act(64, 64, 80, 80, "cuda").normal_()
o2 = ops.new_act(64, 64, 80, 80, "cuda")
print(f"MSCSpatialAttention B=64 c=64 80x80: {timeit(lambda: m2(x2, out=o2)):.1f} us")
PY
timeout 300 python tools/one_gc.py 64 32 160 10 > gpurun_out/plain_gc.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"gc_" -s 16 -c 4 --csv --log-file gpurun_out/r02_gc_kernels.csv python tools/one_gc.py 64 32 160 3 > gpurun_out/ncu_gc.log 2>&1
echo "ncu gc exit $?"; grep -E "gpu__time_duration|issue_active" gpurun_out/r02_gc_kernels.csv | awk -F'","' '{print $5, $(NF-2), $NF}'
