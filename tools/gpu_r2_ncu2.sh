#!/bin/bash
# ncu evidence for the kernels added late in round 2: the training criterion (det_loss.cu), the EMA update, the
# SobelSpatialAttention gate.  Each target program runs plain first.
mkdir -p gpurun_out
python tools/one_loss.py 64 2 8 > gpurun_out/plain_loss.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"loss_|tal_" -s 18 -c 6 --csv --log-file gpurun_out/r02_loss_kernels.csv python tools/one_loss.py 64 2 8 > gpurun_out/ncu_loss.log 2>&1
echo "ncu loss exit $?"
cat > /tmp/one_gate.py <<'PY'
import sys, torch
sys.path.insert(0, "spectrogram-yolov11_b200")
from specyolo import ops
from specyolo.nn.modules import SobelSpatialAttention
m = SobelSpatialAttention(7)
for (c, h) in ((128, 80), (256, 40), (512, 20)):
    x = ops.new_act(64, c, h, h, "cuda").normal_()
    for _ in range(5): m(x)
torch.cuda.synchronize()
print("ok")
PY
python /tmp/one_gate.py > gpurun_out/plain_gate.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"chan_meanmax|sobel_gate" -s 24 -c 6 --csv --log-file gpurun_out/r02_gate_kernels.csv python /tmp/one_gate.py > gpurun_out/ncu_gate.log 2>&1
echo "ncu gate exit $?"
