#!/bin/bash
# kernels whose dynamic shared-memory attribute goes through ensure_dynamic_smem (common.h) + the *_GC block timings
mkdir -p gpurun_out
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_kernels.py tests/test_loss_oracle.py tests/test_metrics.py -k "stft or bottlenect or loss or match or metric" -x > gpurun_out/t_smem.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/t_smem.log
timeout 300 python tools/one_gc.py 64 32 160 10 > gpurun_out/plain_gc.log 2>&1 && cat gpurun_out/plain_gc.log && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"gc_" -s 16 -c 4 --csv --log-file gpurun_out/r02_gc_kernels.csv python tools/one_gc.py 64 32 160 3 > gpurun_out/ncu_gc.log 2>&1
echo "ncu gc exit $?"; grep -E "gpu__time_duration" gpurun_out/r02_gc_kernels.csv | awk -F'","' '{print $5, $NF}'
