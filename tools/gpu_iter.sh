#!/bin/bash
# quick iteration: conv parity, then per-layer profile
mkdir -p gpurun_out
timeout 600 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_kernels.py -k "conv" > gpurun_out/conv.log 2>&1
echo "conv exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/conv.log | head -30
timeout 600 python tools/profile_layers.py 64 > gpurun_out/layers_b64_v3.txt 2>&1; head -1 gpurun_out/layers_b64_v3.txt
for sh in "96 128 1 1 160 160 64" "64 64 1 1 160 160 64" "128 128 3 2 160 160 64" "128 128 3 1 40 40 64" "256 256 3 2 80 80 64"; do timeout 120 python tools/one_conv.py $sh; done
