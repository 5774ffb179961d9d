#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_kernels.py -k "conv" > gpurun_out/conv.log 2>&1
echo "conv tests exit $?"; grep -E "^(FAILED|ERROR)|passed|failed|Error|error" gpurun_out/conv.log | head -30
for sh in "32 16 3 1 160 160 64" "16 32 3 1 160 160 64" "64 32 3 1 80 80 64" "32 64 3 1 80 80 64" "128 64 3 1 80 80 64" "64 64 3 1 80 80 64" "64 64 3 1 40 40 64" "128 64 3 1 40 40 64" "32 32 3 1 20 20 64" "64 64 3 1 20 20 64" "128 64 3 1 20 20 64" "128 128 7 2 160 160 64 20 8 2" "256 128 3 2 80 80 64 20 8 2"; do
  timeout 120 python tools/one_conv.py $sh; SPECYOLO_HALO=0 timeout 120 python tools/one_conv.py $sh | sed 's/^/   per-tap: /'
done
for sh in "96 128 1 1 160 160 64" "64 64 1 1 160 160 64" "128 128 1 1 80 80 64" "192 256 1 1 80 80 64" "256 256 1 1 40 40 64" "512 512 1 1 20 20 64" "32 64 3 2 320 320 64" "128 128 3 2 160 160 64" "256 256 3 2 80 80 64"; do
  timeout 120 python tools/one_conv.py $sh
done
