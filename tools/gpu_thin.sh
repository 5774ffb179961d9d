#!/bin/bash
# A/B of the paired-task thin epilogue: kernel tests, then the thin layers with and without it
mkdir -p gpurun_out
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_kernels.py -k "conv or stem" > gpurun_out/conv.log 2>&1
echo "conv tests exit $?"; grep -E "^(FAILED|ERROR)|passed|failed|Error|error" gpurun_out/conv.log | head -30
for sh in "32 16 3 1 160 160 64" "16 32 3 1 160 160 64" "16 32 3 1 160 160 64 20 1 1 1" "64 32 3 1 80 80 64" "64 32 3 1 40 40 64" "32 32 3 1 20 20 64" "32 32 3 1 20 20 64 20 1 1 1"; do
  timeout 120 python tools/one_conv.py $sh; SPECYOLO_NO_THIN=1 timeout 120 python tools/one_conv.py $sh | sed 's/^/   old: /'
done
timeout 120 python tools/one_stem.py 64 640 640; SPECYOLO_NO_THIN=1 timeout 120 python tools/one_stem.py 64 640 640 | sed 's/^/   old: /'
