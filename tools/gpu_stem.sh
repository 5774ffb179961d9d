#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest -q -m gpu -p no:cacheprovider -x tests/test_gpu_kernels.py -k "stem" > gpurun_out/stem.log 2>&1
echo "stem tests exit $?"; grep -E "^(FAILED|ERROR)|passed|failed|Error|error|assert" gpurun_out/stem.log | head -20
timeout 120 python tools/one_stem.py 64 640 640
