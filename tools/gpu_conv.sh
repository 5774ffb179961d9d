#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_kernels.py -k "conv" > gpurun_out/conv.log 2>&1
echo "conv exit $?"
grep -E "^(FAILED|PASSED|ERROR)|passed|failed|max err" gpurun_out/conv.log | head -60
