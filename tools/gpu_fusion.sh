#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_kernels.py -k "fusion" > gpurun_out/fusion.log 2>&1
echo "fusion tests exit $?"; grep -E "^(FAILED|ERROR)|passed|failed|Error|error" gpurun_out/fusion.log | head
timeout 60 python tools/one_fusion.py 64 80 80 128 3 1; timeout 60 python tools/one_fusion.py 64 40 40 128 3 1; timeout 60 python tools/one_fusion.py 64 40 40 128 2 0; timeout 60 python tools/one_fusion.py 64 20 20 128 2 0
timeout 60 python tools/one_fusion.py 64 40 40 256 2 1; timeout 60 python tools/one_fusion.py 64 80 80 64 3 1
