#!/bin/bash
# Runs the GPU kernel test-suite in separate processes (a trapped kernel poisons its CUDA context),
# each under its own timeout; logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
ls /root/reference > gpurun_out/ref_ls.txt 2>&1
run() { # name, pytest args...
  local name=$1; shift
  timeout 300 python -m pytest -q -m gpu -p no:cacheprovider "$@" > gpurun_out/$name.log 2>&1
  echo "$name exit $?" | tee -a gpurun_out/summary.txt
  tail -n 5 gpurun_out/$name.log
}
: > gpurun_out/summary.txt
run layout   tests/test_gpu_kernels.py -k "layout"
run conv1x1  tests/test_gpu_kernels.py -k "test_conv_igemm and (1-1-1-1 or 1-1-1-1-1)"
run conv     tests/test_gpu_kernels.py -k "conv"
run misc     tests/test_gpu_kernels.py -k "stem or depthwise or sppf or fusion or attention"
run decode   tests/test_gpu_kernels.py -k "decode"
run nms      tests/test_gpu_kernels.py -k "nms"
run stft     tests/test_gpu_kernels.py -k "stft"
cat gpurun_out/summary.txt
