#!/bin/bash
# *_OMN sibling config: MSCSpatialAttention kernels vs the oracle, the whole model vs the real reference's fixture
mkdir -p gpurun_out
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_kernels.py -k "msc_spatial or sobel" -x > gpurun_out/t_msc.log 2>&1; echo "msc tests exit $?"; tail -15 gpurun_out/t_msc.log
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_model.py -k "golden" -x > gpurun_out/t_golden.log 2>&1; echo "golden exit $?"; tail -15 gpurun_out/t_golden.log
