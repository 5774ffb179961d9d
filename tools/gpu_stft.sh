#!/bin/bash
# STFT front end: parity tests, kernel timing at the BASELINE configs[2] share (32 bursts of 2^20 samples), smoke
mkdir -p gpurun_out
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_kernels.py -k "stft" -x > gpurun_out/t_stft.log 2>&1; echo "stft tests exit $?"; tail -15 gpurun_out/t_stft.log
timeout 300 python tools/one_stft.py 32 20 50 2>&1 | tail -2
timeout 300 python tools/one_stft.py 32 16 50 2>&1 | tail -2
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider tests/test_parity_baseline.py -k "c3 or iq" -x > gpurun_out/t_c3.log 2>&1; echo "c3 parity exit $?"; tail -5 gpurun_out/t_c3.log
