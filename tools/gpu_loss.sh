#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider tests/test_loss_oracle.py -x > gpurun_out/t_loss.log 2>&1; echo "loss tests exit $?"; tail -25 gpurun_out/t_loss.log
timeout 600 python tools/one_loss.py 64 2 8 2>&1 | tail -3
timeout 600 python tools/one_loss.py 64 80 12 2>&1 | tail -3
