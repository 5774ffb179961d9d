#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider tests/test_loss_oracle.py tests/test_shim.py -x > gpurun_out/t_loss.log 2>&1; echo "loss+shim tests exit $?"; tail -12 gpurun_out/t_loss.log
