#!/bin/bash
# Reference trainer, stock vs with specyolo's criterion + EMA bound in: median train-step time on one GPU.
mkdir -p gpurun_out
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider tests/test_shim.py tests/test_gpu_model.py tests/test_parity_trained.py 2>&1 | tail -3
for mode in "" "--shim" "" "--shim"; do
  timeout 900 python oracle/train_fixture.py --out /tmp/fix_$RANDOM.pt --device 0 --epochs 4 --images 320 --batch ${BATCH:-16} $mode > gpurun_out/train_step.log 2>&1; grep "train step" gpurun_out/train_step.log || tail -5 gpurun_out/train_step.log
done
