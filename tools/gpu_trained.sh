#!/bin/bash
# Trains the fixture checkpoint with the unmodified reference on the GPU (oracle/train_fixture.py), then runs the
# trained-weights parity test.  The checkpoint comes back under gpurun_out/_trained/ (copy it to tests/golden/_trained/).
mkdir -p gpurun_out/_trained tests/golden/_trained
if [ ! -f tests/golden/_trained/spec_s.pt ]; then
  timeout 1500 python oracle/train_fixture.py --out tests/golden/_trained/spec_s.pt --device 0 --epochs ${EPOCHS:-60} --images ${IMAGES:-320} > gpurun_out/train_fixture.log 2>&1
  echo "train exit $?"; tail -5 gpurun_out/train_fixture.log
  cp tests/golden/_trained/spec_s.pt gpurun_out/_trained/
fi
timeout 600 python -m pytest -q -m gpu -p no:cacheprovider tests/test_parity_trained.py -x > gpurun_out/parity_trained.log 2>&1
echo "test exit $?"; tail -30 gpurun_out/parity_trained.log
cat gpurun_out/parity_trained_b64_640_vs_real_reference.json
