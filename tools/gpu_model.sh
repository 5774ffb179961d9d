#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_model.py > gpurun_out/model.log 2>&1
echo "model exit $?"
grep -E "^(FAILED|ERROR)|passed|failed|Error|rel-L2|assert " gpurun_out/model.log | head -40
timeout 900 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
tail -c 6000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
