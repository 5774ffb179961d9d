#!/bin/bash
# quick check: conv/stem kernel tests + end-to-end model parity + product bench line (no cpu baseline)
mkdir -p gpurun_out
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_kernels.py -k "conv or stem" > gpurun_out/conv.log 2>&1
echo "conv tests exit $?"; grep -E "^(FAILED|ERROR)|passed|failed|Error|error" gpurun_out/conv.log | head -30
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_model.py > gpurun_out/model.log 2>&1
echo "model tests exit $?"; grep -E "^(FAILED|ERROR)|passed|failed|Error|error" gpurun_out/model.log | head -30
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; echo "bench exit $?"
python - <<'PY'
import json; d=json.loads(open('gpurun_out/bench_q.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','launches_per_step')}, d['e2e'], d['clocks'])
PY
