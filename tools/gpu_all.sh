#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_kernels.py > gpurun_out/kernels.log 2>&1
echo "kernels exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/kernels.log | head -30
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_model.py > gpurun_out/model.log 2>&1
echo "model exit $?"; grep -E "^(FAILED|ERROR)|passed|failed|Error" gpurun_out/model.log | head -30
timeout 600 python tools/profile_layers.py 64 > gpurun_out/layers_b64_cur.txt 2>&1; head -1 gpurun_out/layers_b64_cur.txt
sort -k2 -n -r gpurun_out/layers_b64_cur.txt | head -14
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python -c "
import json; d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','launches_per_step')}, d['e2e'], d['clocks'], d['cpu_baseline'])
print(d['roofline'])
for k,v in d['stages'].items(): print(k, v)
"; tail -3 gpurun_out/bench.err
