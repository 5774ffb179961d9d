#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_model.py > gpurun_out/model.log 2>&1
echo "model exit $?"; grep -E "^(FAILED|ERROR)|passed|failed|Error" gpurun_out/model.log | head -30
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python -c "
import json; d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','launches_per_step')}, d['e2e']['value'], d['clocks'])
print({k: d['roofline'][k] for k in ('achieved','frac','ms_per_step','hbm_gbs_same_launches')})
for k,v in d['stages'].items(): print(k, {a: (round(b,3) if isinstance(b,float) else b) for a,b in v.items()})
"; tail -3 gpurun_out/bench.err
