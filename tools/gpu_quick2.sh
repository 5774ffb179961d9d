#!/bin/bash
# kernel-iteration pass: model-level tests, in-graph layer table, short bench line without the extra workloads
mkdir -p gpurun_out
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_model.py tests/test_parity_baseline.py -x > gpurun_out/t_model.log 2>&1; echo "model tests exit $?"; tail -4 gpurun_out/t_model.log
timeout 600 python tools/profile_layers.py 64 > gpurun_out/layers_cur.txt 2>&1; echo "layers exit $?"; head -4 gpurun_out/layers_cur.txt | tail -2
timeout 900 python bench.py --no-extra --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench exit $?"
python - <<'PY'
import json; d=json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','launches_per_step')}, d['e2e']['value'], d['clocks'])
r=d['roofline']; print({k:r[k] for k in ('achieved','frac','group_ms_per_step','launches_per_step','frac_tensor_same_launches')}, r['timing']['serial_graph_ms_per_step'])
for k,v in d['stages'].items(): print(k, {kk:(round(vv,3) if isinstance(vv,float) else vv) for kk,vv in v.items()})
PY
