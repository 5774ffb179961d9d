#!/bin/bash
# Round-2 first GPU pass: the new parity tests first (fast feedback), then the whole GPU suite, smoke, bench line.
mkdir -p gpurun_out
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider tests/test_parity_baseline.py -x > gpurun_out/parity_tests.log 2>&1
echo "parity tests exit $?"; tail -15 gpurun_out/parity_tests.log
timeout 1500 python -m pytest -q -m gpu -p no:cacheprovider tests --deselect tests/test_parity_baseline.py > gpurun_out/gpu_tests.log 2>&1
echo "gpu tests exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/gpu_tests.log | head -30
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 1200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
tail -5 gpurun_out/bench.err
python - <<'PY'
import json; d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','launches_per_step')}, d['e2e'], d['clocks'], d['cpu_baseline'])
print(d['roofline'])
for k,v in d['stages'].items(): print(k, v)
print('c3', d['c3']); print('c4', d['c4']); print('lib', d['library_bar'])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"; cat gpurun_out/bench_ref.json
timeout 600 python tools/profile_layers.py 64 > gpurun_out/layers_r2a.txt 2>&1; echo "layers exit $?"
