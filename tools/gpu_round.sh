#!/bin/bash
# Full GPU check for a milestone: parity tests, bench line, ncu launch list of the bench command.
mkdir -p gpurun_out
timeout 1200 python -m pytest -q -m gpu -p no:cacheprovider tests > gpurun_out/gpu_tests.log 2>&1
echo "gpu tests exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/gpu_tests.log | head -30
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json; d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','launches_per_step')}, d['e2e'], d['clocks'], d['cpu_baseline'])
print(d['roofline'])
for k,v in d['stages'].items(): print(k, v)
PY
tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"; cat gpurun_out/bench_ref.json
if [ "$1" = "ncu" ]; then
# (--no-extra: the c3 / c4 / c5 / library legs run after the timed region and would only add thousands of library launches)
python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu_launches.log; wc -l gpurun_out/launches.csv
fi
