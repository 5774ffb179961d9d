#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/bench_c5_n1.json 2> gpurun_out/bench_c5_n1.err; echo "bench n1 exit $?"
python - <<'PY'
import json; d=json.loads(open('gpurun_out/bench_c5_n1.json').read().strip().splitlines()[-1]); print(d['value'], d['c5'])
PY
tail -3 gpurun_out/bench_c5_n1.err
