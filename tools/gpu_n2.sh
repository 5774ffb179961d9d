#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 50 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 exit $?"
python - <<'PY'
import json; d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['c3']['value'], d['c4']['value']); print(d['c5'])
PY
tail -5 gpurun_out/bench_n2.err
