#!/bin/bash
mkdir -p gpurun_out
N=${1:-4}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 50 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N exit $?"
python - <<PY
import json; d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['e2e']['value'], d['c3']['value'], d['c4']['value']); print({k: d['c5'].get(k) for k in ('value','ms_per_step','stock_value','loss_last_step','error','unavailable')})
PY
tail -3 gpurun_out/bench_n$N.err
