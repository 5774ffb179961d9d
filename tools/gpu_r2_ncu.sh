#!/bin/bash
# Round-2 ncu evidence: launch list of the bench command (C2 legs only: --no-extra keeps the serialised replay short),
# then one full capture each of the kernels that changed this round.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline"
$CMD > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"; wc -l gpurun_out/r02_launches.csv
python tools/one_bneck.py 64 32 80 80 64 6 1 > gpurun_out/plain_bneck.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:bneck_pair -s 3 -c 1 -o gpurun_out/r02_prof_bneck64 -f python tools/one_bneck.py 64 32 80 80 64 6 1 > gpurun_out/ncu_bneck.log 2>&1
echo "ncu bneck exit $?"
python tools/one_nms.py > gpurun_out/plain_nms.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"psa_attention_small|nms_kernel|sppf_pool|detect_decode" -s 8 -c 4 -o gpurun_out/r02_prof_tails -f python tools/one_nms.py > gpurun_out/ncu_tails.log 2>&1
echo "ncu tails exit $?"; tail -2 gpurun_out/ncu_tails.log
