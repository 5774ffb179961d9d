#!/bin/bash
# full ncu capture of both Fusion kernels at the 80x80 level (k=3, c=128, first input upsampled)
mkdir -p gpurun_out
python tools/one_fusion.py 64 80 80 128 3 1 > gpurun_out/p.log 2>&1; cat gpurun_out/p.log
python tools/one_fusion.py 64 40 40 128 3 1; python tools/one_fusion.py 64 40 40 128 2 0; python tools/one_fusion.py 64 20 20 128 2 0
ncu --set full --clock-control none --import-source on -k regex:fusion_ -s 6 -c 2 -o gpurun_out/prof_fusion -f python tools/one_fusion.py 64 80 80 128 3 1 5 > gpurun_out/n_fusion.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/n_fusion.log
