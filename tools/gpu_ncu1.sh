#!/bin/bash
mkdir -p gpurun_out
python tools/profile_layers.py 64 > gpurun_out/layers_b64_v2.txt 2>&1; head -1 gpurun_out/layers_b64_v2.txt
for sh in "96 128 1 1 160 160 64" "64 64 1 1 160 160 64" "128 128 3 2 160 160 64" "128 128 3 1 40 40 64"; do python tools/one_conv.py $sh; done
python tools/one_conv.py 96 128 1 1 160 160 64 5 > gpurun_out/plain_ncu.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 3 -c 2 -o gpurun_out/prof_conv1x1 python tools/one_conv.py 96 128 1 1 160 160 64 5 > gpurun_out/ncu1.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu1.log
