#!/bin/bash
mkdir -p gpurun_out
for sh in "64 64 3 1 40 40 64" "128 128 1 1 40 40 64" "32 16 3 1 160 160 64" "128 128 3 1 20 20 64"; do python tools/one_conv.py $sh 50; done
python tools/one_conv.py 64 64 3 1 40 40 64 5 > gpurun_out/p.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 3 -c 1 -o gpurun_out/prof_c3_64_40 python tools/one_conv.py 64 64 3 1 40 40 64 5 > gpurun_out/n1.log 2>&1
python tools/one_conv.py 32 16 3 1 160 160 64 5 > gpurun_out/p.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 3 -c 1 -o gpurun_out/prof_c3_32_160 python tools/one_conv.py 32 16 3 1 160 160 64 5 > gpurun_out/n2.log 2>&1
echo done
