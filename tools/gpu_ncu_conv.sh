#!/bin/bash
# usage: gpu_ncu_conv.sh <name> <one_conv args...>
name=$1; shift
mkdir -p gpurun_out
python tools/one_conv.py "$@" 5 > gpurun_out/p.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"conv_halo|conv_igemm" -s 3 -c 1 -o gpurun_out/prof_$name -f python tools/one_conv.py "$@" 5 > gpurun_out/n_$name.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/n_$name.log
