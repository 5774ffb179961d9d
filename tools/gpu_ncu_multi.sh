#!/bin/bash
# usage: gpu_ncu_multi.sh <name> <kernel regex> <skip> <count> <python script + args...>
name=$1; regex=$2; skip=$3; cnt=$4; shift 4
mkdir -p gpurun_out
python "$@" > gpurun_out/p.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"$regex" -s $skip -c $cnt -o gpurun_out/prof_$name -f python "$@" > gpurun_out/n_$name.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/n_$name.log
