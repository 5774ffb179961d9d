#!/bin/bash
mkdir -p gpurun_out
python tools/profile_layers.py 64 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"stem_conv|dwconv3x3|fusion_|psa_attention" -c 9 -o gpurun_out/prof_misc python tools/profile_layers.py 64 > gpurun_out/ncu2.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu2.log
