#!/bin/bash
# extra full captures for profiles/: the fused DWConv+1x1 kernel and the DDWConv k7 layer (MMA-rate floor at N = 16)
mkdir -p gpurun_out
python tools/one_dwpw.py 128 128 80 80 64 5 > gpurun_out/p.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"dwpw_kernel" -s 3 -c 1 -o gpurun_out/prof_dwpw_128_128_80 -f python tools/one_dwpw.py 128 128 80 80 64 5 > gpurun_out/n8.log 2>&1
echo "dwpw exit $?"
python tools/one_conv.py 128 128 7 2 160 160 64 5 8 2 > gpurun_out/p.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"conv_halo" -s 3 -c 1 -o gpurun_out/prof_halo_ddw_k7 -f python tools/one_conv.py 128 128 7 2 160 160 64 5 8 2 > gpurun_out/n9.log 2>&1
echo "ddw exit $?"
