#!/bin/bash
# Profiles for profiles/: launch list of the bench command, DRAM traffic of the conv kernels, full captures of the top kernels
mkdir -p gpurun_out
python tools/profile_layers.py 64 > gpurun_out/layers_b64_cur.txt 2>&1 && \
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"conv_|dwpw_kernel|stem_pair_kernel|bneck_pair_kernel" --csv --log-file gpurun_out/conv_traffic.csv python tools/profile_layers.py 64 > gpurun_out/ncu_traffic.log 2>&1
echo "traffic exit $?"; wc -l gpurun_out/conv_traffic.csv
python tools/one_conv.py 64 64 3 1 80 80 64 5 > gpurun_out/p.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"conv_halo" -s 3 -c 1 -o gpurun_out/prof_halo_64_64_k3_80 -f python tools/one_conv.py 64 64 3 1 80 80 64 5 > gpurun_out/n1.log 2>&1
python tools/one_conv.py 96 128 1 1 160 160 64 5 > gpurun_out/p.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"conv_igemm" -s 3 -c 1 -o gpurun_out/prof_igemm_96_128_k1_160 -f python tools/one_conv.py 96 128 1 1 160 160 64 5 > gpurun_out/n2.log 2>&1
python tools/one_conv.py 256 256 3 2 80 80 64 5 > gpurun_out/p.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"conv_igemm" -s 3 -c 1 -o gpurun_out/prof_igemm_256_256_k3s2_80 -f python tools/one_conv.py 256 256 3 2 80 80 64 5 > gpurun_out/n3.log 2>&1
python tools/one_stem_pair.py 5 > gpurun_out/p.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"stem_pair" -s 3 -c 1 -o gpurun_out/prof_stem_pair -f python tools/one_stem_pair.py 5 > gpurun_out/n5.log 2>&1
python tools/one_conv.py 16 32 3 1 160 160 64 5 1 1 1 > gpurun_out/p.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"conv_halo" -s 3 -c 1 -o gpurun_out/prof_halo_16_32_k3_160_res -f python tools/one_conv.py 16 32 3 1 160 160 64 5 1 1 1 > gpurun_out/n6.log 2>&1
python tools/one_fusion.py 64 80 80 128 3 1 5 > gpurun_out/p.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"fusion_" -s 6 -c 2 -o gpurun_out/prof_fusion -f python tools/one_fusion.py 64 80 80 128 3 1 5 > gpurun_out/n7.log 2>&1
python tools/one_attn.py 64 20 20 4 5 > gpurun_out/p.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"psa_attention" -s 3 -c 1 -o gpurun_out/prof_attention -f python tools/one_attn.py 64 20 20 4 5 > gpurun_out/n4.log 2>&1
echo "full captures done"; ls -la gpurun_out/*.ncu-rep | tail -6
