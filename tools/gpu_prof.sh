#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/profile_layers.py 64 > gpurun_out/layers_b64.txt 2>&1; echo "prof exit $?"
cat gpurun_out/layers_b64.txt
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_power_cap --format=csv > gpurun_out/smi_query.txt 2>&1; cat gpurun_out/smi_query.txt
