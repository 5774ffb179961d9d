#!/bin/bash
for d in 4 8 12 16 24 64; do echo "D=$d"; SPECYOLO_FUSION_D=$d timeout 60 python tools/one_fusion.py 64 80 80 128 3 1; SPECYOLO_FUSION_D=$d timeout 60 python tools/one_fusion.py 64 40 40 128 3 1;  SPECYOLO_FUSION_D=$d timeout 60 python tools/one_fusion.py 64 20 20 128 2 0; done
