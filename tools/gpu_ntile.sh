#!/bin/bash
# small-level 1x1 convs: does a narrower N tile (more, smaller work items) beat the whole-group tile?
cd /root/repo
for shape in "256 128 1 1 20 20" "512 512 1 1 20 20" "512 256 1 1 20 20" "1024 512 1 1 20 20" "256 512 1 1 20 20" "768 512 1 1 20 20" "128 128 1 1 40 40" "256 256 1 1 40 40" "384 256 1 1 40 40" "128 64 1 1 40 40" "192 128 1 1 40 40" "128 128 3 2 40 40" "256 512 3 2 40 40"; do
  for cap in 0 128 64; do
    if [ $cap = 0 ]; then unset SPECYOLO_NTILE_CAP; else export SPECYOLO_NTILE_CAP=$cap; fi
    echo -n "cap=$cap  "; timeout 120 python tools/one_conv.py $shape 64 50 2>&1 | tail -1
  done
done
