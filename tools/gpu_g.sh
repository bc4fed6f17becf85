#!/bin/bash
for g in 1 2 3; do
  for shp in 96x68x120 192x34x60 384x17x30; do
    CAL_CONV_G=$g CAL_DEBUG_CONFIG=1 python tools/gpu_ablate2.py $shp 0 2>&1 | grep -E "res=no|halo conv" | sort -u | sed "s/^/G=$g /" | cut -c1-260
  done
done
