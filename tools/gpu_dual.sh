#!/bin/bash
# A/B of the halo kernel on the dominant shapes (profiling experiment)
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -x 2>&1 | tail -3
for shp in 48x135x240 64x135x240 96x68x120 192x34x60 384x17x30; do
  python tools/gpu_ablate2.py $shp 0
done
LO=20 python tools/gpu_timeline.py 48 135 240 res | tail -26
