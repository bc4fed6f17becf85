#!/bin/bash
TAG=r3d
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q -s > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest_gpu.log
grep -E "passed|failed|Error|error|assert" $OUT/${TAG}_pytest_gpu.log | tail -12
for ab in 0 2 8 10 4 6; do
echo "--- ablate $ab (1: no TMA stores, 2: no epilogue math/stores, 4: no MMA, 8: no A loads)"; CAL_DEBUG_ABLATE=$ab timeout 300 python tools/ncu_shapes.py 64 2 2>&1 | tee -a $OUT/${TAG}_ablate.txt
done
echo "--- DX off ablate 10"; CAL_CONV_DX=0 CAL_DEBUG_ABLATE=10 timeout 300 python tools/ncu_shapes.py 64 2 2>&1 | tee -a $OUT/${TAG}_ablate.txt
echo "--- mma rate probe"; timeout 300 python tools/gpu_mma_rate.py 2>&1 | tail -30 | tee $OUT/${TAG}_mma_rate.txt
