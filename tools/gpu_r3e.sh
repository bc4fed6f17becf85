#!/bin/bash
TAG=r3e
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -s > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest_gpu.log
grep -E "passed|failed|Error|^E |FAILED|EvalAI" $OUT/${TAG}_pytest_gpu.log | tail -30
for st in 4 5 6; do echo "--- A stages (resident) $st"; CAL_A_STAGES_RES=$st timeout 300 python tools/ncu_shapes.py 64 2 2>&1 | tee -a $OUT/${TAG}_astages.txt; done
