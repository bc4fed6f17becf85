#!/bin/bash
TAG=r3c
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_hrnet_gpu.py -m gpu -x -q -s > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest_gpu.log
grep -E "end-to-end|camera from CUDA|passed|failed|Error|error|assert" $OUT/${TAG}_pytest_gpu.log | tail -12
echo "--- shapes DX on"; timeout 300 python tools/ncu_shapes.py 64 2>&1 | tee $OUT/${TAG}_shapes_dx1.txt
echo "--- shapes DX off"; CAL_CONV_DX=0 timeout 300 python tools/ncu_shapes.py 64 2 2>&1 | tee $OUT/${TAG}_shapes_dx0.txt
echo "--- shapes DX on, headroom 40960"; CAL_SMEM_HEADROOM=40960 timeout 300 python tools/ncu_shapes.py 64 2>&1 | tee $OUT/${TAG}_shapes_dx1_hr.txt
for hr in 40960 0; do
echo "--- bench full overlap headroom $hr"; CAL_SMEM_HEADROOM=$hr timeout 600 python bench.py --workload full --steps 10 --warmup 3 --no-cpu-baseline --shapes-out $OUT/${TAG}_shapes_full_hr$hr.csv > $OUT/${TAG}_bench_full_hr$hr.json 2> $OUT/${TAG}_bench_full.err; echo "rc=$?"
python -c "import json;d=json.load(open('$OUT/${TAG}_bench_full_hr$hr.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac'],d['kernels_ms_per_step'])"
done
timeout 600 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "cap/" -f -o $OUT/${TAG}_shapes \
    python tools/ncu_shapes.py 64 2 > $OUT/${TAG}_ncu_shapes.txt 2>&1; tail -2 $OUT/${TAG}_ncu_shapes.txt
tail -5 $OUT/${TAG}_bench_full.err
