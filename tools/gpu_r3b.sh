#!/bin/bash
# round-2 visit B: new end-to-end tests, filter-row grouped 3x3 kernel (DX) A/B, overlapped camera solve A/B
TAG=r3b
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q -s > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest_gpu.log
grep -E "end-to-end|camera from CUDA|passed|failed|Error|error" $OUT/${TAG}_pytest_gpu.log | tail -12
echo "--- shapes DX on"; timeout 300 python tools/ncu_shapes.py 64 2>&1 | tee $OUT/${TAG}_shapes_dx1.txt
echo "--- shapes DX off"; CAL_CONV_DX=0 timeout 300 python tools/ncu_shapes.py 64 2>&1 | tee $OUT/${TAG}_shapes_dx0.txt
echo "--- shapes DX on, headroom 40960"; CAL_SMEM_HEADROOM=40960 timeout 300 python tools/ncu_shapes.py 64 2>&1 | tee $OUT/${TAG}_shapes_dx1_hr.txt
echo "--- bench full (overlap)"; timeout 600 python bench.py --workload full --steps 10 --warmup 3 --shapes-out $OUT/${TAG}_shapes_full.csv > $OUT/${TAG}_bench_full.json 2> $OUT/${TAG}_bench_full.err; echo "rc=$?"
python -c "import json;d=json.load(open('$OUT/${TAG}_bench_full.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac'],d['kernels_ms_per_step'])"
echo "--- bench full (no overlap)"; CAL_SOLVE_OVERLAP=0 timeout 600 python bench.py --workload full --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_full_nooverlap.json 2>> $OUT/${TAG}_bench_full.err
python -c "import json;d=json.load(open('$OUT/${TAG}_bench_full_nooverlap.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac'])"
echo "--- bench kp_decode headroom 0 / 40960"
CAL_SMEM_HEADROOM=0 timeout 600 python bench.py --workload kp_decode --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_kp_hr0.json 2>> $OUT/${TAG}_bench_full.err
python -c "import json;d=json.load(open('$OUT/${TAG}_bench_kp_hr0.json'));print(d['value'],d['ms_per_step'],d['roofline']['frac'])"
CAL_SMEM_HEADROOM=40960 timeout 600 python bench.py --workload kp_decode --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_kp_hr40.json 2>> $OUT/${TAG}_bench_full.err
python -c "import json;d=json.load(open('$OUT/${TAG}_bench_kp_hr40.json'));print(d['value'],d['ms_per_step'],d['roofline']['frac'])"
timeout 600 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "cap/" -f -o $OUT/${TAG}_shapes \
    python tools/ncu_shapes.py 64 2 > $OUT/${TAG}_ncu_shapes.txt 2>&1; tail -2 $OUT/${TAG}_ncu_shapes.txt
tail -5 $OUT/${TAG}_bench_full.err
