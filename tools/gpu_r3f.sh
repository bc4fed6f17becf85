#!/bin/bash
TAG=r3f
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -x -s > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest_gpu.log
grep -E "passed|failed|Error|^E |FAILED" $OUT/${TAG}_pytest_gpu.log | tail -20
echo "--- bench full engine"; timeout 600 python bench.py --workload full --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_full.json 2> $OUT/${TAG}_bench_full.err; echo "rc=$?"
python -c "import json;d=json.load(open('$OUT/${TAG}_bench_full.json'));print(d['value'],d['ms_per_step'],d['e2e'],d['roofline']['frac'],d['gpu_launches'])"
echo "--- bench full python schedule"; CAL_ENGINE=0 timeout 600 python bench.py --workload full --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_full_py.json 2>> $OUT/${TAG}_bench_full.err; echo "rc=$?"
python -c "import json;d=json.load(open('$OUT/${TAG}_bench_full_py.json'));print(d['value'],d['ms_per_step'],d['e2e'],d['roofline']['frac'],d['gpu_launches'])"
tail -5 $OUT/${TAG}_bench_full.err
