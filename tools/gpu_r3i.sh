#!/bin/bash
TAG=r3i
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest_gpu.log
grep -E "passed|failed|^E   |FAILED" $OUT/${TAG}_pytest_gpu.log | cut -c1-250 | tail -10
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/${TAG}_smoke.log
echo "--- bench full"; timeout 600 python bench.py --workload full --steps 10 --warmup 3 --shapes-out $OUT/${TAG}_shapes_full.csv > $OUT/${TAG}_bench_full.json 2> $OUT/${TAG}_bench_full.err; echo "rc=$?"
python -c "import json;d=json.load(open('$OUT/${TAG}_bench_full.json'));print(d['value'],d['ms_per_step'],d['e2e'],d['roofline']['frac'],d['gpu_launches'],d['kernels_ms_per_step'])"
echo "--- bench full float frames"; timeout 600 python bench.py --workload full --steps 10 --warmup 3 --no-cpu-baseline --float-frames > $OUT/${TAG}_bench_full_float.json 2>> $OUT/${TAG}_bench_full.err
python -c "import json;d=json.load(open('$OUT/${TAG}_bench_full_float.json'));print(d['value'],d['ms_per_step'],d['e2e'])"
echo "--- bench kp_decode"; timeout 600 python bench.py --workload kp_decode --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_kp_decode.json 2>> $OUT/${TAG}_bench_full.err
python -c "import json;d=json.load(open('$OUT/${TAG}_bench_kp_decode.json'));print(d['value'],d['ms_per_step'],d['e2e'],d['roofline']['frac'])"
tail -5 $OUT/${TAG}_bench_full.err
