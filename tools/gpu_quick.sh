#!/bin/bash
# quick visit: op parity + whole-net parity + one bench line with per-shape times
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_hrnet_gpu.py tests/test_engine_gpu.py -m gpu -q -x 2>&1 | tail -4
timeout 900 python bench.py --workload full --steps 10 --warmup 3 --no-cpu-baseline --shapes-out $OUT/${TAG}_shapes_full.csv > $OUT/${TAG}_bench_full.json 2> $OUT/${TAG}_bench_full.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('$OUT/${TAG}_bench_full.json'));print('full', d['value'],d['ms_per_step'],d['ms_per_step_steady_state'],d['e2e']['value'],d['roofline']['frac'],d['kernels_ms_per_step'])"
head -30 $OUT/${TAG}_shapes_full.csv
