#!/bin/bash
# round-2 visit A: parity on the device (new solvePnPRansac restatement), solve timing, counter-backed captures
TAG=r3a
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest_gpu.log
tail -15 $OUT/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -20 $OUT/${TAG}_smoke.log
timeout 600 python bench.py --workload full --steps 10 --warmup 3 --shapes-out $OUT/${TAG}_shapes_full.csv > $OUT/${TAG}_bench_full.json 2> $OUT/${TAG}_bench_full.err; echo "bench rc=$?"
cat $OUT/${TAG}_bench_full.json
head -12 $OUT/${TAG}_shapes_full.csv
timeout 300 python tools/ncu_shapes.py 64 > $OUT/${TAG}_shapes_alone.txt 2>&1; cat $OUT/${TAG}_shapes_alone.txt
timeout 900 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "cap/" -f -o $OUT/${TAG}_shapes \
    python tools/ncu_shapes.py 64 > $OUT/${TAG}_ncu_shapes.txt 2>&1; tail -3 $OUT/${TAG}_ncu_shapes.txt
for spec in "head:head_:2:1" "convtc256:conv_tc_kernel:360:8" "solve:camera_solve:1:1"; do
  IFS=: read name regex skip count <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c $count -f -o $OUT/${TAG}_$name \
      python tools/ncu_pipeline.py 64 > $OUT/${TAG}_ncu_$name.txt 2>&1
done
ls -la $OUT | tail -12
