#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list, ncu --set full captures.
# usage (under gpurun): bash tools/gpu_round.sh [tag] [workload]
TAG=${1:-r1}
WL=${2:-full}
NCU=${3:-1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
nproc > $OUT/${TAG}_nproc.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest_gpu.log
tail -3 $OUT/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/${TAG}_smoke.log
timeout 900 python bench.py --workload $WL --steps 10 --warmup 3 --shapes-out $OUT/${TAG}_shapes_${WL}.csv > $OUT/${TAG}_bench_${WL}.json 2> $OUT/${TAG}_bench_${WL}.err; echo "bench rc=$?"
cat $OUT/${TAG}_bench_${WL}.json
timeout 600 python bench.py --workload $WL --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_${WL}_ref.json 2>> $OUT/${TAG}_bench_${WL}.err
cat $OUT/${TAG}_bench_${WL}_ref.json
if [ "$NCU" = "1" ]; then
# launch list of the bench command (shares of the step; numbers printed under ncu are not bench values)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/${TAG}_launches_${WL}.csv \
    python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_under_ncu.txt 2>&1
# full captures (one pass of the whole pipeline each; -s skips the warm-up pass): 3x3 halo convs of
# stage 3/4, generic convs, the fused head, the fuse kernel, the decodes and the camera solve
for spec in "halo:conv3x3_halo:420:6" "convtc:conv_tc_kernel:360:8" "head:head_:2:2" "combine:fuse_combine:40:3" \
            "decode:kp_decode:1:1" "linedecode:line_decode:1:1" "solve:camera_solve:1:1" "stem:stem_conv:2:1"; do
  IFS=: read name regex skip count <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c $count -f -o $OUT/${TAG}_$name \
      python tools/ncu_pipeline.py 64 > $OUT/${TAG}_ncu_$name.txt 2>&1
done
fi
ls -la $OUT | tail -20
