#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench lines (full / kp_decode / reference arm / resolution sweep),
# ncu launch list, ncu --set full captures of the dominant launches.
# usage (under gpurun): bash tools/gpu_round.sh [tag] [ncu 0|1] [sweep 0|1]
TAG=${1:-r4a}
NCU=${2:-1}
SWEEP=${3:-1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
nproc > $OUT/${TAG}_nproc.txt
timeout 1500 python -m pytest tests -m gpu -q -s > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest_gpu.log
grep -E "passed|failed|FAILED|end-to-end|camera from|EvalAI" $OUT/${TAG}_pytest_gpu.log | cut -c1-300 | tail -12
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/${TAG}_smoke.log
timeout 900 python bench.py --workload full --steps 10 --warmup 3 --shapes-out $OUT/${TAG}_shapes_full.csv > $OUT/${TAG}_bench_full.json 2> $OUT/${TAG}_bench_full.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('$OUT/${TAG}_bench_full.json'));print('full', d['value'],d['ms_per_step'],d['ms_per_step_steady_state'],d['e2e']['value'],d['roofline']['frac'],d['kernels_ms_per_step'])"
timeout 600 python bench.py --workload full --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_full_ref.json 2>> $OUT/${TAG}_bench_full.err
cat $OUT/${TAG}_bench_full_ref.json | cut -c1-300
timeout 600 python bench.py --workload kp_decode --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_kp_decode.json 2>> $OUT/${TAG}_bench_full.err
python -c "import json;d=json.load(open('$OUT/${TAG}_bench_kp_decode.json'));print('kp_decode', d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac'],d['decode'])"
timeout 600 python bench.py --workload full --steps 40 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_full_40steps.json 2>> $OUT/${TAG}_bench_full.err
python -c "import json;d=json.load(open('$OUT/${TAG}_bench_full_40steps.json'));print('full 40 steps', d['value'],d['ms_per_step'],d['ms_per_step_steady_state'],d['e2e']['value'])"
if [ "$SWEEP" = "1" ]; then
for spec in "720:1280:32" "1080:1920:16"; do
  IFS=: read h w b <<< "$spec"
  timeout 900 python bench.py --workload full --steps 10 --warmup 3 --no-cpu-baseline --height $h --width $w --batch $b > $OUT/${TAG}_bench_full_${h}p.json 2>> $OUT/${TAG}_bench_full.err
  python -c "import json;d=json.load(open('$OUT/${TAG}_bench_full_${h}p.json'));print('${h}p', d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac'])"
done
fi
if [ "$NCU" = "1" ]; then
# launch list of the bench command (shares of the step; numbers printed under ncu are not bench values)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/${TAG}_launches_full.csv \
    python bench.py --workload full --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_under_ncu.txt 2>&1
# the dominant conv launches, one each (NVTX range 'cap' around a single warm launch, tools/ncu_shapes.py)
timeout 900 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "cap/" -f -o $OUT/${TAG}_shapes \
    python tools/ncu_shapes.py 64 > $OUT/${TAG}_ncu_shapes.txt 2>&1
# one pass of the whole pipeline each (-s skips the warm-up pass)
for spec in "head:head_:2:2" "combine:fuse_combine:40:3" "decode:kp_decode:1:1" "linedecode:line_decode:1:1" "solve:camera_solve:1:1" "stem:stem_conv:2:1"; do
  IFS=: read name regex skip count <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c $count -f -o $OUT/${TAG}_$name \
      python tools/ncu_pipeline.py 64 > $OUT/${TAG}_ncu_$name.txt 2>&1
done
fi
tail -3 $OUT/${TAG}_bench_full.err
# gpurun brings back at most 64 MiB: condense the ncu reports here (profiles/ layout under gpurun_out/prof_<tag>) and drop them
CAL_PROF_DIR=$OUT/prof_${TAG} python tools/summarize_profiles.py ${TAG} full > $OUT/${TAG}_summarize.txt 2>&1
rm -f $OUT/${TAG}_*.ncu-rep
ls $OUT | grep ${TAG} | tr '\n' ' '
du -sh $OUT
