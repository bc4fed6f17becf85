#!/bin/bash
# multi-GPU visit: bash tools/gpu_multi.sh <tag> <n_gpus> [sweep 0|1]
TAG=${1:-r4b}
N=${2:-2}
SWEEP=${3:-0}
OUT=gpurun_out
mkdir -p $OUT
run() {  # name, extra args
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 10 --warmup 3 $2 > $OUT/${TAG}_bench_full_${N}gpu$1.json 2> $OUT/${TAG}_bench_${N}gpu$1.err
  echo "rc=$?"; python -c "import json;d=json.load(open('$OUT/${TAG}_bench_full_${N}gpu$1.json'));print('$N GPUs$1', d['value'],d['ms_per_step'],d['ms_per_step_steady_state'],d['e2e']['value'],d['multi_gpu_check'])" || tail -5 $OUT/${TAG}_bench_${N}gpu$1.err
}
run "" ""
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_full_${N}gpu_ref.json 2>> $OUT/${TAG}_bench_${N}gpu.err; cut -c1-200 $OUT/${TAG}_bench_full_${N}gpu_ref.json
if [ "$SWEEP" = "1" ]; then
  run "_720p" "--height 720 --width 1280 --batch 32"
  run "_1080p" "--height 1080 --width 1920 --batch 16"
fi
