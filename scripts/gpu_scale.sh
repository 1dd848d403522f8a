#!/bin/bash
# Strong-scaling run of bench.py at N = 1,2,4,8 (what the driver does at round end).  bash scripts/gpu_scale.sh <tag> [steps]
TAG=${1:-scale}; STEPS=${2:-200}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps $STEPS --warmup 5 > $OUT/bench_n$n.json 2> $OUT/bench_n$n.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $n --steps $STEPS --warmup 5 > $OUT/bench_n$n.json 2> $OUT/bench_n$n.err
  fi
  echo "N=$n rc=$?"; cat $OUT/bench_n$n.json; tail -2 $OUT/bench_n$n.err | grep -v "^\*\|OMP"
done
