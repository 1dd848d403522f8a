#!/bin/bash
# Size sweep on one GPU: fixed cost vs streaming cost of each kernel.  bash scripts/gpu_sweep.sh <tag>
TAG=${1:-sweep}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for g in 4096 2896 2048 1448 1024 724 512 256; do
  timeout 300 python bench.py --grid $g --steps 200 --warmup 5 --quick 2>> $OUT/err.log | sed "s/^{/{\"grid\": $g, /" | tee -a $OUT/sweep.jsonl
done
