#!/bin/bash
# Size sweep on one GPU: fixed cost vs streaming cost.  bash scripts/gpu_sweep.sh <tag> [grids...]
TAG=${1:-sweep}; shift; OUT=gpurun_out/$TAG; mkdir -p $OUT
GRIDS=${@:-"4096 2896 2048 1448 1024 724 512 256"}
for pers in 1 0; do
for g in $GRIDS; do
  SIGB_CG_PERSISTENT=$pers timeout 300 python bench.py --grid $g --steps 200 --warmup 5 --quick 2>> $OUT/err.log | sed "s/^{/{\"grid\": $g, \"persistent\": $pers, /" | tee -a $OUT/sweep.jsonl
done
done
