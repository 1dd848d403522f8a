#!/bin/bash
TAG=${1:-l2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for l2 in 1 0; do for pers in 1 0; do for g in 2048 1448; do
  SIGB_L2_PERSIST=$l2 SIGB_CG_PERSISTENT=$pers timeout 300 python bench.py --grid $g --steps 200 --warmup 5 --quick 2>> $OUT/err.log | sed "s/^{/{\"grid\": $g, \"persistent\": $pers, \"l2\": $l2, /" | tee -a $OUT/l2.jsonl
done; done; done
tail -3 $OUT/err.log
