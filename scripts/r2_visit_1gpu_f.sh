#!/bin/bash
# Round 2, 1-GPU visit f: A/B of the row-sum batching, the staging of the dot operand and the
# inlining of the SpMV pass into the persistent kernel (build variants, never bench values).
TAG=${1:-r2f}; OUT=gpurun_out/$TAG; mkdir -p $OUT
S=$OUT/summary.txt
date | tee -a $S
for v in "" _u3b0 _u3b1 _u2b0 _u3b0in _u2b1in; do
for g in 4096 1448; do
  SIGB_LIB_VARIANT=$v timeout 300 python bench.py --grid $g --steps 200 --warmup 5 --quick --no-parity 2>> $OUT/ab.err | sed "s/^{/{\"grid\": $g, /" | tee -a $OUT/ab.jsonl | tee -a $S
done; done
for v in "" _u3b0 _u2b0; do
  SIGB_LIB_VARIANT=$v timeout 300 python scripts/spmv_probe.py --kind er --n 2000000 2>> $OUT/er.err | tee -a $OUT/er.jsonl | tee -a $S
done
date | tee -a $S
