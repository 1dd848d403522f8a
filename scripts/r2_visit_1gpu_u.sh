#!/bin/bash
# Round 2, 1-GPU visit u: ILDU(0) statically scheduled sweeps, one slab per trip and a straight-line body
# for short rows -- parity, apply time by CTA size (1024^2), 2048^2, phase cycles (diagnostic build).
TAG=${1:-r2u}; OUT=gpurun_out/$TAG; mkdir -p $OUT
S=$OUT/summary.txt
date | tee -a $S
timeout 900 python -m pytest tests/test_gpu_ldu.py -x -q -m gpu > $OUT/pytest.log 2>&1; echo "rc=$?" | tee -a $S
tail -3 $OUT/pytest.log | tee -a $S
for th in 992 512 256; do
  SIGB_LDU_SWEEP_THREADS=$th timeout 300 python bench.py --rows ldu > $OUT/ldu_$th.jsonl 2> $OUT/ldu_$th.err; echo "rc=$? threads $th" | tee -a $S
  grep -E "ldu setup|ldu apply|CG iterations" $OUT/ldu_$th.jsonl | cut -c1-250 | tee -a $S
done
for th in 992 512; do
  SIGB_LDU_SWEEP_THREADS=$th timeout 300 python bench.py --rows ldu --lgrid 2048 > $OUT/ldu2048_$th.jsonl 2> $OUT/ldu2048_$th.err; echo "rc=$? threads $th (2048^2)" | tee -a $S
  grep -E "ldu setup|ldu apply|CG iterations" $OUT/ldu2048_$th.jsonl | cut -c1-250 | tee -a $S
done
for th in 992 512; do
SIGB_LDU_SWEEP_THREADS=$th SIGB_LIB_VARIANT=_sweepstats timeout 200 python bench.py --rows ldu 2>&1 | grep -E "^sweep" | tail -6 | tee -a $S
done
date | tee -a $S
