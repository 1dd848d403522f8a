#!/bin/bash
# Round 2, 1-GPU visit o: statically scheduled ILDU(0) sweeps (parity, apply time at 1024^2 and 2048^2,
# the chunked form beside it), small tile shape with batched row sums on the ER operators.
TAG=${1:-r2o}; OUT=gpurun_out/$TAG; mkdir -p $OUT
S=$OUT/summary.txt
date | tee -a $S
echo "== parity: ldu, spmv, solvers" | tee -a $S
timeout 900 python -m pytest tests/test_gpu_ldu.py tests/test_gpu_spmv.py tests/test_gpu_solvers.py -x -q -m gpu > $OUT/pytest.log 2>&1; echo "rc=$?" | tee -a $S
tail -5 $OUT/pytest.log | tee -a $S
echo "== ILDU" | tee -a $S
for cfg in "SIGB_LDU_STATIC=1" "SIGB_LDU_STATIC=0"; do
  env $cfg timeout 300 python bench.py --rows ldu > $OUT/ldu_$cfg.jsonl 2> $OUT/ldu_$cfg.err; echo "rc=$?" | tee -a $S
  cut -c1-420 $OUT/ldu_$cfg.jsonl | tee -a $S
done
timeout 300 python bench.py --rows ldu --lgrid 2048 > $OUT/ldu2048.jsonl 2> $OUT/ldu2048.err; echo "rc=$?" | tee -a $S
cut -c1-420 $OUT/ldu2048.jsonl | tee -a $S
echo "== ER operators, small tile shape with batched row sums" | tee -a $S
run() { env $1 timeout 400 python scripts/spmv_probe.py $2 2>> $OUT/er.err | tee -a $OUT/er.jsonl | cut -c1-400 | tee -a $S; }
run "SIGB_X=0" "--kind er --n 2000000"
run "SIGB_SMALL_TILE_CTAS=5" "--kind er --n 2000000"
run "SIGB_X=0" "--kind er --n 2000000 --dot"
run "SIGB_X=0" "--kind surrogate --n 20000000 --reps 10"
run "SIGB_X=0" "--kind poisson --n 16777216 --dot"
date | tee -a $S
