#!/bin/bash
# Round 2, 1-GPU visit s: ILDU(0) statically scheduled sweeps -- what bounds a trip: plain stores in front of
# the barrier vs bulk stores through shared memory, L2 prefetch ahead of the staging (A/B), parity.
TAG=${1:-r2s}; OUT=gpurun_out/$TAG; mkdir -p $OUT
S=$OUT/summary.txt
date | tee -a $S
timeout 900 python -m pytest tests/test_gpu_ldu.py -x -q -m gpu > $OUT/pytest.log 2>&1; echo "rc=$?" | tee -a $S
tail -3 $OUT/pytest.log | tee -a $S
i=0
for cfg in "SIGB_LDU_SWEEP_TMA_STORE=1 SIGB_LDU_SWEEP_L2_AHEAD=8" "SIGB_LDU_SWEEP_TMA_STORE=0 SIGB_LDU_SWEEP_L2_AHEAD=8" \
           "SIGB_LDU_SWEEP_TMA_STORE=1 SIGB_LDU_SWEEP_L2_AHEAD=0" "SIGB_LDU_SWEEP_TMA_STORE=0 SIGB_LDU_SWEEP_L2_AHEAD=0" \
           "SIGB_LDU_SWEEP_TMA_STORE=1 SIGB_LDU_SWEEP_L2_AHEAD=24" "SIGB_LDU_SWEEP_TMA_STORE=1 SIGB_LDU_SWEEP_L2_AHEAD=8 SIGB_LDU_SWEEP_THREADS=512"; do
  i=$((i+1))
  env $cfg timeout 300 python bench.py --rows ldu > $OUT/ldu_$i.jsonl 2> $OUT/ldu_$i.err; echo "rc=$? $cfg" | tee -a $S
  grep -E "ldu apply|CG iterations" $OUT/ldu_$i.jsonl | cut -c1-100 | tee -a $S
done
timeout 300 python bench.py --rows ldu --lgrid 2048 > $OUT/ldu2048.jsonl 2> $OUT/ldu2048.err; echo "rc=$?" | tee -a $S
grep -E "ldu apply|CG iterations" $OUT/ldu2048.jsonl | cut -c1-330 | tee -a $S
date | tee -a $S
