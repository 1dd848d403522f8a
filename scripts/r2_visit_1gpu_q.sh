#!/bin/bash
# Round 2, 1-GPU visit q: statically scheduled ILDU(0) sweeps with prefetched trip descriptors -- parity,
# apply time at 1024^2 with 1024 / 512 / 256 threads, 2048^2.
TAG=${1:-r2q}; OUT=gpurun_out/$TAG; mkdir -p $OUT
S=$OUT/summary.txt
date | tee -a $S
timeout 900 python -m pytest tests/test_gpu_ldu.py -x -q -m gpu > $OUT/pytest.log 2>&1; echo "rc=$?" | tee -a $S
tail -3 $OUT/pytest.log | tee -a $S
for th in 1024 512 256; do
  SIGB_LDU_SWEEP_THREADS=$th timeout 300 python bench.py --rows ldu > $OUT/ldu_$th.jsonl 2> $OUT/ldu_$th.err; echo "rc=$?" | tee -a $S
  grep -E "ldu apply|CG iterations" $OUT/ldu_$th.jsonl | cut -c1-330 | tee -a $S
done
for th in 1024 512; do
SIGB_LDU_SWEEP_THREADS=$th timeout 300 python bench.py --rows ldu --lgrid 2048 > $OUT/ldu2048_$th.jsonl 2> $OUT/ldu2048_$th.err; echo "rc=$?" | tee -a $S
grep -E "ldu apply|CG iterations" $OUT/ldu2048_$th.jsonl | cut -c1-330 | tee -a $S
done
date | tee -a $S
