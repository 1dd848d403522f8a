#!/bin/bash
# Round 2, 1-GPU visit p: statically scheduled ILDU(0) sweeps -- parity and apply time (1024^2, 2048^2).
TAG=${1:-r2p}; OUT=gpurun_out/$TAG; mkdir -p $OUT
S=$OUT/summary.txt
date | tee -a $S
echo "== parity: ldu" | tee -a $S
timeout 900 python -m pytest tests/test_gpu_ldu.py -x -q -m gpu > $OUT/pytest.log 2>&1; echo "rc=$?" | tee -a $S
tail -5 $OUT/pytest.log | tee -a $S
echo "== ILDU" | tee -a $S
timeout 300 python bench.py --rows ldu > $OUT/ldu.jsonl 2> $OUT/ldu.err; echo "rc=$?" | tee -a $S
cut -c1-420 $OUT/ldu.jsonl | tee -a $S
timeout 300 python bench.py --rows ldu --lgrid 2048 > $OUT/ldu2048.jsonl 2> $OUT/ldu2048.err; echo "rc=$?" | tee -a $S
cut -c1-420 $OUT/ldu2048.jsonl | tee -a $S
date | tee -a $S
