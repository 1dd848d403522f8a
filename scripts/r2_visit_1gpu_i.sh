#!/bin/bash
# Round 2, 1-GPU visit i: chunked ILDU sweeps with prefetch, device-resident Lanczos row, where the
# add_value stream spends its time.
TAG=${1:-r2i}; OUT=gpurun_out/$TAG; mkdir -p $OUT
S=$OUT/summary.txt
date | tee -a $S
echo "== parity: ldu, solvers, cxx programs (incl. the multi-GPU program on one GPU), mgpu with one GPU" | tee -a $S
timeout 900 python -m pytest tests/test_gpu_ldu.py tests/test_gpu_solvers.py tests/test_cxx_host.py tests/test_gpu_mgpu.py -x -q -m gpu > $OUT/pytest.log 2>&1; echo "rc=$?" | tee -a $S
tail -4 $OUT/pytest.log | tee -a $S
echo "== ILDU" | tee -a $S
SIGB_LIB_VARIANT=_sweepstats timeout 300 python bench.py --rows ldu > $OUT/ldu_stats.log 2>&1
grep "^sweep" $OUT/ldu_stats.log | tail -12 | tee -a $S
timeout 300 python bench.py --rows ldu > $OUT/ldu.jsonl 2> $OUT/ldu.err; echo "rc=$?" | tee -a $S
cut -c1-330 $OUT/ldu.jsonl | tee -a $S
timeout 300 python bench.py --rows ldu --lgrid 2048 > $OUT/ldu2048.jsonl 2> $OUT/ldu2048.err; echo "rc=$?" | tee -a $S
cut -c1-330 $OUT/ldu2048.jsonl | tee -a $S
echo "== Lanczos, device-resident" | tee -a $S
timeout 600 python bench.py --rows lanczos --lanczos-n-big 20000000 > $OUT/lanczos.jsonl 2> $OUT/lanczos.err; echo "rc=$?" | tee -a $S
cut -c1-1200 $OUT/lanczos.jsonl | tee -a $S; tail -3 $OUT/lanczos.err | tee -a $S
echo "== widened rows, add_values stage timing" | tee -a $S
SIGB_VERBOSE=1 timeout 400 python bench.py --rows widened > $OUT/widened.jsonl 2> $OUT/widened.err; echo "rc=$?" | tee -a $S
grep "copy_matrix\|add_value" $OUT/widened.jsonl | cut -c1-260 | tee -a $S
grep "add_values" $OUT/widened.err | tail -14 | tee -a $S
date | tee -a $S
