#!/bin/bash
# Round 2, 1-GPU visit e: the consolidated library (pipelined SpMV as the only form, 3-CTA persistent
# CG with the initial residual fused in, chunked ILDU sweeps, device tilings + stream-ordered scratch).
TAG=${1:-r2e}; OUT=gpurun_out/$TAG; mkdir -p $OUT
S=$OUT/summary.txt
want() { [[ -z "$SECTIONS" || " $SECTIONS " == *" $1 "* ]]; }
date | tee -a $S
if want 1; then
echo "== 1. pytest -m gpu" | tee -a $S
timeout 1200 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" | tee -a $S
tail -15 $OUT/pytest_gpu.log | tee -a $S
fi
if want 2; then
echo "== 2. bench: driver flags, then 200 steps; shard sizes (quick)" | tee -a $S
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_20.json 2> $OUT/bench_20.err; echo "rc=$?" | tee -a $S
cut -c1-2500 $OUT/bench_20.json | tee -a $S
timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu > $OUT/bench_200.json 2> $OUT/bench_200.err; echo "rc=$?" | tee -a $S
cut -c1-400 $OUT/bench_200.json | tee -a $S
for g in 2048 1448; do for st in 200 20; do
  timeout 300 python bench.py --grid $g --steps $st --warmup 5 --quick 2>> $OUT/quick.err | sed "s/^{/{\"grid\": $g, \"steps\": $st, /" | tee -a $OUT/quick.jsonl | tee -a $S
done; done
SIGB_LIB_VARIANT=_timers timeout 300 python bench.py --grid 1448 --steps 400 --warmup 5 --quick > /dev/null 2> $OUT/phases.err
grep "phase_us\|spmv_cta" $OUT/phases.err | tee -a $S
SIGB_LIB_VARIANT=_timers timeout 300 python bench.py --steps 50 --warmup 5 --quick > /dev/null 2> $OUT/tiles.err
grep "spmv_cta" $OUT/tiles.err | tee -a $S
fi
if want 3; then
echo "== 3. ILDU (chunked sweeps) and widened rows" | tee -a $S
timeout 300 python bench.py --rows ldu > $OUT/ldu.jsonl 2> $OUT/ldu.err; echo "rc=$?" | tee -a $S
cut -c1-400 $OUT/ldu.jsonl | tee -a $S
timeout 300 python bench.py --rows ldu --lgrid 2048 > $OUT/ldu2048.jsonl 2> $OUT/ldu2048.err; echo "rc=$?" | tee -a $S
cut -c1-400 $OUT/ldu2048.jsonl | tee -a $S
timeout 400 python bench.py --rows widened > $OUT/widened.jsonl 2> $OUT/widened.err; echo "rc=$?" | tee -a $S
cut -c1-300 $OUT/widened.jsonl | tee -a $S
fi
if want 4; then
echo "== 4. irregular SpMV" | tee -a $S
timeout 300 python scripts/spmv_probe.py --kind er --n 2000000 2>> $OUT/er.err | tee -a $OUT/er.jsonl | tee -a $S
timeout 300 python scripts/spmv_probe.py --kind er --n 2000000 --dot 2>> $OUT/er.err | tee -a $OUT/er.jsonl | tee -a $S
timeout 400 python scripts/spmv_probe.py --kind surrogate --n 20000000 --reps 10 2>> $OUT/er.err | tee -a $OUT/er.jsonl | tee -a $S
fi
date | tee -a $S
