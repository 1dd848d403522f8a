#!/bin/bash
# Round 2, 1-GPU visit g: consolidated library after the register fix of the persistent kernel's vector
# phases and the shuffle-based chunked ILDU sweeps; the round-1 persistent kernel (3 CTAs/SM build)
# beside it for reference.
TAG=${1:-r2g}; OUT=gpurun_out/$TAG; mkdir -p $OUT
S=$OUT/summary.txt
date | tee -a $S
echo "== pytest -m gpu" | tee -a $S
timeout 1200 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" | tee -a $S
tail -4 $OUT/pytest_gpu.log | tee -a $S
echo "== persistent CG at the shard size: new vs round-1 kernel (3 CTAs/SM build)" | tee -a $S
for v in "" _oldpb3; do for st in 200 20; do
  SIGB_VERBOSE=1 SIGB_CG_PERSISTENT=1 SIGB_LIB_VARIANT=$v timeout 300 python bench.py --grid 1448 --steps $st --warmup 5 --quick --no-parity 2>> $OUT/quick.err | sed "s/^{/{\"grid\": 1448, \"steps\": $st, /" | tee -a $OUT/quick.jsonl | tee -a $S
done; done
grep "sigma_b200:" $OUT/quick.err | sort | uniq -c | tee -a $S
for g in 2048; do for st in 200 20; do
  timeout 300 python bench.py --grid $g --steps $st --warmup 5 --quick 2>> $OUT/quick.err | sed "s/^{/{\"grid\": $g, \"steps\": $st, /" | tee -a $OUT/quick.jsonl | tee -a $S
done; done
SIGB_LIB_VARIANT=_timers timeout 300 python bench.py --grid 1448 --steps 400 --warmup 5 --quick > /dev/null 2> $OUT/phases.err
grep "phase_us" $OUT/phases.err | tee -a $S
echo "== bench, driver flags and 200 steps" | tee -a $S
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_20.json 2> $OUT/bench_20.err; echo "rc=$?" | tee -a $S
cut -c1-300 $OUT/bench_20.json | tee -a $S
timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu > $OUT/bench_200.json 2> $OUT/bench_200.err; echo "rc=$?" | tee -a $S
cut -c1-300 $OUT/bench_200.json | tee -a $S
echo "== ILDU" | tee -a $S
timeout 300 python bench.py --rows ldu > $OUT/ldu.jsonl 2> $OUT/ldu.err; echo "rc=$?" | tee -a $S
cut -c1-330 $OUT/ldu.jsonl | tee -a $S
timeout 300 python bench.py --rows ldu --lgrid 2048 > $OUT/ldu2048.jsonl 2> $OUT/ldu2048.err; echo "rc=$?" | tee -a $S
cut -c1-330 $OUT/ldu2048.jsonl | tee -a $S
date | tee -a $S
