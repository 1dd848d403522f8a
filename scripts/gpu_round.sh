#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list, ncu full capture.
# Usage (from the repo root on the box): bash scripts/gpu_round.sh [tag]
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu" | tee $OUT/summary.txt
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -15 $OUT/pytest_gpu.log | tee -a $OUT/summary.txt
echo "== smoke" | tee -a $OUT/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/summary.txt
tail -3 $OUT/smoke.log | tee -a $OUT/summary.txt
echo "== bench" | tee -a $OUT/summary.txt
timeout 600 python bench.py --steps 200 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/summary.txt
cat $OUT/bench.json | tee -a $OUT/summary.txt; tail -5 $OUT/bench.err | tee -a $OUT/summary.txt
echo "== ncu launch list" | tee -a $OUT/summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 32 --warmup 3 --no-cpu > $OUT/bench_under_ncu.log 2>&1; echo "ncu list rc=$?" | tee -a $OUT/summary.txt
echo "== ncu full (csr_stream)" | tee -a $OUT/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:csr_tma -s 20 -c 3 -f -o $OUT/prof_csr \
    python bench.py --steps 48 --warmup 3 --no-cpu > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?" | tee -a $OUT/summary.txt
echo "== ncu full (CgUpdate)" | tee -a $OUT/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"CgUpdateOp|CgDirectionOp" -s 34 -c 2 -f -o $OUT/prof_update \
    python bench.py --steps 48 --warmup 3 --no-cpu > $OUT/ncu_full2.log 2>&1; echo "ncu full2 rc=$?" | tee -a $OUT/summary.txt
ls -la $OUT | tee -a $OUT/summary.txt
