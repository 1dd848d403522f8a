#!/bin/bash
# Round 2, 1-GPU visit x: the state of the library at the end of the round -- full GPU test suite, smoke,
# both bench arms with the driver's flags, 200-step bench, launch list and ncu --set full of the dominant
# kernel, the widened rows (operators / copies / assembly, ILDU, Lanczos), the ER probe.
TAG=${1:-r2x}; OUT=gpurun_out/$TAG; mkdir -p $OUT
S=$OUT/summary.txt
want() { [[ -z "$SECTIONS" || " $SECTIONS " == *" $1 "* ]]; }
date | tee -a $S
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv,noheader | tee -a $S
if want 1; then
echo "== 1. pytest -m gpu, smoke" | tee -a $S
timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" | tee -a $S
tail -6 $OUT/pytest_gpu.log | tee -a $S
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $S
tail -2 $OUT/smoke.log | tee -a $S
fi
if want 2; then
echo "== 2. bench: reference arm, our arm (driver flags), 200 steps" | tee -a $S
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_ref_20.json 2> $OUT/bench_ref_20.err; echo "rc=$?" | tee -a $S
cut -c1-300 $OUT/bench_ref_20.json | tee -a $S
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_20.json 2> $OUT/bench_20.err; echo "rc=$?" | tee -a $S
cut -c1-3000 $OUT/bench_20.json | tee -a $S
timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu > $OUT/bench_200.json 2> $OUT/bench_200.err; echo "rc=$?" | tee -a $S
cut -c1-400 $OUT/bench_200.json | tee -a $S
fi
if want 3; then
echo "== 3. launch list of the bench step; ncu --set full of the dominant kernel (SpMV + dot) and of the CG vector kernels" | tee -a $S
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu --no-parity > $OUT/launches_bench.log 2>&1; echo "rc=$?" | tee -a $S
timeout 900 ncu --set full --clock-control none --import-source on -k regex:csr_tma -s 20 -c 2 -f \
    -o $OUT/prof_csr python bench.py --steps 48 --warmup 3 --no-cpu --no-parity > $OUT/ncu_csr.log 2>&1; echo "rc=$?" | tee -a $S
fi
if want 4; then
echo "== 4. widened rows: operators / copies / assembly, ILDU, Lanczos" | tee -a $S
timeout 400 python bench.py --rows widened > $OUT/widened.jsonl 2> $OUT/widened.err; echo "rc=$?" | tee -a $S
cut -c1-260 $OUT/widened.jsonl | tee -a $S
timeout 300 python bench.py --rows ldu > $OUT/ldu.jsonl 2> $OUT/ldu.err; echo "rc=$?" | tee -a $S
cut -c1-330 $OUT/ldu.jsonl | tee -a $S
timeout 300 python bench.py --rows ldu --lgrid 2048 > $OUT/ldu2048.jsonl 2> $OUT/ldu2048.err; echo "rc=$?" | tee -a $S
cut -c1-330 $OUT/ldu2048.jsonl | tee -a $S
timeout 400 python bench.py --rows lanczos --lanczos-n-big 20000000 > $OUT/lanczos.jsonl 2> $OUT/lanczos.err; echo "rc=$?" | tee -a $S
cut -c1-700 $OUT/lanczos.jsonl | tee -a $S
fi
if want 5; then
echo "== 5. ER operators" | tee -a $S
timeout 300 python scripts/spmv_probe.py --kind er --n 2000000 2>> $OUT/er.err | tee -a $OUT/er.jsonl | cut -c1-400 | tee -a $S
timeout 300 python scripts/spmv_probe.py --kind er --n 2000000 --dot 2>> $OUT/er.err | tee -a $OUT/er.jsonl | cut -c1-400 | tee -a $S
timeout 400 python scripts/spmv_probe.py --kind surrogate --n 20000000 --reps 10 2>> $OUT/er.err | tee -a $OUT/er.jsonl | cut -c1-400 | tee -a $S
fi
date | tee -a $S
