#!/bin/bash
# Round 2, 1-GPU visit k: the ceiling of the x gathers (pure random-gather probe) beside the CSR kernel
# on the Erdos-Renyi operators, an ncu capture of the CURRENT kernel on the 2 M-row operator, and the
# first GPU run of the strict-order solves (whole solve bit for bit against the serial loops).
TAG=${1:-r2k}; OUT=gpurun_out/$TAG; mkdir -p $OUT
S=$OUT/summary.txt
date | tee -a $S
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv,noheader | tee -a $S
echo "== strict-order solves, C++ programs" | tee -a $S
timeout 600 python -m pytest tests/test_gpu_solvers.py tests/test_cxx_host.py -x -q -m gpu > $OUT/pytest_strict.log 2>&1; echo "rc=$?" | tee -a $S
tail -5 $OUT/pytest_strict.log | tee -a $S
echo "== random-gather probe (no matrix, no row sums: what the memory system gives the gathers alone)" | tee -a $S
timeout 300 scripts/_build/gather_probe > $OUT/gather.jsonl 2> $OUT/gather.err; echo "rc=$?" | tee -a $S
cat $OUT/gather.jsonl | tee -a $S
echo "== CSR kernel on the ER operators" | tee -a $S
timeout 300 python scripts/spmv_probe.py --kind er --n 2000000 2>> $OUT/er.err | tee -a $OUT/er.jsonl | tee -a $S
timeout 400 python scripts/spmv_probe.py --kind surrogate --n 20000000 --reps 10 2>> $OUT/er.err | tee -a $OUT/er.jsonl | tee -a $S
echo "-- ncu --set full, ER 2 M rows, current kernel" | tee -a $S
timeout 600 ncu --set full --clock-control none --import-source on -k regex:csr_tma -s 6 -c 1 -f -o $OUT/prof_er2m \
    python scripts/spmv_probe.py --kind er --n 2000000 --reps 3 > $OUT/ncu_er2m.log 2>&1; echo "rc=$?" | tee -a $S
date | tee -a $S
