#!/bin/bash
# Round 2, 1-GPU visit c: the software-pipelined SpMV (build variant _pipe) -- parity first, then
# A/B against the default and the row-direct form; the N=4 shard size on the persistent kernel;
# the irregular (Erdos-Renyi) SpMV with an ncu capture.
#   gpurun --timeout 1800 -- 'bash scripts/r2_visit_1gpu_c.sh r2c'
TAG=${1:-r2c}; OUT=gpurun_out/$TAG; mkdir -p $OUT
S=$OUT/summary.txt
want() { [[ -z "$SECTIONS" || " $SECTIONS " == *" $1 "* ]]; }
date | tee -a $S
if want 1; then
echo "== 1. parity of the pipelined form: SpMV / solver / operator tests under SIGB_LIB_VARIANT=_pipe" | tee -a $S
for v in _pipe _pipe_t1536r256; do
SIGB_LIB_VARIANT=$v timeout 600 python -m pytest tests/test_gpu_spmv.py tests/test_gpu_solvers.py tests/test_gpu_operators.py -x -q -m gpu > $OUT/pytest$v.log 2>&1
echo "$v rc=$?" | tee -a $S; tail -2 $OUT/pytest$v.log | tee -a $S
done
fi
if want 2; then
echo "== 2. A/B, full size (4096^2) and the 8-GPU shard size (1448^2): default / row-direct / pipelined" | tee -a $S
for g in 4096 1448; do
for cfg in "" "SIGB_SPMV_ROWDIRECT=1" "SIGB_LIB_VARIANT=_pipe" "SIGB_LIB_VARIANT=_pipe_t1536r256" "SIGB_LIB_VARIANT=_pb3" "SIGB_LIB_VARIANT=_pb3 SIGB_SPMV_ROWDIRECT=1" "SIGB_LIB_VARIANT=_pipe_pb3"; do
  env $cfg timeout 300 python bench.py --grid $g --steps 200 --warmup 5 --quick 2>> $OUT/ab.err | sed "s/^{/{\"grid\": $g, \"env\": \"$cfg\", /" | tee -a $OUT/ab.jsonl | tee -a $S
done; done
fi
if want 3; then
echo "== 3. the N=4 shard size (2048^2 = 4.2 M rows) on one GPU: kernel per phase vs persistent (4 and 3 CTAs per SM)" | tee -a $S
for cfg in "SIGB_CG_PERSISTENT=0" "SIGB_CG_PERSISTENT=1" "SIGB_CG_PERSISTENT=1 SIGB_LIB_VARIANT=_pb3" "SIGB_CG_PERSISTENT=1 SIGB_LIB_VARIANT=_pipe_pb3"; do
  env $cfg timeout 300 python bench.py --grid 2048 --steps 200 --warmup 5 --quick 2>> $OUT/n4shard.err | sed "s/^{/{\"grid\": 2048, \"env\": \"$cfg\", /" | tee -a $OUT/n4shard.jsonl | tee -a $S
done
fi
if want 4; then
echo "== 4. irregular SpMV (Erdos-Renyi, config 5 style): 2 M rows exact generator, 20 M rows surrogate" | tee -a $S
for cfg in "" "SIGB_SPMV_ROWDIRECT=1" "SIGB_LIB_VARIANT=_pipe"; do
  env $cfg timeout 300 python scripts/spmv_probe.py --kind er --n 2000000 2>> $OUT/er.err | tee -a $OUT/er.jsonl | tee -a $S
done
for cfg in "" "SIGB_SPMV_ROWDIRECT=1" "SIGB_LIB_VARIANT=_pipe"; do
  env $cfg timeout 400 python scripts/spmv_probe.py --kind surrogate --n 20000000 --reps 10 2>> $OUT/er.err | tee -a $OUT/er.jsonl | tee -a $S
done
echo "-- ncu --set full, ER 2 M rows, default kernel" | tee -a $S
timeout 600 ncu --set full --clock-control none --import-source on -k regex:csr_tma -s 6 -c 1 -f -o $OUT/prof_er2m \
    python scripts/spmv_probe.py --kind er --n 2000000 --reps 3 > $OUT/ncu_er2m.log 2>&1; echo "rc=$?" | tee -a $S
echo "-- ncu --set full, surrogate 20 M rows, default kernel" | tee -a $S
timeout 600 ncu --set full --clock-control none --import-source on -k regex:csr_tma -s 6 -c 1 -f -o $OUT/prof_er20m \
    python scripts/spmv_probe.py --kind surrogate --n 20000000 --reps 3 > $OUT/ncu_er20m.log 2>&1; echo "rc=$?" | tee -a $S
fi
date | tee -a $S
ls -la $OUT >> $S
