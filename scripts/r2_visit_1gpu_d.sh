#!/bin/bash
# Round 2, 1-GPU visit d: pipelined SpMV with the fused dot taken out of the row sums (_pipe2).
TAG=${1:-r2d}; OUT=gpurun_out/$TAG; mkdir -p $OUT
S=$OUT/summary.txt
date | tee -a $S
echo "== parity under _pipe2" | tee -a $S
SIGB_LIB_VARIANT=_pipe2 timeout 600 python -m pytest tests/test_gpu_spmv.py tests/test_gpu_solvers.py tests/test_gpu_operators.py -x -q -m gpu > $OUT/pytest_pipe2.log 2>&1
echo "rc=$?" | tee -a $S; tail -2 $OUT/pytest_pipe2.log | tee -a $S
for g in 4096 2048 1448; do
for cfg in "SIGB_LIB_VARIANT=_pipe2" "SIGB_LIB_VARIANT=_pipe2_pb3" "SIGB_CG_PERSISTENT=1 SIGB_LIB_VARIANT=_pipe2_pb3"; do
  env $cfg timeout 300 python bench.py --grid $g --steps 200 --warmup 5 --quick 2>> $OUT/ab.err | sed "s/^{/{\"grid\": $g, \"env\": \"$cfg\", /" | tee -a $OUT/ab.jsonl | tee -a $S
done; done
for cfg in "SIGB_LIB_VARIANT=_pipe2"; do
  env $cfg timeout 300 python scripts/spmv_probe.py --kind er --n 2000000 --dot 2>> $OUT/er.err | tee -a $OUT/er.jsonl | tee -a $S
done
date | tee -a $S
