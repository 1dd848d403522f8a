#!/bin/bash
# Round 2, 1-GPU visit m: the small tile shape of the CSR kernel (gather-bound operators) against the
# large one on the Erdos-Renyi operators, residency / carve-out sweep, parity of the SpMV tests.
TAG=${1:-r2m}; OUT=gpurun_out/$TAG; mkdir -p $OUT
S=$OUT/summary.txt
date | tee -a $S
echo "== parity: SpMV / convert / solver tests" | tee -a $S
timeout 900 python -m pytest tests/test_gpu_spmv.py tests/test_gpu_convert.py tests/test_gpu_solvers.py tests/test_gpu_operators.py -x -q -m gpu > $OUT/pytest.log 2>&1; echo "rc=$?" | tee -a $S
tail -5 $OUT/pytest.log | tee -a $S
echo "== ER 2 M rows: large shape, small shape (residency x carve-out)" | tee -a $S
for cfg in "SIGB_TILE_CLASS=0" "SIGB_TILE_CLASS=1" "SIGB_SMALL_TILE_CTAS=3" "SIGB_SMALL_TILE_CTAS=5" "SIGB_SMALL_TILE_CTAS=6" \
           "SIGB_SMALL_TILE_CTAS=4 SIGB_SMALL_TILE_CARVEOUT_KB=100" "SIGB_SMALL_TILE_CTAS=8 SIGB_SMALL_TILE_CARVEOUT_KB=228"; do
  env $cfg timeout 300 python scripts/spmv_probe.py --kind er --n 2000000 2>> $OUT/er.err | tee -a $OUT/er.jsonl | cut -c1-420 | tee -a $S
done
env SIGB_TILE_CLASS=1 timeout 300 python scripts/spmv_probe.py --kind er --n 2000000 --dot 2>> $OUT/er.err | tee -a $OUT/er.jsonl | cut -c1-420 | tee -a $S
echo "== surrogate 20 M rows" | tee -a $S
for cfg in "SIGB_TILE_CLASS=0" "SIGB_TILE_CLASS=1" "SIGB_SMALL_TILE_CTAS=6"; do
  env $cfg timeout 400 python scripts/spmv_probe.py --kind surrogate --n 20000000 --reps 10 2>> $OUT/er.err | tee -a $OUT/er.jsonl | cut -c1-420 | tee -a $S
done
echo "== Poisson 4096^2 with the small shape forced (the large one is the default there)" | tee -a $S
for cfg in "SIGB_TILE_CLASS=0" "SIGB_TILE_CLASS=1" "SIGB_TILE_CLASS=1 SIGB_SMALL_TILE_CTAS=6"; do
  env $cfg timeout 300 python scripts/spmv_probe.py --kind poisson --n 16777216 2>> $OUT/er.err | tee -a $OUT/er.jsonl | cut -c1-420 | tee -a $S
done
echo "-- ncu --set full, ER 2 M rows, small shape" | tee -a $S
timeout 600 ncu --set full --clock-control none --import-source on -k regex:csr_tma -s 6 -c 1 -f -o $OUT/prof_er2m_small \
    python scripts/spmv_probe.py --kind er --n 2000000 --reps 3 > $OUT/ncu_er2m.log 2>&1; echo "rc=$?" | tee -a $S
date | tee -a $S
