#!/bin/bash
# Round 2, 1-GPU sanitizer visit: compute-sanitizer memcheck and racecheck over the kernels written at the end
# of the round (statically scheduled ILDU sweeps with their shared-memory ring, small tile shape of the CSR
# kernel), then the regular GPU suite once more on the final library.
TAG=${1:-r2san}; OUT=gpurun_out/$TAG; mkdir -p $OUT
S=$OUT/summary.txt
date | tee -a $S
SEL="tests/test_gpu_ldu.py::test_deep_sweeps_bit_exact tests/test_gpu_spmv.py::test_csr_matvec_bit_exact tests/test_gpu_spmv.py::test_small_tile_shape_rows_at_the_cap"
for tool in memcheck racecheck; do
  timeout 500 compute-sanitizer --tool $tool --print-limit 5 python -m pytest $SEL -x -q -m gpu > $OUT/$tool.log 2>&1; echo "$tool rc=$?" | tee -a $S
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Race reported|Invalid" $OUT/$tool.log | sort | uniq -c | head -8 | tee -a $S
done
echo "== pytest -m gpu on the final library" | tee -a $S
timeout 1200 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" | tee -a $S
tail -4 $OUT/pytest_gpu.log | tee -a $S
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_20.json 2> $OUT/bench_20.err; echo "bench rc=$?" | tee -a $S
cut -c1-200 $OUT/bench_20.json | tee -a $S
date | tee -a $S
