#!/bin/bash
# Round 2, 1-GPU visit: regression, then every opt-in path written after the round-1 GPU budget
# ran out (none of them has run on a GPU yet), each next to its default.
#   gpurun --timeout 3000 -- 'bash scripts/r2_visit_1gpu.sh r2a'            # everything (~45 min)
#   gpurun --timeout 1200 -- 'SECTIONS="1 2" bash scripts/r2_visit_1gpu.sh r2a'   # a subset
# Sections: 1 regression  2 experimental parity tests  3 bench + SpMV A/B  4 ILDU sweeps
#           5 copies / assembly  6 persistent CG at shard size + phase clocks  7 configs 4 / 5 full size
# Read gpurun_out/<tag>/summary.txt first.
make -s -C sigma_b200/csrc all variants > /dev/null 2>&1 || echo "variant build failed (prebuilt .so files are used if present)"
TAG=${1:-r2a}; OUT=gpurun_out/$TAG; mkdir -p $OUT
S=$OUT/summary.txt
want() { [[ -z "$SECTIONS" || " $SECTIONS " == *" $1 "* ]]; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
date | tee -a $S
if want 1; then
echo "== 1. regression: pytest -m gpu" | tee -a $S
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" | tee -a $S
tail -5 $OUT/pytest_gpu.log | tee -a $S
fi
if want 2; then
echo "== 2. experimental paths (SIGB_TEST_EXPERIMENTAL=1), one test at a time so a hang costs one timeout" | tee -a $S
for t in test_device_built_tiles_equal_the_host_tiling test_copy_and_transpose_parity_with_device_tiles \
         test_copy_assembly_parity_with_async_scratch \
         test_parity_with_rowdirect_spmv test_persistent_cg_with_rowdirect_spmv test_ldu_parity_with_syncfree_sweeps \
         test_single_reduction_persistent_cg test_bicgstab_with_ldu_preconditioner test_matrix_test_strategy_on_the_device test_matrix_test_set_multiple_entries_on_the_device; do
  SIGB_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_experimental.py -x -q -k $t > $OUT/exp_$t.log 2>&1
  echo "$t rc=$?" | tee -a $S; tail -3 $OUT/exp_$t.log | tee -a $S
done
SIGB_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_cxx_host.py -x -q -k not_yet_run > $OUT/exp_cxx.log 2>&1; echo "cxx not-yet-run programs rc=$?" | tee -a $S
tail -3 $OUT/exp_cxx.log | tee -a $S
fi
if want 3; then
echo "== 3. bench (default path)" | tee -a $S
timeout 600 python bench.py --steps 200 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?" | tee -a $S
cat $OUT/bench.json | tee -a $S
echo "== 3b. SpMV A/B: two-pass (default) vs row-direct, full size and the 8-GPU shard size" | tee -a $S
for g in 4096 1448; do for rd in 0 1; do
  SIGB_SPMV_ROWDIRECT=$rd timeout 300 python bench.py --grid $g --steps 200 --warmup 5 --quick 2>> $OUT/rowdirect.err | sed "s/^{/{\"grid\": $g, \"rowdirect\": $rd, /" | tee -a $OUT/rowdirect.jsonl | tee -a $S
done; done
# 256-row tiles: with row-direct every thread owns exactly one row of a 5-point tile (built here as
# make VARIANT=_t1536r256 DEFS="-DSIGB_TILE_NNZ=1536 -DSIGB_TILE_ROWS=256")
for rd in 0 1; do
  SIGB_LIB_VARIANT=_t1536r256 SIGB_SPMV_ROWDIRECT=$rd timeout 300 python bench.py --steps 200 --warmup 5 --quick 2>> $OUT/rowdirect.err | sed "s/^{/{\"grid\": 4096, \"rowdirect\": $rd, /" | tee -a $OUT/rowdirect.jsonl | tee -a $S
done
# the two-pass kernel compiled with a minimum of 4 resident CTAs named (ptxas: 32 -> up to 64 registers),
# built here as make VARIANT=_mb4 DEFS=-DSIGB_SPMV_MINBLOCKS=4
SIGB_LIB_VARIANT=_mb4 timeout 300 python bench.py --steps 200 --warmup 5 --quick 2>> $OUT/rowdirect.err | sed "s/^{/{\"grid\": 4096, \"rowdirect\": 0, /" | tee -a $OUT/rowdirect.jsonl | tee -a $S
SIGB_LIB_VARIANT=_timers SIGB_SPMV_ROWDIRECT=1 timeout 300 python bench.py --steps 50 --warmup 3 --quick > /dev/null 2> $OUT/spmv_tiles_rowdirect.err
grep spmv_cta_pass $OUT/spmv_tiles_rowdirect.err | tee -a $S
fi
if want 4; then
echo "== 4. ILDU: per-level launches vs sync-free sweeps" | tee -a $S
timeout 300 python bench.py --rows ldu > $OUT/ldu_default.jsonl 2> $OUT/ldu_default.err; echo "rc=$?" | tee -a $S
for cfg in "PER_SM=1 SLEEP=0" "PER_SM=2 SLEEP=0" "PER_SM=1 SLEEP=32" "CTAS=64 SLEEP=0" "CTAS=64 SLEEP=32" "CTAS=32 SLEEP=0" "CTAS=16 SLEEP=0"; do
  eval $cfg; tag=$(echo $cfg | tr ' =' '__')
  SIGB_LDU_SYNCFREE=1 SIGB_LDU_SF_CTAS_PER_SM=${PER_SM:-1} SIGB_LDU_SF_CTAS=${CTAS:-0} SIGB_LDU_SF_SLEEP_NS=$SLEEP \
    timeout 300 python bench.py --rows ldu > $OUT/ldu_syncfree_$tag.jsonl 2> $OUT/ldu_syncfree_$tag.err
  echo "syncfree $cfg rc=$?" | tee -a $S
  unset PER_SM CTAS SLEEP
done
grep -h "ldu apply\|ldu setup\|CG iterations" $OUT/ldu_*.jsonl | cut -c1-300 | tee -a $S
fi
if want 5; then
echo "== 5. copies / assembly: host tiling vs device tiling" | tee -a $S
timeout 400 python bench.py --rows widened > $OUT/widened_default.jsonl 2> $OUT/widened_default.err; echo "rc=$?" | tee -a $S
SIGB_DEVICE_TILES=1 timeout 400 python bench.py --rows widened > $OUT/widened_devtiles.jsonl 2> $OUT/widened_devtiles.err; echo "rc=$?" | tee -a $S
SIGB_DEVICE_TILES=1 SIGB_ASYNC_ALLOC=1 timeout 400 python bench.py --rows widened > $OUT/widened_devtiles_async.jsonl 2> $OUT/widened_devtiles_async.err; echo "rc=$?" | tee -a $S
grep -h "copy_matrix\|add_value" $OUT/widened_*.jsonl | cut -c1-260 | tee -a $S
fi
if want 6; then
echo "== 6. persistent CG at the 8-GPU shard size on one GPU: default / single reduction / phase breakdown" | tee -a $S
for v in "" 1; do
  SIGB_CG_PERSISTENT=1 SIGB_CG_SINGLE_REDUCE=$v timeout 300 python bench.py --grid 1448 --steps 400 --warmup 5 --quick 2>> $OUT/pers.err | sed "s/^{/{\"single_reduce\": \"$v\", /" | tee -a $OUT/pers.jsonl | tee -a $S
done
# persistent kernels compiled for 3 resident CTAs per SM (80 registers, almost no spills), built here
# as make VARIANT=_pb3 DEFS=-DSIGB_PERSIST_MINBLOCKS=3
for rd in 0 1; do
  SIGB_LIB_VARIANT=_pb3 SIGB_CG_PERSISTENT=1 SIGB_SPMV_ROWDIRECT=$rd timeout 300 python bench.py --grid 1448 --steps 400 --warmup 5 --quick 2>> $OUT/pers.err | sed "s/^{/{\"pb3\": 1, \"rowdirect\": $rd, /" | tee -a $OUT/pers.jsonl | tee -a $S
done
SIGB_CG_PERSISTENT=1 SIGB_SPMV_ROWDIRECT=1 timeout 300 python bench.py --grid 1448 --steps 400 --warmup 5 --quick 2>> $OUT/pers.err | sed "s/^{/{\"rowdirect\": 1, /" | tee -a $OUT/pers.jsonl | tee -a $S
for c in 2 3; do
  SIGB_CG_PERSISTENT=1 SIGB_CG_PERSIST_CTAS_PER_SM=$c timeout 300 python bench.py --grid 1448 --steps 400 --warmup 5 --quick 2>> $OUT/pers.err | sed "s/^{/{\"ctas_per_sm\": $c, /" | tee -a $OUT/pers.jsonl | tee -a $S
done
for v in "" 1; do
  SIGB_LIB_VARIANT=_timers SIGB_CG_PERSISTENT=1 SIGB_CG_SINGLE_REDUCE=$v timeout 300 python bench.py --grid 1448 --steps 400 --warmup 5 --quick > /dev/null 2> $OUT/phases_single$v.err
  grep phase_us $OUT/phases_single$v.err | tee -a $S
done
echo "== 6b. where a CTA of the SpMV kernel spends its pass (diagnostic build), full size" | tee -a $S
SIGB_LIB_VARIANT=_timers timeout 300 python bench.py --steps 50 --warmup 3 --quick > /dev/null 2> $OUT/spmv_tiles.err
grep spmv_cta_pass $OUT/spmv_tiles.err | tee -a $S
fi
if want 7; then
echo "== 7. BASELINE configs 4 and 5 at full size on one GPU" | tee -a $S
timeout 900 python scripts/bench_configs_dist.py > $OUT/configs_full_1gpu.jsonl 2> $OUT/configs_full_1gpu.err; echo "rc=$?" | tee -a $S
cut -c1-700 $OUT/configs_full_1gpu.jsonl | tee -a $S
fi
date | tee -a $S
ls -la $OUT >> $S
