#!/bin/bash
# Round 2, 2-GPU visit (charged 2x): multi-GPU PARITY of every opt-in that touches the row-sharded
# path, and the N=2 A/B of the fused all-reduce.  Parity does not need 8 GPUs; the 8-GPU visit
# (charged 8x) is kept for numbers.
#   gpurun --gpus 2 --timeout 1800 -- 'bash scripts/r2_visit_2gpu.sh r2n2'
# Sections: 1 regression (tests/test_gpu_dist.py)  2 opt-in parity  3 N=2 bench A/Bs
make -s -C sigma_b200/csrc all variants > /dev/null 2>&1 || echo "variant build failed (prebuilt .so files are used if present)"
TAG=${1:-r2n2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
S=$OUT/summary.txt
want() { [[ -z "$SECTIONS" || " $SECTIONS " == *" $1 "* ]]; }
run2() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 "$@"; }
date | tee -a $S
if want 1; then
echo "== 1. regression: sharded parity, default path" | tee -a $S
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q > $OUT/pytest_dist.log 2>&1; echo "rc=$?" | tee -a $S
tail -3 $OUT/pytest_dist.log | tee -a $S
fi
if want 2; then
echo "== 2. opt-in parity at world 1..2 (SIGB_TEST_EXPERIMENTAL=1), one test at a time" | tee -a $S
for t in "fence_free_halo" "fused_allreduce" "push_by_the_last_ctas" "rowdirect_spmv and sharded" "single_reduction"; do
  name=$(echo $t | tr ' ' '_')
  SIGB_TEST_EXPERIMENTAL=1 timeout 1200 python -m pytest tests/test_gpu_experimental.py -x -q -k "$t" > $OUT/exp_$name.log 2>&1
  echo "$t rc=$?" | tee -a $S; tail -3 $OUT/exp_$name.log | tee -a $S
done
fi
if want 3; then
echo "== 3. N=2 bench: default, fused all-reduce, fence-free halo, push-last, row-direct" | tee -a $S
for cfg in "" "SIGB_FUSED_ALLREDUCE=1" "SIGB_HALO_LL=1" "SIGB_PUSH_LAST=1" "SIGB_SPMV_ROWDIRECT=1" "SIGB_FUSED_ALLREDUCE=1 SIGB_HALO_LL=1 SIGB_SPMV_ROWDIRECT=1"; do
  env $cfg bash -c "$(declare -f run2); run2 bench.py --gpus 2 --steps 200 --warmup 5 --quick" 2>> $OUT/n2.err | sed "s/^{/{\"env\": \"$cfg\", /" | tee -a $OUT/n2.jsonl | tee -a $S
done
fi
date | tee -a $S
