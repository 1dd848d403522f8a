#!/bin/bash
# Round 2, 2-GPU visit r (charged 2x): sharded parity with the small tile shape and the push-by-every-CTA
# halo exchange of operators without locality (both launch models, both transports), the multi-GPU C++
# program, and config 5 at a quarter of its size with the two push modes side by side.
TAG=${1:-r2r}; OUT=gpurun_out/$TAG; mkdir -p $OUT
S=$OUT/summary.txt
run2() { timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
date | tee -a $S
echo "== 1. parity" | tee -a $S
timeout 1200 python -m pytest tests/test_gpu_dist.py tests/test_gpu_mgpu.py -x -q > $OUT/pytest_multi.log 2>&1; echo "rc=$?" | tee -a $S
tail -4 $OUT/pytest_multi.log | tee -a $S
make -s -C tests/cxx > /dev/null 2>&1
timeout 300 tests/cxx/_build/solver_test_multi_gpu -v > $OUT/cxx_multi_gpu.log 2>&1; echo "solver_test_multi_gpu rc=$?" | tee -a $S
cat $OUT/cxx_multi_gpu.log | tee -a $S
echo "== 2. config 5 (ER, BiCGSTAB) at n = 5 M on 2 GPUs: push by every CTA vs dedicated communication CTAs; large tile shape" | tee -a $S
for cfg in "SIGB_PUSH_ALL=1" "SIGB_PUSH_ALL=0" "SIGB_PUSH_ALL=1 SIGB_TILE_CLASS=0"; do
  env $cfg bash -c "$(declare -f run2); run2 29531 scripts/bench_configs_dist.py --er-n 5000000 --skip c4,c5l" 2>> $OUT/c5.err | sed "s/^{/{\"env\": \"$cfg\", /" | tee -a $OUT/c5.jsonl | cut -c1-700 | tee -a $S
done
date | tee -a $S
