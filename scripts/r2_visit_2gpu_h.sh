#!/bin/bash
# Round 2, 2-GPU visit h (charged 2x): the consolidated library on two GPUs -- sharded parity (both
# launch models: one process per GPU, and one process for all GPUs through sigb_mgpu_*), the missing-peer
# fault test, the persistent kernel at the 8-GPU shard size with its phase breakdown, and N=2 at full size.
TAG=${1:-r2h}; OUT=gpurun_out/$TAG; mkdir -p $OUT
S=$OUT/summary.txt
run2() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 "$@"; }
date | tee -a $S
echo "== 1. parity: row-sharded operators, one process per GPU and single-process multi-GPU; C++ program" | tee -a $S
timeout 1200 python -m pytest tests/test_gpu_dist.py tests/test_gpu_mgpu.py -x -q > $OUT/pytest_multi.log 2>&1; echo "rc=$?" | tee -a $S
tail -5 $OUT/pytest_multi.log | tee -a $S
make -s -C tests/cxx > /dev/null 2>&1
timeout 300 tests/cxx/_build/solver_test_multi_gpu -v > $OUT/cxx_multi_gpu.log 2>&1; echo "solver_test_multi_gpu rc=$?" | tee -a $S
cat $OUT/cxx_multi_gpu.log | tee -a $S
echo "== 2. persistent CG, 2 ranks x 2.1 M rows (--grid 2048), 200 and 20 steps; phase breakdown" | tee -a $S
for st in 200 20; do
  run2 bench.py --gpus 2 --grid 2048 --steps $st --warmup 5 --quick 2>> $OUT/pers2.err | sed "s/^{/{\"steps\": $st, /" | tee -a $OUT/pers2.jsonl | tee -a $S
done
SIGB_LIB_VARIANT=_oldpb3 SIGB_CG_PERSISTENT=1 bash -c "$(declare -f run2); run2 bench.py --gpus 2 --grid 2048 --steps 200 --warmup 5 --quick --no-parity" 2>> $OUT/pers2.err | sed "s/^{/{\"round1_kernel_3ctas\": 1, /" | tee -a $OUT/pers2.jsonl | tee -a $S
SIGB_LIB_VARIANT=_timers bash -c "$(declare -f run2); run2 bench.py --gpus 2 --grid 2048 --steps 200 --warmup 5 --quick" > /dev/null 2> $OUT/phases.err
grep phase_us $OUT/phases.err | tee -a $S
echo "== 3. full size at N=2: driver flags, then 200 steps" | tee -a $S
run2 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_n2_20.json 2> $OUT/bench_n2_20.err; echo "rc=$?" | tee -a $S
cut -c1-330 $OUT/bench_n2_20.json | tee -a $S
run2 bench.py --gpus 2 --steps 200 --warmup 5 > $OUT/bench_n2_200.json 2> $OUT/bench_n2_200.err; echo "rc=$?" | tee -a $S
cut -c1-330 $OUT/bench_n2_200.json | tee -a $S
echo "== 4. chunked ILDU sweeps: trip statistics (diagnostic build)" | tee -a $S
SIGB_LIB_VARIANT=_sweepstats timeout 300 python bench.py --rows ldu > $OUT/ldu_stats.log 2>&1
grep "^sweep" $OUT/ldu_stats.log | tail -12 | tee -a $S
date | tee -a $S
