#!/bin/bash
# Round 2, 8-GPU visit j (charged 8x -- every minute here is 8 GPU-minutes): the headline at N=8 and
# N=4 with the parity gate inside, sharded parity at worlds 8 and 4 (both transports, both launch
# models), the phase breakdown at N=8, BASELINE configs 4 and 5 at their named sizes.
#   gpurun --gpus 8 --timeout 1200 -- 'bash scripts/r2_visit_8gpu_j.sh r2j'
TAG=${1:-r2j}; OUT=gpurun_out/$TAG; mkdir -p $OUT
S=$OUT/summary.txt
want() { [[ -z "$SECTIONS" || " $SECTIONS " == *" $1 "* ]]; }
runN() { n=$1; port=$2; shift 2; timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port "$@"; }
date | tee -a $S
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv,noheader > $OUT/gpu.txt 2>&1
if want 1; then
echo "== 1. headline, N=8: driver flags (20 steps), then 200 steps" | tee -a $S
runN 8 29511 bench.py --gpus 8 --steps 20 --warmup 5 > $OUT/bench_n8_20.json 2> $OUT/bench_n8_20.err; echo "rc=$?" | tee -a $S
cut -c1-330 $OUT/bench_n8_20.json | tee -a $S
runN 8 29512 bench.py --gpus 8 --steps 200 --warmup 5 > $OUT/bench_n8_200.json 2> $OUT/bench_n8_200.err; echo "rc=$?" | tee -a $S
cut -c1-330 $OUT/bench_n8_200.json | tee -a $S
echo "-- N=4: driver flags, then 200 steps (quick line)" | tee -a $S
runN 4 29513 bench.py --gpus 4 --steps 20 --warmup 5 > $OUT/bench_n4_20.json 2> $OUT/bench_n4_20.err; echo "rc=$?" | tee -a $S
cut -c1-330 $OUT/bench_n4_20.json | tee -a $S
runN 4 29514 bench.py --gpus 4 --steps 200 --warmup 5 --quick 2>> $OUT/quick.err | tee -a $OUT/quick.jsonl | tee -a $S
fi
if want 2; then
echo "== 2. phase breakdown of the persistent kernel at N=8 (diagnostic build, never a bench value)" | tee -a $S
SIGB_LIB_VARIANT=_timers bash -c "$(declare -f runN); runN 8 29515 bench.py --gpus 8 --steps 200 --warmup 5 --quick" > /dev/null 2> $OUT/phases.err
grep phase_us $OUT/phases.err | cut -c1-700 | tee -a $S
fi
if want 3; then
echo "== 3. sharded parity against the serial oracle: world 8 (peer memory, NCCL), single-process 8 GPUs, C++ program" | tee -a $S
runN 8 29521 tests/dist_gpu_worker.py > $OUT/dist_w8_p2p.log 2>&1; echo "world 8 p2p rc=$? $(grep 'dist gpu ok' $OUT/dist_w8_p2p.log)" | tee -a $S
SIGB_TRANSPORT=nccl bash -c "$(declare -f runN); runN 8 29522 tests/dist_gpu_worker.py" > $OUT/dist_w8_nccl.log 2>&1; echo "world 8 nccl rc=$? $(grep 'dist gpu ok' $OUT/dist_w8_nccl.log)" | tee -a $S
timeout 420 python tests/mgpu_worker.py 8 > $OUT/mgpu_8.log 2>&1; echo "single-process 8 GPUs rc=$? $(grep 'mgpu ok' $OUT/mgpu_8.log)" | tee -a $S
make -s -C tests/cxx > /dev/null 2>&1
timeout 300 tests/cxx/_build/solver_test_multi_gpu -v > $OUT/cxx_multi_gpu_8.log 2>&1; echo "solver_test_multi_gpu (8 GPUs) rc=$?" | tee -a $S
cat $OUT/cxx_multi_gpu_8.log | tee -a $S
echo "-- world 4, both transports side by side on disjoint GPUs; single-process 4 GPUs; the missing-peer fault test" | tee -a $S
( CUDA_VISIBLE_DEVICES=0,1,2,3 bash -c "$(declare -f runN); runN 4 29523 tests/dist_gpu_worker.py" > $OUT/dist_w4_p2p.log 2>&1; echo "world 4 p2p rc=$? $(grep 'dist gpu ok' $OUT/dist_w4_p2p.log)" >> $OUT/par_a.txt ) &
( CUDA_VISIBLE_DEVICES=4,5,6,7 SIGB_TRANSPORT=nccl bash -c "$(declare -f runN); runN 4 29524 tests/dist_gpu_worker.py" > $OUT/dist_w4_nccl.log 2>&1; echo "world 4 nccl rc=$? $(grep 'dist gpu ok' $OUT/dist_w4_nccl.log)" >> $OUT/par_b.txt ) &
wait
( CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 420 python tests/mgpu_worker.py 4 > $OUT/mgpu_4.log 2>&1; echo "single-process 4 GPUs rc=$? $(grep 'mgpu ok' $OUT/mgpu_4.log)" >> $OUT/par_a.txt ) &
( CUDA_VISIBLE_DEVICES=4,5 SIGB_WAIT_TIMEOUT_MS=1500 bash -c "$(declare -f runN); runN 2 29525 tests/dist_fault_worker.py" > $OUT/fault.log 2>&1; echo "missing peer rc=$? $(grep 'fault ok' $OUT/fault.log)" >> $OUT/par_b.txt ) &
wait
cat $OUT/par_a.txt $OUT/par_b.txt | tee -a $S
fi
if want 4; then
echo "== 4. BASELINE configs 4 and 5 at their named sizes on 8 GPUs" | tee -a $S
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
  scripts/bench_configs_dist.py > $OUT/configs_full_8gpu.jsonl 2> $OUT/configs_full_8gpu.err; echo "rc=$?" | tee -a $S
cut -c1-900 $OUT/configs_full_8gpu.jsonl | tee -a $S
tail -3 $OUT/configs_full_8gpu.err | tee -a $S
fi
date | tee -a $S
