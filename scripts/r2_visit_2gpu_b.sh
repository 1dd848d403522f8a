#!/bin/bash
# Round 2, second 2-GPU visit (charged 2x): parity of the row-sharded opt-ins at world 2 straight
# through the worker (one torchrun per combination), then where a persistent-kernel iteration goes
# at the 8-GPU SHARD size (2 ranks x 2.1 M rows = --grid 2048) and the candidates next to it.
#   gpurun --gpus 2 --timeout 1500 -- 'bash scripts/r2_visit_2gpu_b.sh r2b'
TAG=${1:-r2b}; OUT=gpurun_out/$TAG; mkdir -p $OUT
S=$OUT/summary.txt
want() { [[ -z "$SECTIONS" || " $SECTIONS " == *" $1 "* ]]; }
run2() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 "$@"; }
date | tee -a $S
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
if want 1; then
echo "== 1. regression: tests/test_gpu_dist.py (worlds 1, 2; both transports)" | tee -a $S
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q > $OUT/pytest_dist.log 2>&1; echo "rc=$?" | tee -a $S
tail -3 $OUT/pytest_dist.log | tee -a $S
fi
if want 2; then
echo "== 2. opt-in parity at world 2 (tests/dist_gpu_worker.py under each environment)" | tee -a $S
i=0
for cfg in "SIGB_HALO_LL=1" "SIGB_HALO_LL=1 SIGB_SPMV_ROWDIRECT=1" "SIGB_HALO_LL=1 SIGB_CG_PERSISTENT=0" \
           "SIGB_FUSED_ALLREDUCE=1 SIGB_CG_PERSISTENT=0" "SIGB_PUSH_LAST=1" "SIGB_SPMV_ROWDIRECT=1" \
           "SIGB_LIB_VARIANT=_pb3 SIGB_HALO_LL=1 SIGB_SPMV_ROWDIRECT=1"; do
  i=$((i+1))
  env $cfg bash -c "$(declare -f run2); run2 tests/dist_gpu_worker.py" > $OUT/worker_$i.log 2>&1
  echo "[$cfg] rc=$? $(grep 'dist gpu ok' $OUT/worker_$i.log | head -1)" | tee -a $S
done
fi
if want 3; then
echo "== 3. persistent CG, 2 ranks x 2.1 M rows (--grid 2048): default and candidates (200 iterations)" | tee -a $S
for cfg in "" "SIGB_HALO_LL=1" "SIGB_SPMV_ROWDIRECT=1" "SIGB_LIB_VARIANT=_pb3" "SIGB_PUSH_LAST=1" \
           "SIGB_HALO_LL=1 SIGB_SPMV_ROWDIRECT=1" "SIGB_HALO_LL=1 SIGB_SPMV_ROWDIRECT=1 SIGB_LIB_VARIANT=_pb3"; do
  env SIGB_CG_PERSISTENT=1 $cfg bash -c "$(declare -f run2); run2 bench.py --gpus 2 --grid 2048 --steps 200 --warmup 5 --quick" 2>> $OUT/pers2.err \
    | sed "s/^{/{\"env\": \"$cfg\", /" | tee -a $OUT/pers2.jsonl | tee -a $S
done
echo "-- per-phase breakdown (diagnostic build, never a bench value): default, then fence-free halo" | tee -a $S
for cfg in "" "SIGB_HALO_LL=1"; do
  env SIGB_CG_PERSISTENT=1 SIGB_LIB_VARIANT=_timers $cfg bash -c "$(declare -f run2); run2 bench.py --gpus 2 --grid 2048 --steps 200 --warmup 5 --quick" > /dev/null 2> $OUT/phases_$(echo $cfg | tr -d ' =').err
  grep phase_us $OUT/phases_$(echo $cfg | tr -d ' =').err | tee -a $S
done
fi
if want 4; then
echo "== 4. full size at N=2 (kernel-per-phase path): default, fused all-reduce, fence-free halo, row-direct" | tee -a $S
for cfg in "" "SIGB_FUSED_ALLREDUCE=1" "SIGB_HALO_LL=1" "SIGB_SPMV_ROWDIRECT=1" "SIGB_FUSED_ALLREDUCE=1 SIGB_HALO_LL=1 SIGB_SPMV_ROWDIRECT=1"; do
  env $cfg bash -c "$(declare -f run2); run2 bench.py --gpus 2 --steps 200 --warmup 5 --quick" 2>> $OUT/n2.err \
    | sed "s/^{/{\"env\": \"$cfg\", /" | tee -a $OUT/n2.jsonl | tee -a $S
done
echo "-- the contract line at N=2 with the driver's own flags (parity gate inside)" | tee -a $S
run2 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "rc=$?" | tee -a $S
cut -c1-1500 $OUT/bench_n2.json | tee -a $S
fi
date | tee -a $S
