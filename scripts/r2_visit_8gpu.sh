#!/bin/bash
# Round 2, 8-GPU visit (charged 8x: keep it short).  Default scaling line, the single-reduction
# persistent CG, the per-phase breakdown of both, then configs 4 / 5 at full size.
#   gpurun --gpus 8 --timeout 1200 -- 'bash scripts/r2_visit_8gpu.sh r2n8'
make -s -C sigma_b200/csrc all variants > /dev/null 2>&1 || echo "variant build failed (prebuilt .so files are used if present)"
TAG=${1:-r2n8}; OUT=gpurun_out/$TAG; mkdir -p $OUT
S=$OUT/summary.txt
run8() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 "$@"; }
echo "== N=8 default" | tee $S
run8 bench.py --gpus 8 --steps 200 --warmup 5 > $OUT/bench_n8.json 2> $OUT/bench_n8.err; echo "rc=$?" | tee -a $S
cat $OUT/bench_n8.json | tee -a $S
echo "== N=8 single reduction (opt-in; parity bars of tests/test_gpu_experimental.py must be green first)" | tee -a $S
SIGB_CG_SINGLE_REDUCE=1 run8 bench.py --gpus 8 --steps 200 --warmup 5 > $OUT/bench_n8_single.json 2> $OUT/bench_n8_single.err; echo "rc=$?" | tee -a $S
cat $OUT/bench_n8_single.json | tee -a $S
echo "== N=8 with the halo push moved to the last CTAs (SIGB_PUSH_LAST=1), default and single reduction" | tee -a $S
for v in "" 1; do
  SIGB_PUSH_LAST=1 SIGB_CG_SINGLE_REDUCE=$v run8 bench.py --gpus 8 --steps 200 --warmup 5 --quick 2>> $OUT/pushlast.err | sed "s/^{/{\"push_last\": 1, \"single_reduce\": \"$v\", /" | tee -a $OUT/pushlast.jsonl | tee -a $S
done
echo "== N=8 fence-free halo (SIGB_HALO_LL=1): parity at world 2..8, then bench alone and with push-last off/on" | tee -a $S
SIGB_TEST_EXPERIMENTAL=1 timeout 1500 python -m pytest tests/test_gpu_experimental.py -x -q -k "fence_free_halo and default" > $OUT/exp_halo_ll.log 2>&1; echo "parity rc=$?" | tee -a $S
tail -3 $OUT/exp_halo_ll.log | tee -a $S
for v in "" 1; do
  SIGB_HALO_LL=1 SIGB_CG_SINGLE_REDUCE=$v run8 bench.py --gpus 8 --steps 200 --warmup 5 --quick 2>> $OUT/halo_ll.err | sed "s/^{/{\"halo_ll\": 1, \"single_reduce\": \"$v\", /" | tee -a $OUT/halo_ll.jsonl | tee -a $S
done
echo "== N=8 row-direct SpMV inside the persistent kernel, alone and with everything else" | tee -a $S
SIGB_SPMV_ROWDIRECT=1 run8 bench.py --gpus 8 --steps 200 --warmup 5 --quick 2>> $OUT/rd8.err | sed "s/^{/{\"rowdirect\": 1, /" | tee -a $OUT/rd8.jsonl | tee -a $S
SIGB_SPMV_ROWDIRECT=1 SIGB_HALO_LL=1 run8 bench.py --gpus 8 --steps 200 --warmup 5 --quick 2>> $OUT/rd8.err | sed "s/^{/{\"rowdirect\": 1, \"halo_ll\": 1, /" | tee -a $OUT/rd8.jsonl | tee -a $S
SIGB_SPMV_ROWDIRECT=1 SIGB_PUSH_LAST=1 SIGB_CG_SINGLE_REDUCE=1 run8 bench.py --gpus 8 --steps 200 --warmup 5 --quick 2>> $OUT/rd8.err | sed "s/^{/{\"rowdirect\": 1, \"push_last\": 1, \"single_reduce\": 1, /" | tee -a $OUT/rd8.jsonl | tee -a $S
echo "== N=8 with the persistent kernels compiled for 3 CTAs per SM (variant _pb3)" | tee -a $S
SIGB_LIB_VARIANT=_pb3 run8 bench.py --gpus 8 --steps 200 --warmup 5 --quick 2>> $OUT/pb3.err | sed "s/^{/{\"pb3\": 1, /" | tee -a $OUT/pb3.jsonl | tee -a $S
echo "== phase breakdown (diagnostic build; its timings are not bench values)" | tee -a $S
for v in "" 1; do
  SIGB_LIB_VARIANT=_timers SIGB_CG_SINGLE_REDUCE=$v run8 bench.py --gpus 8 --steps 200 --warmup 5 --quick > $OUT/phases_single$v.json 2> $OUT/phases_single$v.err
  grep phase_us $OUT/phases_single$v.err | tee -a $S
done
echo "== N=2 and N=4 (kernel-per-phase path): separate all-reduce launches vs fused into their producers" | tee -a $S
SIGB_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_gpu_experimental.py -x -q -k fused_allreduce > $OUT/exp_fused.log 2>&1; echo "parity rc=$?" | tee -a $S
for n in 2 4; do for f in 0 1; do
  SIGB_FUSED_ALLREDUCE=$f timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $n --steps 200 --warmup 5 --quick 2>> $OUT/fused.err | sed "s/^{/{\"fused_allreduce\": $f, /" | tee -a $OUT/fused.jsonl | tee -a $S
done; done
echo "== sharded parity at world 8 (both transports)" | tee -a $S
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q -k "8" > $OUT/pytest_dist8.log 2>&1; echo "rc=$?" | tee -a $S
tail -3 $OUT/pytest_dist8.log | tee -a $S
echo "== configs 4 / 5, full size, 8 GPUs" | tee -a $S
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
  scripts/bench_configs_dist.py > $OUT/configs_full_8gpu.jsonl 2> $OUT/configs_full_8gpu.err; echo "rc=$?" | tee -a $S
cut -c1-700 $OUT/configs_full_8gpu.jsonl | tee -a $S
