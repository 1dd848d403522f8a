#!/bin/bash
# Round 2, 8-GPU visit (charged 8x: every run here costs ~4-8 GPU-minutes, pick sections).  Parity of
# the opt-ins is checked on 2 GPUs first (scripts/r2_visit_2gpu.sh); this visit is for numbers.
#   gpurun --gpus 8 --timeout 1200 -- 'SECTIONS="1 2" bash scripts/r2_visit_8gpu.sh r2n8'
# Sections: 1 default line + per-phase breakdown   2 each candidate alone
#           3 combinations                         4 N=4 fused all-reduce A/B
#           5 sharded parity at world 8            6 BASELINE configs 4 / 5 at full size
make -s -C sigma_b200/csrc all variants > /dev/null 2>&1 || echo "variant build failed (prebuilt .so files are used if present)"
TAG=${1:-r2n8}; OUT=gpurun_out/$TAG; mkdir -p $OUT
S=$OUT/summary.txt
want() { [[ -z "$SECTIONS" || " $SECTIONS " == *" $1 "* ]]; }
runN() { n=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 "$@"; }
quick8() {   # quick8 "<env assignments>": one --quick line at N=8 tagged with its environment
  env $1 bash -c "$(declare -f runN); runN 8 bench.py --gpus 8 --steps 200 --warmup 5 --quick" 2>> $OUT/quick8.err \
    | sed "s/^{/{\"env\": \"$1\", /" | tee -a $OUT/quick8.jsonl | tee -a $S
}
date | tee -a $S
if want 1; then
echo "== 1. N=8 default (full JSON line), then the per-phase breakdown of the default (diagnostic build, not a bench value)" | tee -a $S
runN 8 bench.py --gpus 8 --steps 200 --warmup 5 > $OUT/bench_n8.json 2> $OUT/bench_n8.err; echo "rc=$?" | tee -a $S
cat $OUT/bench_n8.json | tee -a $S
SIGB_LIB_VARIANT=_timers runN 8 bench.py --gpus 8 --steps 200 --warmup 5 --quick > $OUT/phases.json 2> $OUT/phases.err
grep phase_us $OUT/phases.err | tee -a $S
fi
if want 2; then
echo "== 2. candidates, one at a time" | tee -a $S
for cfg in "SIGB_HALO_LL=1" "SIGB_PUSH_LAST=1" "SIGB_LIB_VARIANT=_pb3" "SIGB_SPMV_ROWDIRECT=1" "SIGB_CG_SINGLE_REDUCE=1"; do quick8 "$cfg"; done
fi
if want 3; then
echo "== 3. combinations (reference statement order first, then with the single reduction)" | tee -a $S
quick8 "SIGB_HALO_LL=1 SIGB_SPMV_ROWDIRECT=1"
quick8 "SIGB_HALO_LL=1 SIGB_SPMV_ROWDIRECT=1 SIGB_LIB_VARIANT=_pb3"
quick8 "SIGB_HALO_LL=1 SIGB_SPMV_ROWDIRECT=1 SIGB_CG_SINGLE_REDUCE=1"
fi
if want 4; then
echo "== 4. N=4 (kernel-per-phase path): separate all-reduce launches vs fused into their producers" | tee -a $S
for f in 0 1; do
  SIGB_FUSED_ALLREDUCE=$f runN 4 bench.py --gpus 4 --steps 200 --warmup 5 --quick 2>> $OUT/fused.err | sed "s/^{/{\"fused_allreduce\": $f, /" | tee -a $OUT/fused.jsonl | tee -a $S
done
fi
if want 5; then
echo "== 5. sharded parity at world 8 (both transports)" | tee -a $S
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q -k "8" > $OUT/pytest_dist8.log 2>&1; echo "rc=$?" | tee -a $S
tail -3 $OUT/pytest_dist8.log | tee -a $S
fi
if want 6; then
echo "== 6. configs 4 / 5, full size, 8 GPUs (generation ~1-2 min per rank, in parallel)" | tee -a $S
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
  scripts/bench_configs_dist.py > $OUT/configs_full_8gpu.jsonl 2> $OUT/configs_full_8gpu.err; echo "rc=$?" | tee -a $S
cut -c1-700 $OUT/configs_full_8gpu.jsonl | tee -a $S
fi
date | tee -a $S
