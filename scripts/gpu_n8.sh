#!/bin/bash
# 8-GPU visit, short: sharded parity at world 8 is covered by bench residuals; runs bench N=1 (quick) and N=8.
TAG=${1:-n8}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 300 python bench.py --gpus 1 --steps 200 --warmup 5 --quick > $OUT/bench_n1_quick.json 2> $OUT/n1.err; cat $OUT/bench_n1_quick.json
for pers in auto 0; do
  if [ $pers = auto ]; then unset SIGB_CG_PERSISTENT; else export SIGB_CG_PERSISTENT=$pers; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 8 --steps 200 --warmup 5 > $OUT/bench_n8_pers_$pers.json 2> $OUT/bench_n8_pers_$pers.err
  echo "pers=$pers rc=$?"; cat $OUT/bench_n8_pers_$pers.json; tail -3 $OUT/bench_n8_pers_$pers.err
done
