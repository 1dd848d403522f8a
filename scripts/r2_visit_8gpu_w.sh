#!/bin/bash
# Round 2, 8-GPU visit w (charged 8x): BASELINE config 5 (ER, 20 M vertices, BiCGSTAB) at its named size with
# the small tile shape and the push-by-every-CTA halo exchange (visit j had 4.37 ms per SpMV, 115.6 it/s).
TAG=${1:-r2w}; OUT=gpurun_out/$TAG; mkdir -p $OUT
S=$OUT/summary.txt
date | tee -a $S
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
  scripts/bench_configs_dist.py --skip c4,c5l > $OUT/c5_8gpu.jsonl 2> $OUT/c5_8gpu.err; echo "rc=$?" | tee -a $S
cut -c1-1200 $OUT/c5_8gpu.jsonl | tee -a $S
tail -3 $OUT/c5_8gpu.err | tee -a $S
date | tee -a $S
