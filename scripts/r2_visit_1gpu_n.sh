#!/bin/bash
# Round 2, 1-GPU visit n: small tile shape at 512 staged entries (build variant _s512) against 1024,
# residency sweep, on the ER operators and the Poisson matrix; fused dot on Poisson.
TAG=${1:-r2n}; OUT=gpurun_out/$TAG; mkdir -p $OUT
S=$OUT/summary.txt
date | tee -a $S
run() { env $1 timeout 400 python scripts/spmv_probe.py $2 2>> $OUT/er.err | tee -a $OUT/er.jsonl | cut -c1-400 | tee -a $S; }
echo "== ER 2 M rows" | tee -a $S
for cfg in "SIGB_SMALL_TILE_CTAS=5" "SIGB_LIB_VARIANT=_s512 SIGB_SMALL_TILE_CTAS=8" "SIGB_LIB_VARIANT=_s512 SIGB_SMALL_TILE_CTAS=9" \
           "SIGB_LIB_VARIANT=_s512 SIGB_SMALL_TILE_CTAS=7" "SIGB_LIB_VARIANT=_s512 SIGB_SMALL_TILE_CTAS=6" "SIGB_LIB_VARIANT=_s512 SIGB_SMALL_TILE_CTAS=10"; do
  run "$cfg" "--kind er --n 2000000"
done
echo "== surrogate 20 M rows" | tee -a $S
for cfg in "SIGB_SMALL_TILE_CTAS=5" "SIGB_LIB_VARIANT=_s512 SIGB_SMALL_TILE_CTAS=8" "SIGB_LIB_VARIANT=_s512 SIGB_SMALL_TILE_CTAS=6"; do
  run "$cfg" "--kind surrogate --n 20000000 --reps 10"
done
echo "== Poisson 4096^2, plain and fused with the dot" | tee -a $S
for cfg in "SIGB_TILE_CLASS=0" "SIGB_TILE_CLASS=1 SIGB_SMALL_TILE_CTAS=6" "SIGB_TILE_CLASS=1 SIGB_SMALL_TILE_CTAS=5" "SIGB_TILE_CLASS=1 SIGB_SMALL_TILE_CTAS=7" \
           "SIGB_TILE_CLASS=1 SIGB_LIB_VARIANT=_s512 SIGB_SMALL_TILE_CTAS=8" "SIGB_TILE_CLASS=1 SIGB_LIB_VARIANT=_s512 SIGB_SMALL_TILE_CTAS=10"; do
  run "$cfg" "--kind poisson --n 16777216"
  run "$cfg" "--kind poisson --n 16777216 --dot"
done
date | tee -a $S
