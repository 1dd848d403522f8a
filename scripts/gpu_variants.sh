#!/bin/bash
# A/B of experimental library builds on one GPU.  Usage: bash scripts/gpu_variants.sh <tag> "<variant list>"
TAG=${1:-ab}; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
for v in "$@"; do
  vv=$v; [ "$v" = "base" ] && vv=""
  SIGB_LIB_VARIANT=$vv timeout 300 python bench.py --steps 200 --warmup 5 --quick 2>> $OUT/err.log | tee -a $OUT/variants.jsonl
done
