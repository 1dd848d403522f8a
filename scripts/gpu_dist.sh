#!/bin/bash
# Multi-GPU visit: sharded parity tests + bench at N=1..NG.  Usage: bash scripts/gpu_dist.sh <tag> <ngpus> [steps]
TAG=${1:-r1d}; NG=${2:-2}; STEPS=${3:-200}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt 2>&1
nvidia-smi topo -m > $OUT/topo.txt 2>&1
echo "== pytest dist" | tee $OUT/summary.txt
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q > $OUT/pytest_dist.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -25 $OUT/pytest_dist.log | tee -a $OUT/summary.txt
for n in 1 2 4 8; do
  if [ $n -le $NG ]; then
    echo "== bench N=$n" | tee -a $OUT/summary.txt
    if [ $n -eq 1 ]; then
      timeout 600 python bench.py --gpus 1 --steps $STEPS --warmup 5 --no-cpu > $OUT/bench_n$n.json 2> $OUT/bench_n$n.err
    else
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus $n --steps $STEPS --warmup 5 > $OUT/bench_n$n.json 2> $OUT/bench_n$n.err
    fi
    echo "rc=$?" | tee -a $OUT/summary.txt
    cat $OUT/bench_n$n.json | tee -a $OUT/summary.txt; tail -8 $OUT/bench_n$n.err | tee -a $OUT/summary.txt
  fi
done
if [ "${NCU:-0}" = "1" ]; then
  echo "== ncu full: csr_tma (1 NDOT=0 + 2 NDOT=1 active launches) and CG vector kernels" | tee -a $OUT/summary.txt
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:csr_tma -s 20 -c 3 -f -o $OUT/prof_csr \
      python bench.py --steps 48 --warmup 3 --no-cpu > $OUT/ncu_csr.log 2>&1; echo "rc=$?" | tee -a $OUT/summary.txt
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"CgUpdateOp|CgDirectionOp" -s 40 -c 2 -f -o $OUT/prof_ew \
      python bench.py --steps 48 --warmup 3 --no-cpu > $OUT/ncu_ew.log 2>&1; echo "rc=$?" | tee -a $OUT/summary.txt
fi
