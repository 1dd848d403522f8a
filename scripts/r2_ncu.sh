#!/bin/bash
# Round 2 profiling visit (1 GPU): launch list of the bench step and ncu --set full captures of the
# dominant kernel in its default and row-direct forms (numbers under ncu are never bench values).
#   gpurun --timeout 1500 -- 'bash scripts/r2_ncu.sh r2p'
# then, here:  python scripts/ncu_summary.py gpurun_out/r2p/prof_csr.ncu-rep > profiles/r2_ncu_csr_tma.txt  (etc.)
TAG=${1:-r2p}; OUT=gpurun_out/$TAG; mkdir -p $OUT
S=$OUT/summary.txt
echo "== launch list of bench.py (shares of the step)" | tee $S
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu > $OUT/launches_bench.log 2>&1; echo "rc=$?" | tee -a $S
for rd in 0 1; do
  echo "== ncu full: csr_tma, SIGB_SPMV_ROWDIRECT=$rd" | tee -a $S
  SIGB_SPMV_ROWDIRECT=$rd timeout 900 ncu --set full --clock-control none --import-source on -k regex:csr_tma -s 20 -c 3 -f \
      -o $OUT/prof_csr_rd$rd python bench.py --steps 48 --warmup 3 --no-cpu > $OUT/ncu_csr_rd$rd.log 2>&1; echo "rc=$?" | tee -a $S
done
echo "== ncu full: CG vector kernels" | tee -a $S
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"CgUpdateOp|CgDirectionOp" -s 34 -c 2 -f \
    -o $OUT/prof_ew python bench.py --steps 48 --warmup 3 --no-cpu > $OUT/ncu_ew.log 2>&1; echo "rc=$?" | tee -a $S
echo "== ncu full: persistent CG kernel at the 8-GPU shard size (one launch = the whole solve)" | tee -a $S
SIGB_CG_PERSISTENT=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:cg_persistent -c 1 -f \
    -o $OUT/prof_persistent python bench.py --grid 1448 --steps 48 --warmup 3 --quick > $OUT/ncu_persistent.log 2>&1; echo "rc=$?" | tee -a $S
ls -la $OUT | tee -a $S
