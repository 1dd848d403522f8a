#!/bin/bash
# Round 2, 2-GPU visit zz (charged 2x): lanczos / eigensolve on the single-process multi-GPU operator, the
# host mirror's csc / ellpack row forms after the refactoring (C++ program).
TAG=${1:-r2zz}; OUT=gpurun_out/$TAG; mkdir -p $OUT
S=$OUT/summary.txt
date | tee -a $S
timeout 600 python -m pytest tests/test_gpu_mgpu.py -x -q > $OUT/pytest.log 2>&1; echo "rc=$?" | tee -a $S
tail -4 $OUT/pytest.log | tee -a $S
timeout 300 tests/cxx/_build/solver_test_multi_gpu -v > $OUT/cxx_multi_gpu.log 2>&1; echo "solver_test_multi_gpu rc=$?" | tee -a $S
cat $OUT/cxx_multi_gpu.log | tee -a $S
date | tee -a $S
