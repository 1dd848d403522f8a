#!/bin/bash
# Round 2, 2-GPU visit z (charged 2x): eigensolve on row-sharded operators, csc / ellpack matrices sharded
# through their rows by the host mirror (C++ program), regression of the ILDU tests after the clean-up.
TAG=${1:-r2z}; OUT=gpurun_out/$TAG; mkdir -p $OUT
S=$OUT/summary.txt
date | tee -a $S
timeout 1200 python -m pytest tests/test_gpu_ldu.py tests/test_gpu_dist.py tests/test_gpu_mgpu.py tests/test_cxx_host.py tests/test_gpu_solvers.py -x -q -m gpu > $OUT/pytest.log 2>&1; echo "rc=$?" | tee -a $S
tail -5 $OUT/pytest.log | tee -a $S
timeout 300 tests/cxx/_build/solver_test_multi_gpu -v > $OUT/cxx_multi_gpu.log 2>&1; echo "solver_test_multi_gpu rc=$?" | tee -a $S
cat $OUT/cxx_multi_gpu.log | tee -a $S
date | tee -a $S
