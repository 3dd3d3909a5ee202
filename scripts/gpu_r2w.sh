#!/bin/bash
# round 2, session 3: branch-free right-hand sides (CLODE_BRANCHLESS=1) and the controller trims, A/B per workload,
# plus the production-tier parity and accuracy tests with the branch-free build
mkdir -p gpurun_out
SECONDS=0
timeout 900 bash scripts/gpu_sweep.sh r2w scripts/sweeps/r2_branchless.spec > /dev/null 2>&1
echo "sweep wall seconds: $SECONDS" >> gpurun_out/r2w_sweep.log
SECONDS=0
(CLODE_BRANCHLESS=1 timeout 600 python -m pytest tests/test_fast_exp.py tests/test_gpu_production_parity.py tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -15) > gpurun_out/r2w_tests_branchless.log
echo "tests wall seconds: $SECONDS" >> gpurun_out/r2w_tests_branchless.log
cat gpurun_out/r2w_sweep.log; tail -5 gpurun_out/r2w_tests_branchless.log
