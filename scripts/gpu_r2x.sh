#!/bin/bash
# round 2, session 3, second call: the GPU test-suite with the branch-free build as the default, the fast polar scale
# factor (CLODE_FAST_POLAR=1) A/B and its parity tests
mkdir -p gpurun_out
SECONDS=0
(timeout 600 python -m pytest tests/ -q -m gpu 2>&1 | tail -12) > gpurun_out/r2x_tests.log
echo "pytest -m gpu wall seconds: $SECONDS" >> gpurun_out/r2x_tests.log
SECONDS=0
(CLODE_FAST_POLAR=1 timeout 300 python -m pytest tests/test_gpu_production_parity.py tests/test_fast_exp.py -q -m gpu -k "c4 or polar" 2>&1 | tail -8) > gpurun_out/r2x_tests_polar.log
echo "polar tests wall seconds: $SECONDS" >> gpurun_out/r2x_tests_polar.log
SECONDS=0
timeout 600 bash scripts/gpu_sweep.sh r2x scripts/sweeps/r2_branchless2.spec > /dev/null 2>&1
echo "sweep wall seconds: $SECONDS" >> gpurun_out/r2x_sweep.log
cat gpurun_out/r2x_sweep.log; tail -4 gpurun_out/r2x_tests.log; tail -4 gpurun_out/r2x_tests_polar.log
