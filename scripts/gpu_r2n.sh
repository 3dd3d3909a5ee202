#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_scheduling.py -m gpu -q -k literals 2>&1 | tail -5) > gpurun_out/r2n_tests.log
bash scripts/gpu_sweep.sh r2n scripts/sweeps/r2_imm_hoist.spec > /dev/null 2>&1
cat gpurun_out/r2n_tests.log gpurun_out/r2n_sweep.log; tail -3 gpurun_out/r2n_err.log
