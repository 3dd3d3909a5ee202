#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_host_boundary.py tests/test_gpu_frontend.py -m gpu -x -q 2>&1 | tail -25) > gpurun_out/r2c_tests_host.log
(timeout 900 python bench.py --steps 5 --warmup 3 2>gpurun_out/r2c_bench_err.log | tail -1) > gpurun_out/r2c_bench_full.json
bash scripts/gpu_sweep.sh r2c scripts/sweeps/r2_schedule2.spec > /dev/null 2>&1
(timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1) > gpurun_out/r2c_bench_reference.json
cat gpurun_out/r2c_tests_host.log; tail -5 gpurun_out/r2c_bench_err.log; cut -c1-6000 gpurun_out/r2c_bench_full.json; cat gpurun_out/r2c_sweep.log; cut -c1-600 gpurun_out/r2c_bench_reference.json
