#!/bin/bash
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests/test_gpu_scheduling.py tests/test_gpu_production_parity.py -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r2b_tests_new.log
(timeout 1500 python -m pytest tests/test_reference_python_overlay.py -m gpu -q 2>&1 | tail -40) > gpurun_out/r2b_tests_overlay.log
python scripts/probes/production_parity_probe.py C5 > gpurun_out/r2b_parity_probe_c5.log 2>&1
bash scripts/gpu_sweep.sh r2b scripts/sweeps/r2_schedule.spec > /dev/null 2>&1
(timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_scheduling.py --deselect tests/test_gpu_production_parity.py --deselect tests/test_reference_python_overlay.py 2>&1 | tail -15) > gpurun_out/r2b_tests_all.log
cat gpurun_out/r2b_tests_new.log gpurun_out/r2b_tests_overlay.log gpurun_out/r2b_parity_probe_c5.log gpurun_out/r2b_sweep.log gpurun_out/r2b_tests_all.log
