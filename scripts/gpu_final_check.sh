#!/bin/bash
# exactly what the driver runs at round end, on the final snapshot
mkdir -p gpurun_out
SECONDS=0
(python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4) > gpurun_out/final_tests.log
echo "pytest -m gpu wall seconds: $SECONDS" >> gpurun_out/final_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
SECONDS=0
(python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 2>&1 | tail -1) > gpurun_out/final_bench_reference.json
echo "reference arm wall seconds: $SECONDS" >> gpurun_out/final_tests.log
SECONDS=0
(python bench.py --gpus 1 2>gpurun_out/final_bench_err.log | tail -1) > gpurun_out/final_bench.json
echo "bench.py (defaults) wall seconds: $SECONDS" >> gpurun_out/final_tests.log
cat gpurun_out/final_tests.log; tail -1 gpurun_out/final_smoke.log; cut -c1-300 gpurun_out/final_bench_reference.json; cut -c1-400 gpurun_out/final_bench.json; tail -2 gpurun_out/final_bench_err.log
