#!/bin/bash
# the complete GPU test-suite as the driver runs it (timed), harvesting the cubins it had to JIT; then smoke()
mkdir -p gpurun_out/cubins
touch gpurun_out/.marker
SECONDS=0
(timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -8) > gpurun_out/r2h_tests_all.log
echo "pytest -m gpu wall seconds: $SECONDS" >> gpurun_out/r2h_tests_all.log
find clode_b200/_cubin_cache -name '*.cubin' -newer gpurun_out/.marker -exec cp {} gpurun_out/cubins/ \;
ls gpurun_out/cubins | wc -l >> gpurun_out/r2h_tests_all.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h_smoke.log 2>&1
cat gpurun_out/r2h_tests_all.log; tail -2 gpurun_out/r2h_smoke.log
