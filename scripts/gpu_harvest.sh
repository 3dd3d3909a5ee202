#!/bin/bash
# run the complete GPU test-suite and bring back, as one tarball, every cubin it had to JIT (they then travel with the
# repo snapshot, so the driver's round-end run finds a warm cache)
mkdir -p gpurun_out
touch gpurun_out/.marker
SECONDS=0
(timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -6) > gpurun_out/harvest_tests.log
echo "pytest -m gpu wall seconds: $SECONDS" >> gpurun_out/harvest_tests.log
(cd clode_b200/_cubin_cache && find . -name '*.cubin' -newer ../../gpurun_out/.marker -print0 | tar czf ../../gpurun_out/cubins.tgz --null -T -)
ls -la gpurun_out/cubins.tgz; cat gpurun_out/harvest_tests.log
