#!/bin/bash
# round-2 first call: production-tier parity statistics + both tiers' throughput + baselines of C3/C4
mkdir -p gpurun_out
python scripts/probes/production_parity_probe.py C2 C3 --k 2048 > gpurun_out/r2a_parity_probe.log 2>&1
python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2a_bench_c2.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --bit-exact 1 2>&1 | tail -1 > gpurun_out/r2a_bench_c2_bitexact.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --shuffle 1 2>&1 | tail -1 > gpurun_out/r2a_bench_c2_shuffled.log
cat gpurun_out/r2a_parity_probe.log
for f in c2 c2_bitexact c2_shuffled; do cut -c1-300 gpurun_out/r2a_bench_$f.log; done
