#!/bin/bash
# round-2 ncu captures of the remaining kernels: C3 warm-up (first sorted round), C5 rk4 trajectory, C2 localmax, the bit-exact tier
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k clode_initialize_observer -s 14 -c 1 -o gpurun_out/r2k_c3_warmup -f \
    python bench.py --workload C3 --steps 1 --warmup 1 --quick 1 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k clode_trajectory -s 2 -c 1 -o gpurun_out/r2k_c5_trajectory -f \
    python bench.py --workload C5 --steps 1 --warmup 1 --quick 1 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k clode_features -s 40 -c 1 -o gpurun_out/r2k_c2l_features -f \
    python bench.py --workload C2l --steps 1 --warmup 1 --quick 1 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k clode_features -s 40 -c 1 -o gpurun_out/r2k_c2_bitexact_features -f \
    python bench.py --workload C2 --bit-exact 1 --steps 1 --warmup 1 --quick 1 --no-cpu-baseline > /dev/null 2>&1
for f in gpurun_out/r2k_*.ncu-rep; do python scripts/ncu_summary.py $f 0 > ${f%.ncu-rep}_summary.txt 2>&1; done
for f in gpurun_out/r2k_*_summary.txt; do echo "== $f"; sed -n 2,12p $f; sed -n 17,19p $f; sed -n 22,32p $f; done
