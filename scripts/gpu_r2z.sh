#!/bin/bash
# round 2, final call of the third session: the driver's round-end sequence on the final code, the launch list and the
# full ncu captures of the kernels that changed (branch-free right-hand sides), and the production-tier parity probe
mkdir -p gpurun_out
SECONDS=0
(timeout 900 python -m pytest tests/ -q -m gpu --durations=12 2>&1 | tail -30) > gpurun_out/r2z_tests.log
echo "pytest -m gpu wall seconds: $SECONDS" >> gpurun_out/r2z_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1
(timeout 300 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 2>&1 | tail -1) > gpurun_out/r2z_bench_reference.json
SECONDS=0
(timeout 900 python bench.py --gpus 1 2>gpurun_out/r2z_bench_err.log | tail -1) > gpurun_out/r2z_bench_n1.json
echo "bench.py (defaults) wall seconds: $SECONDS" >> gpurun_out/r2z_tests.log
tail -6 gpurun_out/r2z_tests.log; tail -1 gpurun_out/r2z_smoke.log; cut -c1-330 gpurun_out/r2z_bench_n1.json
# launch list of the quick bench command (per-launch times under ncu are serialised, cold-cache)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2z_launches.csv \
    python bench.py --steps 2 --warmup 3 --quick 1 --no-cpu-baseline > gpurun_out/r2z_under_ncu.log 2>&1
# full captures: C2 pilot + first sorted round; C3 features launch and first sorted warm-up round; C4; C5 rk4
timeout 300 ncu --set full --clock-control none --import-source on -k clode_features -s 39 -c 2 -o gpurun_out/r2z_c2_features -f \
    python bench.py --steps 1 --warmup 3 --quick 1 --no-cpu-baseline > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k clode_features -s 2 -c 1 -o gpurun_out/r2z_c3_features -f \
    python bench.py --workload C3 --steps 1 --warmup 1 --quick 1 --no-cpu-baseline > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k clode_initialize_observer -s 14 -c 1 -o gpurun_out/r2z_c3_warmup -f \
    python bench.py --workload C3 --steps 1 --warmup 1 --quick 1 --no-cpu-baseline > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k clode_features -s 2 -c 1 -o gpurun_out/r2z_c4_features -f \
    python bench.py --workload C4 --steps 1 --warmup 1 --quick 1 --no-cpu-baseline > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k clode_trajectory -s 2 -c 1 -o gpurun_out/r2z_c5_trajectory -f \
    python bench.py --workload C5 --steps 1 --warmup 1 --quick 1 --no-cpu-baseline > /dev/null 2>&1
for f in gpurun_out/r2z_*.ncu-rep; do python scripts/ncu_summary.py $f 0 > ${f%.ncu-rep}_summary.txt 2>&1; python scripts/ncu_summary.py $f 1 >> ${f%.ncu-rep}_summary.txt 2>/dev/null; done
find gpurun_out -name "r2z_*.ncu-rep" ! -name "r2z_c3_features.ncu-rep" -delete; ls -la gpurun_out/*.ncu-rep
(timeout 600 python scripts/probes/production_parity_probe.py C2 C3 C5 2>&1 | tail -40) > gpurun_out/r2z_production_parity_probe.log
for f in gpurun_out/r2z_*_summary.txt; do echo "== $f"; sed -n 2,4p $f; grep -E "fp64|issue_active|warps_active|local_ld|local_st|dram__bytes" $f | head -8; done
tail -12 gpurun_out/r2z_production_parity_probe.log
