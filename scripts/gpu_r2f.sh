#!/bin/bash
# N=1: full bench line, launch list under ncu, full captures of the dominant launches
mkdir -p gpurun_out
(timeout 900 python bench.py --steps 10 --warmup 3 2>gpurun_out/r2f_bench_err.log | tail -1) > gpurun_out/r2f_bench_n1.json
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2f_bench_n1.json").read())
print("C2 ms", d["ms_per_step"], "first", d["config"]["first_call_ms"], "roof", d["roofline"]["frac"], d["roofline"]["frac_of_nominal"], "e2e", d["e2e"]["value"]/d["value"], "frontend", d.get("e2e_frontend"))
PY
# launch list of the default bench command (per-launch times under ncu are serialised, cold-cache)
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2f_launches.csv \
    python bench.py --steps 2 --warmup 3 --quick 1 --no-cpu-baseline > gpurun_out/r2f_under_ncu.log 2>&1
# full captures: the pilot launch and the first sorted round of the C2 features kernel (launches 2 and 3 of the 4th pass)
ncu --set full --clock-control none --import-source on -k clode_features -s 39 -c 2 -o gpurun_out/r2f_c2_features -f \
    python bench.py --steps 1 --warmup 3 --quick 1 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k clode_transient -s 13 -c 2 -o gpurun_out/r2f_c2t_transient -f \
    python bench.py --workload C2t --steps 1 --warmup 1 --quick 1 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k clode_features -s 2 -c 1 -o gpurun_out/r2f_c4_features -f \
    python bench.py --workload C4 --steps 1 --warmup 1 --quick 1 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k clode_features -s 2 -c 1 -o gpurun_out/r2f_c3_features -f \
    python bench.py --workload C3 --steps 1 --warmup 1 --quick 1 --no-cpu-baseline > /dev/null 2>&1
for f in gpurun_out/r2f_*.ncu-rep; do python scripts/ncu_summary.py $f 0 > ${f%.ncu-rep}_summary.txt 2>&1; python scripts/ncu_summary.py $f 1 >> ${f%.ncu-rep}_summary.txt 2>/dev/null; done
ls -la gpurun_out/r2f_*; head -5 gpurun_out/r2f_c2_features_summary.txt
