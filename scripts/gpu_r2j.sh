#!/bin/bash
mkdir -p gpurun_out
touch gpurun_out/.marker
SECONDS=0
(timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -8) > gpurun_out/r2u_tests_all.log
echo "pytest -m gpu wall seconds: $SECONDS" >> gpurun_out/r2u_tests_all.log
(cd clode_b200/_cubin_cache && find . -name '*.cubin' -newer ../../gpurun_out/.marker -print0 | tar czf ../../gpurun_out/cubins.tgz --null -T -)
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2u_smoke.log 2>&1
(timeout 900 python bench.py --steps 10 --warmup 3 2>gpurun_out/r2u_bench_err.log | tail -1) > gpurun_out/r2u_bench_n1.json
cat gpurun_out/r2u_tests_all.log; tail -2 gpurun_out/r2u_smoke.log; tail -3 gpurun_out/r2u_bench_err.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2u_bench_n1.json").read())
print("C2 ms", d["ms_per_step"], "first", d["config"]["first_call_ms"], "roof", d["roofline"]["frac_of_nominal"], "e2e", d["e2e"]["value"]/d["value"])
print("tiers", d["tiers"]["other_tier"]["ms_per_step"], d["tiers"]["other_tier"]["frac_of_nominal_fp64"], "parity", d["parity"])
print("frontend", d["e2e_frontend"])
for k,v in d["config"]["secondary"].items(): print("  ", k, round(v["ms_per_step"],3), "%.4g"%v["value"])
PY
