#!/bin/bash
# A/B sweeps on the GPU box (from the repo root): bash scripts/gpu_sweep.sh <tag> ; results in gpurun_out/<tag>_sweep.log
tag=${1:-sweep}
mkdir -p gpurun_out
out=gpurun_out/${tag}_sweep.log
: > $out
run() { # label, defines, args...
  label=$1; defs=$2; shift 2
  line=$(CLODE_EXTRA_DEFINES="$defs" python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" 2>gpurun_out/${tag}_err.log | tail -1)
  echo "$label | $(echo "$line" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); k=d["config"].get("kernel",{}); print(d["ms_per_step"], d["value"], k.get("registers"), k.get("local_bytes"), k.get("blocks_per_sm"), d.get("roofline",{}).get("frac"), "fwd", d["config"].get("forward_order_ms_per_step"))' 2>&1 | tail -1)" >> $out
}
for w in C2 C3; do
run "$w A: integer norm compares, integer step compares, one-step norm division" "" --workload $w
run "$w B: float norm compares, integer step compares, one-step norm division" "-DCLODE_FLOAT_NORM_COMPARES" --workload $w
run "$w C: integer norm compares, float step compares, one-step norm division" "-DCLODE_FLOAT_STEP_COMPARES" --workload $w
run "$w D: float compares, one-step norm division" "-DCLODE_FLOAT_NORM_COMPARES -DCLODE_FLOAT_STEP_COMPARES" --workload $w
run "$w E: float compares, two-step norm division (previous kernel)" "-DCLODE_FLOAT_NORM_COMPARES -DCLODE_FLOAT_STEP_COMPARES -DCLODE_TWO_STEP_NORM_DIVISION" --workload $w
done
cat $out
