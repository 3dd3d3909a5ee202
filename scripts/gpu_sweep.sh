#!/bin/bash
# A/B sweeps on the GPU box, from the repo root:
#   bash scripts/gpu_sweep.sh <tag> [spec file]
# Each line of the spec is  "label | ENV=VALUE ... | bench.py arguments";  without a spec the five workloads run with
# the default build.  One result line per run goes to gpurun_out/<tag>_sweep.log:
#   label | kernel ms/step, steps/s, registers, local bytes, blocks/SM, roofline fraction, forward-order ms/step
# The logs under profiles/r01_*_ab.log were produced this way (variants selected with --library-exp, --ieee-div,
# --min-blocks, CLODE_BLOCK_ORDER, CLODE_EXT_SMEM, CLODE_KERNEL_MIN_BLOCKS, CLODE_EXTRA_DEFINES).
tag=${1:-sweep}
spec=$2
mkdir -p gpurun_out
out=gpurun_out/${tag}_sweep.log
: > "$out"
run() { # label, "ENV=VALUE ...", bench arguments...
  label=$1; envs=$2; shift 2
  line=$(env $envs python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" 2>"gpurun_out/${tag}_err.log" | tail -1)
  echo "$label | $(echo "$line" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); k=d["config"].get("kernel",{}); print(d["ms_per_step"], d["value"], k.get("registers"), k.get("local_bytes"), k.get("blocks_per_sm"), d.get("roofline",{}).get("frac"), "first", d["config"].get("first_call_ms"), "e2e", d["e2e"]["value"])' 2>&1 | tail -1)" >> "$out"
}
if [ -n "$spec" ]; then
  while IFS='|' read -r label envs args; do
    [ -z "$label" ] && continue
    run "$(echo $label)" "A=1 $envs" $args
  done < "$spec"
else
  for w in C2 C2l C3 C4 C5 C5e; do run "$w default" "A=1" --workload $w; done
fi
cat "$out"
