#!/bin/bash
# A/B sweeps on the GPU box (from the repo root): bash scripts/gpu_sweep.sh <tag> ; results in gpurun_out/<tag>_sweep.log
tag=${1:-sweep}
mkdir -p gpurun_out
out=gpurun_out/${tag}_sweep.log
: > $out
run() { # label, env, args...
  label=$1; envs=$2; shift 2
  line=$(env $envs python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" 2>gpurun_out/${tag}_err.log | tail -1)
  echo "$label | $(echo "$line" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); k=d["config"].get("kernel",{}); print(d["ms_per_step"], d["value"], k.get("registers"), k.get("local_bytes"), k.get("blocks_per_sm"), d.get("roofline",{}).get("frac"))' 2>&1 | tail -1)" >> $out
}
for w in C3 C4 C5 C5e; do
  run "$w default (block order auto, exp table in shared memory, inline edge cases)" "A=1" --workload $w
  run "$w library-exp" "A=1" --workload $w --library-exp 1
  run "$w block order forward" "CLODE_BLOCK_ORDER=forward" --workload $w
done
run "C2 default (block order auto)" "A=1"
run "C2 block order forward" "CLODE_BLOCK_ORDER=forward"
run "C2 block order reverse" "CLODE_BLOCK_ORDER=reverse"
cat $out
