#!/bin/bash
# A/B sweeps on the GPU box (from the repo root): bash scripts/gpu_sweep.sh <tag> ; results in gpurun_out/<tag>_sweep.log
tag=${1:-sweep}
mkdir -p gpurun_out
out=gpurun_out/${tag}_sweep.log
: > $out
run() { # label, env, args...
  label=$1; envs=$2; shift 2
  line=$(env $envs python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" 2>gpurun_out/${tag}_err.log | tail -1)
  echo "$label | $(echo "$line" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); k=d["config"].get("kernel",{}); print(d["ms_per_step"], d["value"], k.get("registers"), k.get("local_bytes"), k.get("blocks_per_sm"), d.get("roofline",{}).get("frac"))' 2>&1 | tail -1)" >> $out
}
run "C3 default" "A=1" --workload C3
run "C3 extents in shared memory" "CLODE_EXT_SMEM=1" --workload C3
run "C3 extents in shared memory, 5 blocks/SM" "CLODE_EXT_SMEM=1" --workload C3 --min-blocks 5
run "C3 extents in shared memory, 3 blocks/SM" "CLODE_EXT_SMEM=1" --workload C3 --min-blocks 3
run "C4 default" "A=1" --workload C4
run "C4 extents in shared memory" "CLODE_EXT_SMEM=1" --workload C4
run "C4 extents in shared memory, 5 blocks/SM" "CLODE_EXT_SMEM=1" --workload C4 --min-blocks 5
cat $out
