#!/bin/bash
# usage: sweep_b.sh  (on the GPU box, from the repo root)
mkdir -p gpurun_out
out=gpurun_out/b_sweep.log
: > $out
run() { # label, env, args...
  label=$1; envs=$2; shift 2
  line=$(env $envs python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" 2>gpurun_out/b_err.log | tail -1)
  echo "$label | $(echo "$line" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); k=d["config"].get("kernel",{}); print(d["ms_per_step"], d["value"], k.get("registers"), k.get("local_bytes"), k.get("blocks_per_sm"), d.get("roofline",{}).get("frac"))' 2>&1 | tail -1)" >> $out
}
run "C2 default" "A=1"
run "C2 ieee-div(old route)" "A=1" --ieee-div 1
run "C3 default" "A=1" --workload C3
run "C3 ieee-div" "A=1" --workload C3 --ieee-div 1
run "C3 init=5" "CLODE_KERNEL_MIN_BLOCKS=0,5,0,0" --workload C3
run "C3 init=6" "CLODE_KERNEL_MIN_BLOCKS=0,6,0,0" --workload C3
run "C3 init=3" "CLODE_KERNEL_MIN_BLOCKS=0,3,0,0" --workload C3
run "C4 default" "A=1" --workload C4
run "C4 ieee-div" "A=1" --workload C4 --ieee-div 1
run "C5 default" "A=1" --workload C5
run "C5 ieee-div" "A=1" --workload C5 --ieee-div 1
cat $out
