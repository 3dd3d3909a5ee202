#!/bin/bash
# usage: bash scripts/gpu_multi.sh <N> <tag> [tests]   — the driver's own launch line for N GPUs
N=$1; tag=$2
mkdir -p gpurun_out
if [ -n "$3" ]; then
  (timeout 900 python -m pytest tests/test_gpu_frontend.py tests/test_gpu_host_boundary.py -m gpu -q 2>&1 | tail -15) > gpurun_out/${tag}_tests_multi.log
  cat gpurun_out/${tag}_tests_multi.log
fi
for n in $(echo $N | tr ',' ' '); do
  if [ "$n" = "1" ]; then
    (timeout 1200 python bench.py --gpus 1 --steps 5 --warmup 3 2>gpurun_out/${tag}_n${n}_err.log | tail -1) > gpurun_out/${tag}_bench_n${n}.json
  else
    (timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 5 --warmup 3 2>gpurun_out/${tag}_n${n}_err.log | tail -1) > gpurun_out/${tag}_bench_n${n}.json
  fi
  tail -3 gpurun_out/${tag}_n${n}_err.log | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench_n${n}.json").read())
    print("N=${n}", "ms", round(d["ms_per_step"],2), "value", "%.4g"%d["value"], "e2e", "%.4g"%d["e2e"]["value"], "strong", d.get("strong"), "frontend", {k:v for k,v in (d.get("e2e_frontend") or {}).items() if k in ("value","ms_per_step","kernel_ms_per_step","fraction_of_kernel_only","error")})
    for k,v in (d["config"].get("secondary") or {}).items(): print("   ", k, round(v["ms_per_step"],2), "%.4g"%v["value"])
except Exception as e:
    print("N=${n} no line:", e)
PY
done
if [ -n "$4" ]; then
  (timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>&1 | tail -1 | cut -c1-700) | tee gpurun_out/${tag}_reference_n2.json
fi
