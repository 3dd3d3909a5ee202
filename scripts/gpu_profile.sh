#!/bin/bash
# Final measurements + ncu captures on the GPU box (from the repo root): bash scripts/gpu_profile.sh <tag>
tag=${1:-prof}
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/${tag}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
python bench.py > gpurun_out/${tag}_bench_c2.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference.log 2>&1
for w in C3 C4 C5 C5e; do
  python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/${tag}_bench_$w.log
done
# launch list of the default bench command (per-launch times under ncu are serialised, cold-cache)
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_under_ncu.log 2>&1
# one full capture of the headline kernel (after the warm-up launches, so the block-order history is in place)
ncu --set full --clock-control none --import-source on -k clode_features -s 3 -c 1 -o gpurun_out/${tag}_c2_features -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
for f in gpurun_out/${tag}_*.ncu-rep; do python scripts/ncu_summary.py $f > ${f%.ncu-rep}_summary.txt 2>&1; done
cat gpurun_out/${tag}_tests.log; cat gpurun_out/${tag}_smoke.log | tail -2
tail -1 gpurun_out/${tag}_bench_c2.log | cut -c1-400
