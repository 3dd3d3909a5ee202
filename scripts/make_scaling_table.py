#!/usr/bin/env python
"""profiles/r02_scaling.md from the four bench lines profiles/r02_bench_n{1,2,4,8}.json (one 8-GPU box, back to back).
usage: python scripts/make_scaling_table.py [prefix]   (default prefix: profiles/r02_bench_n)"""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
prefix = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, "profiles", "r02_bench_n")
d = {n: json.loads(open(f"{prefix}{n}.json").read()) for n in (1, 2, 4, 8)}
out = ["# Scaling, round 2 — one 8xB200 box, the driver's launch line (`python -m torch.distributed.run --nproc-per-node N bench.py "
       "--gpus N --steps 5 --warmup 3`), all four runs back to back\n",
       "Lines: `profiles/r02_bench_n{1,2,4,8}.json` (complete JSON lines incl. tiers / parity / secondary / small_n).  ms = device time per "
       "pass, max over ranks; value = accepted instance-steps/s of the whole job.  Regenerate: `python scripts/make_scaling_table.py`.\n",
       "## C2 (headline): Lorenz dopri5 `basic`, f64\n",
       "| N | weak: 2^20 per GPU, ms | value | weak efficiency | e2e (C ABI, host buffers, async NCCL gather) | e2e efficiency | strong: 2^20 "
       "total, ms | strong efficiency | in-process front end (`FeatureSimulator(device_ids=0..N-1)`), ms per pass of N x 2^20 (upload + rest) | "
       "value | fraction of kernel-only |",
       "|---|---|---|---|---|---|---|---|---|---|---|"]
v1, e1, ms1 = d[1]["value"], d[1]["e2e"]["value"], d[1]["ms_per_step"]
for n in (1, 2, 4, 8):
    x = d[n]
    st, fe = x.get("strong"), x.get("e2e_frontend") or {}
    sp = fe.get("split", {})
    out.append(f"| {n} | {x['ms_per_step']:.2f} | {x['value']:.4g} | {x['value'] / (n * v1):.4f} | {x['e2e']['value']:.4g} | "
               f"{x['e2e']['value'] / (n * e1):.3f} | "
               + (f"{st['ms_per_step']:.2f} | {ms1 / (n * st['ms_per_step']):.3f}" if st else f"{ms1:.2f} | 1")
               + f" | {fe.get('ms_per_step', 0):.1f} ({sp.get('set_ensemble_ms', 0):.1f} + {sp.get('features_and_results_ms', 0):.1f}) | "
                 f"{fe.get('value', 0):.4g} | {fe.get('fraction_of_kernel_only', 0):.3f} |")
out += ["",
        "Strong scaling is bounded by the serial time loop of the dearest instance, not by the launch structure: a chaotic Lorenz instance takes "
        "≈ 6500 attempts of ≈ 1 µs each (one warp's dependent FP64 chain), i.e. ≈ 6.5 ms however many GPUs share the other 2^20 − 1 instances; "
        "at N = 8 the per-GPU share of the work is 7.3 ms.  The cost-sorted schedule starts those instances first.",
        "The in-process front end moves every byte through ONE Python process (N x 48 MB in, N x 48 MB out per pass).  Uploads: contiguous chunk "
        "h of the host matrix to GPU h (dense DMA, one host thread per GPU), then every GPU pulls its interleaved shard out of all chunks with "
        "peer loads over NVLink (`clode_scatter_records`); the first implementation (one strided host pass per shard) took 37 ms at N = 8 and "
        "gave 0.80 / 0.69 / 0.55 of kernel-only at N = 2 / 4 / 8.  Results: NVLink gather + transpose on GPU 0, one device-to-host copy "
        "(403 MB ≈ 11 ms at N = 8: the PCIe link of one GPU).  The configuration that scales is one process per GPU (round 1: e2e efficiency "
        "0.934 at N = 8).\n",
        "## The other BASELINE configs (weak scaling, `config.secondary` of the same runs)\n",
        "| config | " + " | ".join(f"N={n}: ms, steps/s" for n in (1, 2, 4, 8)) + " | efficiency at 8 |", "|---|---|---|---|---|---|"]
for k in d[1]["config"]["secondary"]:
    cells = [f"{d[n]['config']['secondary'][k]['ms_per_step']:.2f}, {d[n]['config']['secondary'][k]['value']:.4g}" for n in (1, 2, 4, 8)]
    eff = d[8]["config"]["secondary"][k]["value"] / (8 * d[1]["config"]["secondary"][k]["value"])
    w = d[1]["config"]["secondary"][k]["workload"]
    out.append(f"| {w.split(',')[0]} ({w.split(',')[1].strip()}) | " + " | ".join(cells) + f" | {eff:.4f} |")
t = d[1]["tiers"]["other_tier"]
out += ["", f"Bit-exact tier on C2 (N = 1): {t['ms_per_step']:.1f} ms, {t['value']:.4g} steps/s = {t['frac_of_nominal_fp64']:.3f} of nominal "
            f"FP64 (production: {d[1]['roofline']['frac_of_nominal']:.3f}).",
        "CPU reference arm (the reference's kernels as host C, OpenMP): 1.554e8 steps/s on 16 cores (N = 1 box), 3.05e8 on 32 cores "
        "(N = 2 arm on the 8-GPU box)."]
path = os.path.join(REPO, "profiles", "r02_scaling.md")
open(path, "w").write("\n".join(out) + "\n")
print(path)
