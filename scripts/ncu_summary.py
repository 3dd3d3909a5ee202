"""Summarise an .ncu-rep (first kernel) into the handful of numbers DESIGN.md / profiles/ quote."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0  # kernel index inside the report
vals = rows[2 + which]
d = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
print(f"# kernel {which} of {len(rows) - 2} in {rep}")
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__cycles_elapsed.max", "smsp__warps_eligible.avg.per_cycle_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum"]
for k in keys:
    if k in d:
        print(f"{k:75s} {d[k][0]:>18s} {d[k][1]}")
st = {h.split("stalled_")[1]: float(v[0] or 0) for h, v in d.items() if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h}
tot = sum(st.values()) or 1
print("warp stall reasons (pc sampling):")
for k, v in sorted(st.items(), key=lambda x: -x[1])[:8]:
    print(f"    {k:30s} {100 * v / tot:5.1f} %")
