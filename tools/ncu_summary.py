#!/usr/bin/env python3
"""Summarise an .ncu-rep (ncu --set full) into the handful of metrics DESIGN.md / profiles/ quote.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [out.csv] [note]
"""
import csv
import io
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "sass__inst_executed_shared_loads", "sass__inst_executed_shared_stores",
    "smsp__sass_average_branch_targets_threads_uniform.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_fp64.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
]


def main():
    rep = sys.argv[1]
    out = sys.argv[2] if len(sys.argv) > 2 else None
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = ["# %s ; source %s" % (note, rep), "kernel,metric,unit,value"]
    for vals in rows[2:]:
        kname = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        for h, u, v in zip(hdr, units, vals):
            if h in KEEP or h.startswith("smsp__average_warps_issue_stalled"):
                lines.append("%s,%s,%s,%s" % (kname.split("(")[0][:60], h, u, v))
    txt = "\n".join(lines) + "\n"
    if out:
        open(out, "w").write(txt)
    sys.stdout.write(txt)


if __name__ == "__main__":
    main()
