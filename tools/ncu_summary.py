#!/usr/bin/env python
"""tools/ncu_summary.py <report.ncu-rep> — the handful of ncu metrics DESIGN.md / profiles/ quote."""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_active.avg", "sm__inst_executed_pipe_fp64.sum",
        "smsp__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.sum",
        "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fmaheavy.sum", "sm__inst_executed_pipe_uniform.sum",
        "sm__inst_executed_pipe_cbu.sum", "sm__inst_executed_pipe_adu.sum"]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    print("kernels:", [r[name_i][:40] for r in data])
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print("%-80s %s  [%s]" % (w, " | ".join(r[i] for r in data), units[i]))
    print("-- stall reasons (warps per issue-active) --")
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            print("%-28s %s" % (h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")],
                                " | ".join("%.2f" % float(r[i] or 0) for r in data)))


if __name__ == "__main__":
    main()
