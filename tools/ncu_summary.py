#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i … --page raw --csv`) into the text format kept under profiles/.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep "header line" > profiles/<round>_ncu_<what>_summary.txt
"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "smsp__branch_targets_threads_divergent",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def main():
    rep, header = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    if header:
        print(header)
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print("kernel:", d.get("Kernel Name"))
        for k in KEYS:
            if k in d:
                print("%-70s %s %s" % (k, d[k], u.get(k, "")))
        for k in hdr:
            if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
                print("  stall %-28s %s" % (k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], d[k]))
        print("---")


if __name__ == "__main__":
    main()
