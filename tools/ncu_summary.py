#!/usr/bin/env python3
"""Summarise one kernel of an .ncu-rep (raw page) into the few numbers DESIGN.md / profiles/ quote."""
import csv
import subprocess
import sys

KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_lsu.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def main(path, which=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    sel = [r for r in rows[2:] if which is None or which in r[hdr.index("Kernel Name")]]
    for vals in (sel if which else sel[-1:]):
        one(hdr, units, vals)
        print()


def one(hdr, units, vals):
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:72s} {vals[i]} {units[i]}")
    print("\nwarp stall reasons (average warps stalled per issue-active cycle):")
    st = []
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                st.append((float(vals[i]), h))
            except ValueError:
                pass
    for v, h in sorted(st, reverse=True):
        if v > 0.05:
            print(f"  {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):40s} {v:6.2f}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
