#!/usr/bin/env python
"""Prints the handful of ncu metrics this repo's profile notes quote from a `.ncu-rep` (one block per captured launch).
Usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__stack_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        print("==", d.get("Kernel Name", "?")[:60])
        for h in hdr:
            if h in WANT or ("issue_stalled" in h and h.endswith("per_issue_active.ratio") and float(d[h] or 0) >= 0.05):
                print(f"  {h} = {d[h]} {u[h]}")


main()
