#!/usr/bin/env python
"""`.ncu-rep` -> JSON summary (one object per captured launch) with the metrics the profile notes quote.
Usage: python tools/ncu_to_json.py gpurun_out/x.ncu-rep > profiles/x_summary.json"""
import csv
import json
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__stack_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fmaheavy.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d, u = dict(zip(hdr, vals)), dict(zip(hdr, units))
        e = {"kernel": d.get("Kernel Name", "?").split("(")[0], "metrics": {}, "stalls_per_issue": {}}
        for h in hdr:
            if h in WANT:
                e["metrics"][h] = f"{d[h]} {u[h]}".strip()
            elif "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                try:
                    v = float(d[h] or 0)
                except ValueError:
                    continue
                if v >= 0.05:
                    e["stalls_per_issue"][h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")] = round(v, 3)
        res.append(e)
    json.dump({"source": sys.argv[1], "command": " ".join(sys.argv[2:]), "launches": res}, sys.stdout, indent=1)


main()
