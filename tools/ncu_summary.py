#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into the handful of metrics the roofline discussion uses.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_static",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    kn = hdr.index("Kernel Name")
    print("# %s" % path)
    for r in data:
        print("\n== %s" % r[kn][:100])
        rd = wr = None
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print("%-82s %s %s" % (w, r[i], units[i]))
                if w == "dram__bytes_read.sum": rd = (float(r[i]), units[i])
                if w == "dram__bytes_write.sum": wr = (float(r[i]), units[i])
        if rd and wr:
            sc = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            print("%-82s %.0f byte" % ("dram traffic (read+write) per launch", rd[0] * sc[rd[1]] + wr[0] * sc[wr[1]]))


if __name__ == "__main__":
    main(sys.argv[1])
