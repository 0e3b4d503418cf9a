#!/usr/bin/env python
"""Text summary of one kernel of an ncu report for profiles/: key metrics of the raw page + stall samples by CUDA line.
  python scripts/ncu_summary.py report.ncu-rep kernel_regex [top_lines] > profiles/xxx.txt"""
import csv
import os
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = sys.argv[3] if len(sys.argv) > 3 else "30"
KEEP = ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_read.sum", "launch__block_size", "launch__grid_size",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active")
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units = rows[0], rows[1]
print("== %s, kernels matching /%s/ (ncu --set full --clock-control none --import-source on)" % (os.path.basename(rep), kern))
for r in rows[2:]:
    for i, h in enumerate(hdr):
        if h in KEEP or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
            print("  %-78s %s %s" % (h, r[i], units[i]))
    print()
print("-- stall samples by CUDA source line (cuda,sass view; rows of a CUDA line include their SASS rows)")
sys.stdout.flush()
subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_lines.py"), rep, top, kern])
