#!/usr/bin/env python3
"""Prints the handful of metrics we track from an .ncu-rep (per profiled launch)."""
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
pat = re.compile(r"^(Kernel Name|gpu__time_duration\.sum|dram__bytes_read\.sum|dram__bytes_write\.sum|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|"
                 r"launch__registers_per_thread|launch__grid_size|launch__block_size|smsp__inst_executed\.sum|smsp__issue_active\.avg\.pct_of_peak_sustained_active|"
                 r"sm__inst_executed_pipe_alu\.avg\.pct_of_peak_sustained_active|sm__inst_executed_pipe_fma\.avg\.pct_of_peak_sustained_active|"
                 r"sm__pipe_fma_cycles_active\.avg\.pct_of_peak_sustained_active|sm__pipe_alu_cycles_active\.avg\.pct_of_peak_sustained_active|"
                 r"smsp__warps_active\.avg\.per_cycle_active|smsp__warps_eligible\.avg\.per_cycle_active|sm__warps_active\.avg\.pct_of_peak_sustained_active|"
                 r"smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio|l1tex__t_sector_hit_rate\.pct|lts__t_sector_hit_rate\.pct|"
                 r"smsp__inst_executed_op_local_(ld|st)\.sum|smsp__thread_inst_executed_per_inst_executed\.ratio|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|sm__throughput\.avg\.pct_of_peak_sustained_elapsed)$")
for r in rows[2:]:
    print("----")
    for i, h in enumerate(hdr):
        if pat.match(h):
            v = r[i]
            if "stalled" in h:
                try:
                    if float(v) < 0.3:
                        continue
                except ValueError:
                    pass
            print("%-90s %-14s %s" % (h, units[i], v))
