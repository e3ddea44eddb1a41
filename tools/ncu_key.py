#!/usr/bin/env python
"""Key metrics of the first kernel in an .ncu-rep (a reading aid for profiles/)."""
import csv
import subprocess
import sys

want = ['gpu__time_duration.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'launch__registers_per_thread', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed.avg.per_cycle_active',
        'launch__grid_size', 'launch__occupancy_limit', 'smsp__average_warp_latency_per_inst_issued.ratio',
        'smsp__average_warps_issue_stalled', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__waves_per_multiprocessor',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed_pipe_fp64', 'sm__inst_executed_pipe_fp64',
        'smsp__warps_eligible.avg.per_cycle_active', 'launch__shared_mem_per_block_dynamic']
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
for i, h in enumerate(hdr):
    if any(h.startswith(w) for w in want):
        print(f"{h:90s} {vals[i]:>16s} {units[i]}")
