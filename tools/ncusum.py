#!/usr/bin/env python
"""Summarise one kernel of an .ncu-rep into the few lines kept under profiles/:
   tools/ncusum.py gpurun_out/<tag>_smax_C3.ncu-rep "<header comment>" > profiles/<tag>_ncu_smax_c3_summary.txt"""
import csv
import io
import subprocess
import sys

KEEP = ["Kernel Name", "Block Size", "Grid Size", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "lts__t_sector_hit_rate.pct", "sm__cycles_active.avg", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.per_cycle_active", "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_uniform.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]

rep = sys.argv[1]
if len(sys.argv) > 2:
    print("# " + sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(out)))
names, units, vals = rows[0], rows[1], rows[2]
for k in KEEP:
    if k in names:
        i = names.index(k)
        print(f"{k:110s} {units[i]:18s} {vals[i]}")
