#!/usr/bin/env python3
"""Summarise an `ncu --set full` report: per captured kernel the metrics the roofline discussion uses.
usage: tools/ncu_summary.py report.ncu-rep > profiles/xxx.txt"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.max",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
    "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__cycles_active.avg", "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
name_i = hdr.index("Kernel Name")
seen = set()
for r in rows[2:]:
    k = r[name_i]
    if k in seen:
        continue
    seen.add(k)
    print("==== " + k[:110])
    for m in WANT + [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]:
        if m in hdr:
            v = r[hdr.index(m)]
            try:
                if m.startswith("smsp__average_warps") and float(v.replace(",", "")) < 0.5:
                    continue
            except ValueError:
                pass
            print(f"  {m:90s} {v:>16s} {units[hdr.index(m)]}")
