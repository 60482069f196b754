#!/usr/bin/env python
"""Headline metrics per launch from an `ncu -i X.ncu-rep --page raw --csv` dump (the format of profiles/r2*_ncu_*.txt).
   python tools/ncu_brief.py raw.csv "<title line>" """
import csv
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__cluster_size", "smsp__cycles_active.avg"]
rows = list(csv.reader(open(sys.argv[1])))
hdr = units = None
print("# " + (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]))
n = 0
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr is not None and units is None:
        units = r
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    print(f"\nlaunch {n}: {d['Kernel Name'][:90]}")
    n += 1
    for k in KEYS:
        if k in d and d[k] != "":
            print(f"  {k} [{u.get(k, '')}] = {d[k]}")
