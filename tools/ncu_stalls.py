#!/usr/bin/env python
"""Summarise an `ncu --set full --import-source on` report: per-kernel headline metrics, stall reasons and the
SASS instructions that collect the most warp-stall samples.
   python tools/ncu_stalls.py report.ncu-rep [kernel-regex] [top-N]"""
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else "."
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_active.avg", "lts__t_sector_hit_rate.pct",
        "launch__grid_size", "launch__block_size"]
for r in rows[2:]:
    d = dict(zip(h, r))
    if not re.search(pat, d["Kernel Name"]):
        continue
    print("==", d["ID"], d["Kernel Name"][:80])
    for k in KEYS:
        if k in d:
            print(f"   {k:70s} {d[k]}")

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
seen = set()
for b in blocks:
    if not re.search(pat, b["name"]) or b["name"] in seen:
        continue
    seen.add(b["name"])
    hdr = b["rows"][0]
    data = [r for r in b["rows"][1:] if len(r) == len(hdr)]
    si = [i for i, x in enumerate(hdr) if x.startswith("stall_") and "Not Issued" not in x]
    tot = sum(int(r[2]) for r in data)
    print("\n#### ", b["name"][:90], " samples:", tot, " instructions:", len(data))
    print("   ", ", ".join(f"{hdr[i][6:]}={sum(int(r[i] or 0) for r in data)}" for i in si))
    for k, r in sorted(enumerate(data), key=lambda kr: -int(kr[1][2]))[:topn]:
        top = sorted(((hdr[i][6:], int(r[i] or 0)) for i in si), key=lambda kv: -kv[1])[:2]
        print(f"   #{k:5d} {r[2]:>6s} exec={r[5]:>8s}  {r[1][:72]:72s} {top}")
