#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one optimiser step (from one
embed_gather_sum_fwd launch to the next), grouped by kernel.   python tools/launch_summary.py launches.csv"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = None
recs = []
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    d = dict(zip(hdr, r))
    try:
        v = float(d["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    unit = d.get("Metric Unit", "ns")
    v_us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    recs.append((d["Kernel Name"], d.get("Grid Size"), d.get("Block Size"), v_us))
starts = [i for i, r in enumerate(recs) if "embed_gather_sum_fwd" in r[0]]
if len(starts) >= 2:
    recs = recs[starts[-2]:starts[-1]]          # the last complete step (the first ones pack weights / create state)
agg = collections.OrderedDict()
for name, grid, block, us in recs:
    key = re.sub(r"\(.*", "", name)
    key = re.sub(r"^void ", "", key)[:70]
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
print(f"one step: {len(recs)} launches, {tot/1e3:.2f} ms of kernel time")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{t:9.1f} us {100*t/tot:5.1f}%  n={n:4d} avg={t/n:7.1f}  {k}")
