#!/usr/bin/env python
"""Prints kernel name, grid and duration from an `ncu --csv` launch list (stdin or a file):
    ncu --metrics gpu__time_duration.sum --clock-control none --csv python tools/profile_bins.py --bins 3 | python tools/ncu_times.py
"""
import csv
import sys

src = open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin
rows = [r for r in csv.reader(l for l in src if l.startswith('"'))]
hdr = rows[0]
name, grid, val, metric = (hdr.index(k) for k in ("Kernel Name", "Grid Size", "Metric Value", "Metric Name"))
tot = {}
for r in rows[1:]:
    if r[metric] != "gpu__time_duration.sum":
        continue
    short = r[name].split("(")[0].replace("void ", "").split("::")[-1]
    us = float(r[val].replace(",", "")) / 1e3
    print("%-40s grid %-16s %8.1f us" % (short[:40], r[grid], us))
    a = tot.setdefault(short, [0, 0.0])
    a[0] += 1
    a[1] += us
print("---- totals")
for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-40s %4d launches %9.1f us  avg %7.1f us" % (k[:40], n, us, us / n))
