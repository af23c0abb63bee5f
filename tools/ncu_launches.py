#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv --log-file X` launch list per kernel.
usage: ncu_launches.py launches.csv ["header line"]"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
k, v, u = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows:
    if r is hdr or r[k] == "Kernel Name":
        continue
    try:
        t = float(r[v].replace(",", ""))
    except ValueError:
        continue
    t = t / 1e3 if r[u] in ("ns", "nsecond") else (t * 1e3 if r[u] in ("ms", "msecond") else t)  # -> us
    a = agg.setdefault(r[k], [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(a[1] for a in agg.values())
if len(sys.argv) > 2:
    print(sys.argv[2])
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:60]:60s} launches {n:4d} total {t / 1e3:9.3f} ms share {100 * t / tot:5.1f}%  avg {t / n:9.1f} us")
