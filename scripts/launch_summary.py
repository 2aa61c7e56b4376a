#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count / mean / min / total (us)."""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
r = csv.reader(lines)
hdr = next(r)
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg = collections.OrderedDict()
for n, row in enumerate(r):
    if n < skip:
        continue
    agg.setdefault(row[ki].split("(")[0][-48:], []).append(float(row[vi].replace(",", "")) / 1000)
tot = sum(sum(v) for v in agg.values())
print(f"{'kernel':50s} {'n':>4s} {'mean us':>9s} {'min us':>9s} {'total us':>10s} {'share':>6s}")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:50s} {len(v):4d} {sum(v)/len(v):9.1f} {min(v):9.1f} {sum(v):10.1f} {sum(v)/tot:6.1%}")
print(f"{'total':50s} {'':4s} {'':9s} {'':9s} {tot:10.1f}")
