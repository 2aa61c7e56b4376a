#!/usr/bin/env python
"""Per-kernel share of ONE training step out of an ncu launch list of the default bench command: the launches between the last two
adam_kernel launches (a graph-replayed step).  python scripts/step_share.py profiles/r1_launches_final_bench.csv [ms_per_step]"""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
r = csv.reader(lines)
hdr = next(r)
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
rows = [(row[ki], float(row[vi].replace(",", "")) / 1000) for row in r]
adam = [i for i, (k, _) in enumerate(rows) if "adam_kernel" in k]
lo, hi = adam[-2] + 1, adam[-1] + 1
agg = collections.OrderedDict()
for k, us in rows[lo:hi]:
    agg.setdefault(k.split("(")[0].replace("nrf::", "").replace("void ", "")[-44:], []).append(us)
tot = sum(sum(v) for v in agg.values())
print(f"# One training step (graph replay) out of {sys.argv[1]}: the launches between the last two adam_kernel launches.")
print("# ncu per-launch times are cold-cache and serialised: compare SHARES." + (f"  bench.py (same command, no profiler): {sys.argv[2]} ms/step;" if len(sys.argv) > 2 else "")
      + f" sum of the launches below: {tot:.1f} us")
print(f"{'kernel':44s} {'n':>3s} {'us (sum)':>10s} {'share':>7s}")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:44s} {len(v):3d} {sum(v):10.1f} {sum(v) / tot:7.1%}")
