"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel and grid:
count, total time and share of the captured GPU time.
usage: python tools/launch_summary.py profiles/r01/launches_<wl>.csv"""
import collections
import csv
import sys

for path in sys.argv[1:]:
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        a = agg.setdefault((r[4][:90], r[8]), [0, 0.0, r[-2]])
        a[0] += 1
        a[1] += float(r[-1].replace(",", ""))
    tot = sum(v[1] for v in agg.values()) or 1.0
    print(f"== {path}: {len(rows)} launches")
    for (name, grid), v in agg.items():
        print(f"{v[0]:4d} x {name:90s} grid {grid:14s} total {v[1]:14.1f} {v[2]} share {v[1] / tot:.3f}")
