#!/usr/bin/env python
"""Per-kernel shares of ONE bench step from an ncu launch list (gpu__time_duration.sum, --csv).
usage: python profiles/launch_shares.py profiles/r01_launches.csv"""
import collections
import csv
import sys

lines = [ln for ln in open(sys.argv[1]) if ln.startswith('"')]
rows = list(csv.DictReader(lines))
ours = [(r["Kernel Name"].split("(")[0].replace("void ", "").replace("rtb::", "")[:60], r["Grid Size"], float(r["Metric Value"]))
        for r in rows if "at::" not in r["Kernel Name"]]
idx = [i for i, o in enumerate(ours) if "accum" in o[0]]
last = ours[idx[-2] + 1: idx[-1] + 1]  # everything between the last two sum launches = the last step
unit = 1e6 if max(o[2] for o in last) > 1e4 else 1e3  # ns or us -> ms
tot = sum(o[2] for o in last)
agg = collections.OrderedDict()
for o in last:
    a = agg.setdefault(o[0], [0, 0.0])
    a[0] += 1
    a[1] += o[2]
for k, v in agg.items():
    print(f"{k:60s} x{v[0]:3d} {v[1] / unit:8.3f} ms {100 * v[1] / tot:5.1f} %")
print(f"{'one step (serialised, cold cache)':60s}      {tot / unit:8.3f} ms")
