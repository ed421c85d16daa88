"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` log: python tools/launch_list.py file.csv"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    k = r[ki][:110]
    v = float(r[vi].replace(",", ""))
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 14]:
    print(f"{a[0]:5d} x {a[1]/a[0]/1e3:10.1f} us = {a[1]/1e3:11.1f} us {100*a[1]/tot:5.1f}%  {k}")
