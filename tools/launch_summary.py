"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: python tools/launch_summary.py file.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 2:]
kn, mv, gs = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
agg = collections.OrderedDict()
for r in data:
    if len(r) <= mv:
        continue
    a = agg.setdefault(r[kn][:80], [0, 0.0, r[gs]])
    a[0] += 1
    a[1] += float(r[mv].replace(",", ""))
tot = sum(a[1] for a in agg.values())
for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{a[1]/1e3:10.1f} us total {a[0]:4d} launches {a[1]/a[0]/1e3:9.1f} us avg {100*a[1]/tot:5.1f}%  {n}  grid {a[2]}")
