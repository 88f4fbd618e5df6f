#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch
list per kernel.  usage: launch_summary.py launches.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ik, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
per = collections.defaultdict(dict)
for r in data:
    if len(r) < len(hdr):
        continue
    per[int(r[iid])]["name"] = r[ik]
    per[int(r[iid])][r[im]] = float(r[iv].replace(",", ""))
agg = collections.OrderedDict()
for i, d in sorted(per.items()):
    n = d["name"].split("(")[0].replace("void ", "")
    a = agg.setdefault(n, [0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += d.get("gpu__time_duration.sum", 0)
    a[2] += d.get("dram__bytes_read.sum", 0)
    a[3] += d.get("dram__bytes_write.sum", 0)
tot = sum(a[1] for a in agg.values())
print("%-34s %8s %10s %8s %12s %12s" % ("kernel", "launches", "avg us", "share", "dram rd MB", "dram wr MB"))
for n, a in agg.items():
    print("%-34s %8d %10.1f %7.1f%% %12.2f %12.2f" % (n[:34], a[0], a[1] / a[0] / 1e3, 100 * a[1] / tot,
                                                      a[2] / a[0] / 1e6, a[3] / a[0] / 1e6))
