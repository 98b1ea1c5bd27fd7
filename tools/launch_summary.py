#!/usr/bin/env python3
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel launches, total and share.
usage: launch_summary.py launches.csv"""
import csv, sys, re, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[4]); name = re.sub(r"^void ", "", name)
    ns = float(r[14].replace(",", ""))
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += ns
tot = sum(a[1] for a in agg.values())
print("%-90s %8s %12s %7s" % ("kernel", "launches", "total_us", "share"))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-90s %8d %12.1f %6.1f%%" % (k[:90], a[0], a[1] / 1e3, 100 * a[1] / tot))
