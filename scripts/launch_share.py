#!/usr/bin/env python
"""Per-kernel share of the device time from an ncu launch list:
   ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv <command>
usage: python scripts/launch_share.py launches.csv > profiles/NAME.txt"""
import csv
import re
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[start]
ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
ui = hdr.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[start + 1:]:
    if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[ki]).strip()
    v = float(r[vi].replace(",", ""))
    unit = r[ui]
    ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}.get(unit, 1e-6)
    tot[name] += ms
    cnt[name] += 1
total = sum(tot.values())
print(f"{'kernel':60s} {'launches':>9s} {'ms':>10s} {'share':>7s}")
for k in sorted(tot, key=lambda k: -tot[k]):
    print(f"{k[:60]:60s} {cnt[k]:9d} {tot[k]:10.3f} {100 * tot[k] / total:6.1f}%")
print(f"{'total':60s} {sum(cnt.values()):9d} {total:10.3f}")
