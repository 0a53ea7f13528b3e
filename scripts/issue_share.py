#!/usr/bin/env python
"""Per-kernel issue-slot accounting from an ncu metrics pass:
   ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__inst_executed_pipe_alu.sum,sm__cycles_active.sum \
       --clock-control none --csv --log-file issue.csv <command>
usage: python scripts/issue_share.py issue.csv > profiles/NAME.txt
Columns: warp instructions executed (all pipes), warp instructions on the integer/logic pipe, SM-cycles with a resident
warp, device time -- summed over the launches of a kernel."""
import csv
import re
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[start]
idi, ki, mi, vi, ui = (hdr.index(x) for x in ("ID", "Kernel Name", "Metric Name", "Metric Value", "Metric Unit"))
scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}
acc = defaultdict(lambda: defaultdict(float))
launches = defaultdict(set)
for r in rows[start + 1:]:
    if len(r) <= vi:
        continue
    name = re.sub(r"\(.*", "", r[ki]).strip()
    v = float(r[vi].replace(",", "") or 0)
    if r[mi] == "gpu__time_duration.sum":
        v *= scale.get(r[ui], 1e-6)
    acc[name][r[mi]] += v
    launches[name].add(r[idi])
tot = defaultdict(float)
for k in acc:
    for m, v in acc[k].items():
        tot[m] += v
print(f"{'kernel':44s} {'launches':>8s} {'ms':>9s} {'warp inst (M)':>14s} {'share':>6s} {'alu pipe (M)':>13s} {'share':>6s} {'SM-cycles (M)':>14s}")
for k in sorted(acc, key=lambda k: -acc[k]["smsp__inst_executed_pipe_alu.sum"]):
    a = acc[k]
    print(f"{k[:44]:44s} {len(launches[k]):8d} {a['gpu__time_duration.sum']:9.3f} {a['smsp__inst_executed.sum'] / 1e6:14.2f} "
          f"{100 * a['smsp__inst_executed.sum'] / max(tot['smsp__inst_executed.sum'], 1):5.1f}% {a['smsp__inst_executed_pipe_alu.sum'] / 1e6:13.2f} "
          f"{100 * a['smsp__inst_executed_pipe_alu.sum'] / max(tot['smsp__inst_executed_pipe_alu.sum'], 1):5.1f}% {a['sm__cycles_active.sum'] / 1e6:14.2f}")
print(f"{'total':44s} {sum(len(v) for v in launches.values()):8d} {tot['gpu__time_duration.sum']:9.3f} {tot['smsp__inst_executed.sum'] / 1e6:14.2f} {'':6s} "
      f"{tot['smsp__inst_executed_pipe_alu.sum'] / 1e6:13.2f} {'':6s} {tot['sm__cycles_active.sum'] / 1e6:14.2f}")
