"""Per-kernel share of device time from an `ncu --metrics gpu__time_duration.sum --csv` launch list
(usage: python scripts/launch_share.py launches.csv)."""
import csv
import re
import sys

lines = open(sys.argv[1]).read().splitlines()
start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
by = {}
total = 0.0
for r in csv.DictReader(lines[start:]):
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"].replace("void ", ""))
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    ns = v * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
    e = by.setdefault(name, [0, 0.0])
    e[0] += 1
    e[1] += ns
    total += ns
print(f"{'kernel':48s} {'launches':>9s} {'ms':>10s} {'share':>7s}")
for name, (n, ns) in sorted(by.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:48s} {n:9d} {ns / 1e6:10.3f} {ns / total:7.1%}")
print(f"{'total':48s} {sum(n for n, _ in by.values()):9d} {total / 1e6:10.3f}")
