"""Development aid: the figures of a bench.py JSON line as a table.  usage: bench_table.py line.json"""
import json, sys
d = json.load(open(sys.argv[1]))
def row(k, v):
    p = v.get("e2e_pageable")
    print(f"{k:20s} reads/s {v['reads_per_s']:10.0f}  GCUPS {v['value']:9.0f}  frac {v['roofline']['frac']:.3f}  dominant launch {v['roofline']['dominant_launch']['frac']:.3f}"
          f"  e2e {v['e2e']['reads_per_s']:9.0f}" + (f" (pageable {p['reads_per_s']:.0f})" if p else "") +
          f"  cpu {v['cpu_baseline']['reads_per_s']:8.1f} reads/s {v['cpu_baseline']['value']:6.1f} GCUPS on {v['cpu_baseline']['cores']} cores  ms/batch {v.get('ms_per_batch', 0):.3f}")
row("config2", d)
for k, v in d.get("sub", {}).items():
    if k != "config5":
        row(k, v)
print("clocks", d.get("clocks"), "n_gpus", d["n_gpus"], "ms_per_step", d["ms_per_step"], "host cores", d.get("host_cores"))
for c in d.get("sub", {}).get("config5", {}).get("cells", []):
    print(c["m"], c["error"], {kk: dict(wall_ms=round(c[kk]["ms"], 1), device_ms=round(c[kk]["device_ms"], 1), gcups=round(c[kk]["gcups"]), gcups_device=round(c[kk]["gcups_device"]),
                                        cpu_gcups=round(c[kk]["cpu_gcups"]), frac_device=round(c[kk]["roofline_frac_device"], 3)) for kk in ("exists", "cigar")})
