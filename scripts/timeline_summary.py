"""Summary of a timeline written by scripts/timeline.py: device time per kernel, how many kernels run side by side,
and the gaps inside one batch's stream.  usage: timeline_summary.py <timeline.json.gz>"""
import collections, gzip, json, sys

rows = json.load(gzip.open(sys.argv[1]))
k = [r for r in rows if r["cat"] == "kernel"]
t0 = min(r["ts"] for r in k)
t1 = max(r["ts"] + r["dur"] for r in k)
span = t1 - t0
print(f"span {span / 1e3:.2f} ms, {len(k)} kernel launches on {len(set(r['stream'] for r in k))} streams, {len(rows) - len(k)} copies / memsets")
agg = collections.defaultdict(lambda: [0, 0.0, 0])
for r in k:
    a = agg[r["name"].split("(")[0][:60]]
    g = r["grid"] or [0, 0, 0]
    a[0] += 1; a[1] += r["dur"]; a[2] += g[0] * g[1] * g[2]
print(f"{'kernel':62s} {'launches':>8s} {'sum ms':>9s} {'avg us':>9s} {'avg grid':>9s}")
for n, a in sorted(agg.items(), key=lambda x: -x[1][1])[:16]:
    print(f"{n:62s} {a[0]:8d} {a[1] / 1e3:9.2f} {a[1] / a[0]:9.1f} {a[2] / a[0]:9.0f}")
print(f"sum of kernel durations {sum(r['dur'] for r in k) / 1e3:.1f} ms = {sum(r['dur'] for r in k) / span:.2f} x the span")
pts = []
for r in k:
    pts.append((r["ts"], 1)); pts.append((r["ts"] + r["dur"], -1))
pts.sort()
hist = collections.Counter(); cur = 0; last = pts[0][0]
for t, d in pts:
    hist[cur] += t - last; last = t; cur += d
print("kernels running side by side (fraction of the span): " + ", ".join(f"{c}: {hist[c] / span:.3f}" for c in sorted(hist)))
