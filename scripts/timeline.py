"""Development aid: a device timeline of the config-2 step as bench.py drives it (LANES host threads, each running its
staged job), taken with torch.profiler (CUPTI activity records of every kernel and copy of the process, ours included).
Writes gpurun_out/<tag>_timeline.json.gz (kernel name, stream, start, duration, grid, block, registers, shared memory)
and prints how busy the device was.  usage: timeline.py <tag> [lanes] [steps] [flush: each|one|none]"""
import gzip, json, os, sys, time, threading
sys.path.insert(0, os.getcwd())
import bench
from floxer_b200 import gpu as g
from floxer_b200.batch import VerifyConfig

tag = sys.argv[1]
lanes = int(sys.argv[2]) if len(sys.argv) > 2 else 32
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
flush_mode = sys.argv[4] if len(sys.argv) > 4 else "each"
refs, batch, _ = bench.build_workload("config2", 0, g.pex_build, None, 8)
import torch
from torch.profiler import profile, ProfilerActivity
ctx = g.Context(0); ctx.set_references(refs)
jobs = [ctx.stage_verify(batch, VerifyConfig()) for _ in range(lanes)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def step(i):
    def f():
        if flush_mode == "each" or (flush_mode == "one" and i == 0):
            flush.fill_(1)
        jobs[i].run()
    return f


fns = [step(i) for i in range(lanes)]
for _ in range(4):
    bench.run_lanes(fns, 3)
torch.cuda.synchronize()
t0 = time.perf_counter(); dt, _ = bench.run_lanes(fns, steps); torch.cuda.synchronize()
print(f"unprofiled: {dt * 1e3 / steps:.2f} ms per step, {lanes * steps * len(batch) / dt:.0f} reads/s", flush=True)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    dt, _ = bench.run_lanes(fns, steps)
    torch.cuda.synchronize()
print(f"profiled: {dt * 1e3 / steps:.2f} ms per step", flush=True)
path = f"gpurun_out/{tag}_trace.json"
prof.export_chrome_trace(path)
ev = json.load(open(path))["traceEvents"]
os.remove(path)
rows = []
for e in ev:
    if e.get("ph") != "X" or e.get("cat") not in ("kernel", "gpu_memcpy", "gpu_memset"):
        continue
    a = e.get("args", {})
    rows.append({"name": e["name"][:80], "cat": e["cat"], "ts": e["ts"], "dur": e["dur"], "stream": a.get("stream"),
                 "grid": a.get("grid"), "block": a.get("block"), "regs": a.get("registers per thread"), "smem": a.get("shared memory"),
                 "occ": a.get("est. achieved occupancy %"), "bps": a.get("blocks per SM"), "wps": a.get("warps per SM"), "bytes": a.get("bytes")})
rows.sort(key=lambda r: r["ts"])
with gzip.open(f"gpurun_out/{tag}_timeline.json.gz", "wt") as f:
    json.dump(rows, f)
# busy fraction: union of kernel intervals, and the time-weighted number of kernels running
pts = []
for r in rows:
    if r["cat"] == "kernel":
        pts.append((r["ts"], 1)); pts.append((r["ts"] + r["dur"], -1))
pts.sort()
busy = 0.0; conc_t = 0.0; cur = 0; last = pts[0][0]
for t, d in pts:
    if cur > 0:
        busy += t - last
    conc_t += cur * (t - last)
    last = t; cur += d
span = pts[-1][0] - pts[0][0]
print(f"{len(rows)} records, span {span / 1e3:.2f} ms, some kernel running {busy / span:.3f} of it, mean kernels in flight {conc_t / span:.2f}")
