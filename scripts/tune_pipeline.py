"""Development aid: throughput of fxg_verify_run with several batches in flight (FXG_GROUPS / FXG_WORKERS knobs);
prints wall, user-CPU and system-CPU milliseconds per step."""
import sys, os, time, threading
sys.path.insert(0, os.getcwd())
import bench
from floxer_b200 import gpu as g
from floxer_b200.batch import VerifyConfig
depth = int(os.environ.get("FXG_GROUPS", "32"))
refs, batch, _ = bench.build_workload("config2", 0, g.pex_build, None, 8)
ctx = g.Context(0); ctx.set_references(refs)
jobs = [ctx.stage_verify(batch, VerifyConfig()) for _ in range(depth)]
def lanes(n):
    ts = [threading.Thread(target=lambda j=j: [j.run() for _ in range(n)]) for j in jobs]
    c0 = os.times(); t0 = time.perf_counter()
    for t in ts: t.start()
    for t in ts: t.join()
    dt = time.perf_counter() - t0; c1 = os.times()
    return round(dt * 1e3 / (n * depth), 2), round((c1.user - c0.user) * 1e3 / (n * depth), 1), round((c1.system - c0.system) * 1e3 / (n * depth), 1)
lanes(3)
print({k: os.environ[k] for k in os.environ if k.startswith("FXG_")}, "ms per step (wall, user, sys)", [lanes(6) for _ in range(3)])
