"""Development aid: host-side phase times (FXG_PROFILE) of config-2 batches on one worker, alone and merged."""
import os, sys, time, threading
os.environ["FXG_PROFILE"] = "1"
os.environ.setdefault("FXG_WORKERS", "1")
os.environ.setdefault("FXG_GROUPS", "1")
sys.path.insert(0, os.getcwd())
import bench
from floxer_b200 import gpu as g
from floxer_b200.batch import VerifyConfig
refs, batch, _ = bench.build_workload("config2", 0, g.pex_build, None, 8)
ctx = g.Context(0); ctx.set_references(refs)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
jobs = [ctx.stage_verify(batch, VerifyConfig()) for _ in range(n)]
for _ in range(2):
    jobs[0].run()
print("---- one job alone", file=sys.stderr, flush=True)
t0 = time.perf_counter(); jobs[0].run(); print("run ms", (time.perf_counter() - t0) * 1e3, file=sys.stderr, flush=True)
for rnd in range(3):
    print(f"---- {n} jobs at once, round {rnd}", file=sys.stderr, flush=True)
    ts = [threading.Thread(target=j.run) for j in jobs]
    t0 = time.perf_counter()
    for t in ts: t.start()
    for t in ts: t.join()
    print("all done ms", (time.perf_counter() - t0) * 1e3, file=sys.stderr, flush=True)
print(ctx.counters(), file=sys.stderr)
