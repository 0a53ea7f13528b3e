"""Development aid: one staged job of a named config run a few times (for ncu launch lists and traces)."""
import os, sys, time
sys.path.insert(0, os.getcwd())
from floxer_b200 import gpu as g, workloads as W
from floxer_b200.batch import VerifyConfig
name, n, runs = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
ivopt = len(sys.argv) > 4 and sys.argv[4] == "ivopt"
refs, table = W.build_references(name)
batch = W.build_reads(name, refs, table, g.pex_build, n_reads=n)
ctx = g.Context(0); ctx.set_references(refs)
st = ctx.stage_verify(batch, VerifyConfig(interval_optimization=ivopt))
for it in range(runs):
    t0 = time.perf_counter(); st.run(); dt = time.perf_counter() - t0
    print(f"staged run {it}: {dt * 1e3:.2f} ms", file=sys.stderr, flush=True)
print(ctx.counters(), file=sys.stderr)
