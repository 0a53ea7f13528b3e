"""Development aid: host-side phase times (FXG_PROFILE) of one large job (config 3 / config 4 sample) through fxg_verify_reads."""
import os, sys, time
os.environ["FXG_PROFILE"] = "1"
sys.path.insert(0, os.getcwd())
from floxer_b200 import gpu as g, workloads as W
from floxer_b200.batch import VerifyConfig
name, n = sys.argv[1], int(sys.argv[2])
refs, table = W.build_references(name)
batch = W.build_reads(name, refs, table, g.pex_build, n_reads=n)
ctx = g.Context(0); ctx.set_references(refs)
for ivopt in (False, True):
    cfg = VerifyConfig(interval_optimization=ivopt)
    for it in range(3):
        print(f"---- {name} ivopt={ivopt} run {it}", file=sys.stderr, flush=True)
        t0 = time.perf_counter(); j = ctx.verify_reads(batch, cfg); dt = time.perf_counter() - t0
        print("verify_reads ms", dt * 1e3, file=sys.stderr, flush=True)
        j.free()
    st = ctx.stage_verify(batch, cfg)
    for it in range(3):
        print(f"---- staged {name} ivopt={ivopt} run {it}", file=sys.stderr, flush=True)
        t0 = time.perf_counter(); st.run(); dt = time.perf_counter() - t0
        print("staged run ms", dt * 1e3, file=sys.stderr, flush=True)
    st.free()
