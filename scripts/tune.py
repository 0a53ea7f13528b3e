"""Development aid: times fxg_verify_run on the bench workload under the environment's FXG_* knobs."""
import sys, os, time
sys.path.insert(0, os.getcwd())
import bench
from floxer_b200 import gpu as g
from floxer_b200.batch import VerifyConfig
refs, batch, _ = bench.build_workload("config2", 0, g.pex_build, None, 8)
ctx = g.Context(0); ctx.set_references(refs)
job = ctx.stage_verify(batch, VerifyConfig())
for _ in range(3): job.run()
ts = []
for _ in range(8):
    t0 = time.perf_counter(); job.run(); ts.append((time.perf_counter() - t0) * 1e3)
print({k: os.environ[k] for k in os.environ if k.startswith("FXG_")}, "run ms min %.2f med %.2f" % (min(ts), sorted(ts)[len(ts)//2]))
