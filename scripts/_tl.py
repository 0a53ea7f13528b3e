import sys, os, time
sys.path.insert(0, os.getcwd())
import bench
from floxer_b200 import gpu as g
from floxer_b200.batch import VerifyConfig
refs, batch = bench.make_workload("config2", 0, g.pex_build)
ctx = g.Context(0); ctx.set_references(refs)
job = ctx.stage_verify(batch, VerifyConfig())
for _ in range(3): job.run()
os.environ["FXG_TRACE_WAVES"] = "1"
t0=time.perf_counter(); job.run(); print("run ms", (time.perf_counter()-t0)*1e3, file=sys.stderr)
