import sys, os, time
sys.path.insert(0, os.getcwd())
import bench
from floxer_b200 import gpu as g
from floxer_b200.batch import VerifyConfig
refs, batch, _ = bench.build_workload("config2", 0, g.pex_build, None, 8)
ctx = g.Context(0); ctx.set_references(refs)
job = ctx.stage_verify(batch, VerifyConfig())
job.run(); 
os.environ["FXG_PROFILE"]="1"
t0=time.perf_counter(); job.run(); print("run ms", (time.perf_counter()-t0)*1e3)
print(ctx.counters())
