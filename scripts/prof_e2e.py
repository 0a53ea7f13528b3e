import sys, os, time
sys.path.insert(0, os.getcwd())
import bench
from floxer_b200 import gpu as g
from floxer_b200.batch import VerifyConfig
refs, batch, _ = bench.build_workload("config2", 0, g.pex_build, None, 8)
ctx = g.Context(0); ctx.set_references(refs)
cfg = VerifyConfig()
for it in range(4):
    t0=time.perf_counter(); job = ctx.stage_verify(batch, cfg); t1=time.perf_counter()
    job.run(); t2=time.perf_counter()
    al, cg = job.alignments(); t3=time.perf_counter()
    job.free(); t4=time.perf_counter()
    print(f"stage {1e3*(t1-t0):.2f} run {1e3*(t2-t1):.2f} fetch {1e3*(t3-t2):.2f} free {1e3*(t4-t3):.2f} ms; {len(al)} alignments, {len(cg)} cigar ops")
