"""Development aid: config 2's job run a few times, alone, printing the word-steps the score passes issued per run
(for the ncu metrics pass behind scripts/issue_share.py)."""
import os, sys
sys.path.insert(0, os.getcwd())
import bench
from floxer_b200 import gpu as g
from floxer_b200.batch import VerifyConfig
runs = int(sys.argv[1]) if len(sys.argv) > 1 else 3
refs, batch, _ = bench.build_workload("config2", 0, g.pex_build, None, 8)
ctx = g.Context(0); ctx.set_references(refs)
job = ctx.stage_verify(batch, VerifyConfig())
prev = ctx.counters()
for it in range(runs):
    job.run()
    c = ctx.counters()
    print(f"run {it}: launches {c["kernel_launches"] - prev["kernel_launches"]} word-steps {c["dp_word_steps"] - prev["dp_word_steps"]} tasks {c["dp_tasks"] - prev["dp_tasks"]}", flush=True)
    prev = c
