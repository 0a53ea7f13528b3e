"""Development aid: host phases (FXG_PROFILE) of one fxg_align_batch_run over 2^20 small tasks (config 5's m = 100 cell)."""
import os, sys, time
os.environ["FXG_PROFILE"] = "1"
sys.path.insert(0, os.getcwd())
import numpy as np
from floxer_b200 import abi, gpu as g, synthetic
m = int(sys.argv[1]) if len(sys.argv) > 1 else 100
ref = synthetic.random_reference(10_000_000, 20240006)
ctx = g.Context(0); ctx.set_references([ref])
base, pool = synthetic.microbench_tasks(ref, [m], [0.05], 1 << 12, 20240011, abi.MODE_EXISTS)
for mode, name in ((abi.MODE_EXISTS, "exists"), (abi.MODE_CIGAR, "cigar")):
    tasks = np.tile(base, (1 << 20) // len(base)); tasks["mode"] = mode
    b = ctx.stage_align_batch(tasks, pool)
    for it in range(3):
        print(f"---- {name} run {it}", file=sys.stderr, flush=True)
        t0 = time.perf_counter(); b.run(); print(f"run ms {(time.perf_counter() - t0) * 1e3:.2f}", file=sys.stderr, flush=True)
    b.free()
