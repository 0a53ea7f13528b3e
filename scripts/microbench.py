#!/usr/bin/env python
"""Config 5 of BASELINE.json: verification-only microbenchmark -- batched alignment::align, query 100..2000 bp against
a window of m + 2k + 1 bases, error rates 2..15 %, half of the queries true positives, modes `exists` and `cigar`.

Prints one line per (m, error rate, mode): tasks/s and GCUPS (m x n cells per task, full-matrix convention) of
fxg_align_batch_run with inputs resident in HBM.  Development / documentation aid; the numbers go to profiles/."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from floxer_b200 import abi, gpu as g, synthetic  # noqa: E402


def main():
    per_cell = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 15
    ref = synthetic.random_reference(10_000_000, 20240001)
    ctx = g.Context(0)
    ctx.set_references([ref])
    print(f"# {per_cell} tasks per cell, seed 20240006")
    for mode, name in ((abi.MODE_EXISTS, "exists"), (abi.MODE_CIGAR, "cigar")):
        for m in (100, 200, 500, 1000, 2000):
            for e in (0.02, 0.05, 0.10, 0.15):
                tasks, pool = synthetic.microbench_tasks(ref, [m], [e], per_cell, 20240006, mode)
                b = ctx.stage_align_batch(tasks, pool)
                b.run()
                ts = []
                for _ in range(3):
                    t0 = time.perf_counter()
                    b.run()
                    ts.append(time.perf_counter() - t0)
                res, _ = b.fetch()
                dt = min(ts)
                cells = float((tasks["ref_len"].astype(np.float64) * tasks["query_len"]).sum())
                print(f"mode={name:6s} m={m:5d} e={e:4.2f} k={int(tasks['max_errors'][0]):4d} n={int(tasks['ref_len'][0]):5d} "
                      f"hits={int(res['exists'].sum()):6d} {len(tasks) / dt / 1e6:8.3f} Mtasks/s {cells / dt / 1e9:10.1f} GCUPS  ({dt * 1e3:.2f} ms)")
                b.free()
    ctx.close()


if __name__ == "__main__":
    main()
