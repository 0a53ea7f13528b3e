#!/usr/bin/env python
"""Development probe: runs samples of configs 2-4 (floxer_b200/workloads.py) through the CUDA path, times them,
and compares the first reads bit for bit with the CPU port (oracle/cpu_baseline.c).

  python scripts/probe_configs.py [config2:1000 config3:2000 config4_shard:500 ...] [--check N] [--runs R]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("specs", nargs="*", default=["config2:1000", "config3:1000", "config4_shard:300"])
    ap.add_argument("--check", type=int, default=6, help="reads compared with the CPU port")
    ap.add_argument("--runs", type=int, default=3)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    from floxer_b200 import build, gpu as g, workloads as W
    from floxer_b200.batch import VerifyConfig, alignment_records
    from oracle import cpu_baseline
    build.build_native()
    results = []
    made = []
    for spec in args.specs:                      # all generation first: the read generator forks, and CUDA must not be up by then
        name, n = spec.split(":")
        t0 = time.time()
        refs, table = W.build_references(name)
        t1 = time.time()
        batch = W.build_reads(name, refs, table, g.pex_build, n_reads=int(n))
        made.append((name, int(n), refs, batch, t0, t1, time.time()))
    for name, n, refs, batch, t0, t1, t2 in made:
        t2b = time.time()
        ctx = g.Context(0)
        ctx.set_references(refs)
        t3 = time.time()
        rec = {"config": name, "reads": n, "anchors": int(len(batch.anchors)), "gen_ref_s": t1 - t0, "gen_reads_s": t2 - t1, "set_refs_s": t3 - t2b}
        for ivopt in (False, True):
            cfg = VerifyConfig(interval_optimization=ivopt)
            times = []
            ctx.reset_counters()
            for _ in range(args.runs):
                t = time.time()
                job = ctx.verify_reads(batch, cfg)
                times.append(time.time() - t)
                al, cg = job.alignments()
                stats = job.stats()
                job.free()
            ctr = ctx.counters()
            key = "ivopt" if ivopt else "plain"
            cells = stats["cells_inner"] + stats["cells_root"]
            rec[key] = {"s": times, "alignments": int(len(al)), "cigar_ops": int(len(cg)), "gcups_full_matrix": cells / min(times) / 1e9,
                        "reads_per_s": n / min(times), "word_steps_per_run": ctr["dp_word_steps"] / args.runs,
                        "dp_kernel_ms_per_run": ctr["dp_kernel_ms"] / args.runs, "trace_kernel_ms_per_run": ctr["trace_kernel_ms"] / args.runs,
                        "launches_per_run": ctr["kernel_launches"] / args.runs, "stats": stats,
                        "shared_score_passes": ctr["shared_score_passes"] / args.runs, "rescored": ctr["rescored_roots"] / args.runs,
                        "inferred_inner": ctr["inferred_inner"] / args.runs, "shared_tracebacks": ctr["shared_tracebacks"] / args.runs}
            if args.check:
                sub = batch.slice(0, min(args.check, n))
                job = ctx.verify_reads(sub, cfg)
                a1, c1 = job.alignments()
                s1 = job.stats()
                job.free()
                t = time.time()
                a2, c2, s2 = cpu_baseline.verify_reads(refs, sub, cfg, threads=os.cpu_count() or 1)
                rec[key]["cpu_s_for_check"] = time.time() - t
                same = alignment_records(a1, c1) == alignment_records(a2, c2) and s1 == s2
                rec[key]["parity_with_cpu_port"] = bool(same)
                rec[key]["cpu_reads_per_s"] = len(sub) / rec[key]["cpu_s_for_check"]
                if not same:
                    print("MISMATCH", name, key, len(a1), len(a2), s1, s2, file=sys.stderr)
        ctx.close()
        print(json.dumps(rec), flush=True)
        results.append(rec)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
