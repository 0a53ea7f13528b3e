"""Randomised differential test of the host queue's shortcuts (run on a GPU box: python scripts/stress_parity.py [cases] [seed] [only-case|-] [oracle-every]).

Every case is a random batch -- several references (some barely longer than a read, so that windows are clipped at both
ends), reads of mixed lengths (different trees in one batch), random error rates / seed errors / tree builder, decoys,
shifted duplicates of anchors -- verified twice through the C ABI: by a context with every shortcut on (tree levels on
the device, inferred inner answers, shared root passes and tracebacks) and by one that computes every window on its own
from the host (FXG_DEVICE_LEVELS=0 FXG_INFER_INNER=0 FXG_SHARE_ROOTS=0).  Alignments, CIGARs and statistics must be
identical; every fourth case is also checked against the CPU oracle.
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.getcwd())
sys.path.insert(0, os.path.join(os.getcwd(), "tests"))

from floxer_b200 import abi, synthetic  # noqa: E402
from floxer_b200 import gpu as g  # noqa: E402
from floxer_b200.batch import BatchBuilder, VerifyConfig, alignment_records  # noqa: E402


def merged_and_densified(batches, ref_lens, rng, spread, prob):
    bb = BatchBuilder()
    for batch in batches:
        for R in batch.reads:
            no, ni, nl = int(R["node_offset"]), int(R["num_inner"]), int(R["num_leaves"])
            qo, ql = int(R["query_offset"]), int(R["query_len"])
            ao, af, ar = int(R["anchor_offset"]), int(R["num_anchors_forward"]), int(R["num_anchors_reverse"])

            def densify(a):
                out = []
                for x in a:
                    out.append(tuple(int(v) for v in x))
                    if rng.random() < prob:
                        shift = int(rng.integers(-spread, spread + 1))
                        out.append((int(x[0]), int(x[1]), max(0, min(ref_lens[int(x[1])] - 1, int(x[2]) + shift)), int(x[3])))
                return np.array(sorted(out), dtype=abi.ANCHOR_DTYPE)
            bb.add(batch.forward_pool[qo:qo + ql], batch.reverse_pool[qo:qo + ql], batch.nodes[no:no + ni],
                   batch.nodes[no + ni:no + ni + nl], densify(batch.anchors[ao:ao + af]), densify(batch.anchors[ao + af:ao + af + ar]))
    return bb.build()


def context(env):
    saved = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    c = g.Context(0)
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    return c


def main(n_cases=None, seed=None, only=None, oracle_every=4):
    if n_cases is None:
        n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
        seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
        only = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3] != "-" else None    # run just this case (the others are generated and skipped)
        oracle_every = int(sys.argv[4]) if len(sys.argv) > 4 else oracle_every
    rng = np.random.default_rng(seed)
    fast = context({})
    plain = context({"FXG_DEVICE_LEVELS": "0", "FXG_INFER_INNER": "0", "FXG_SHARE_ROOTS": "0", "FXG_WORKERS": "2"})
    oracle = None
    t0 = time.time()
    totals = {"alignments": 0, "shared_score_passes": 0, "rescored_roots": 0, "inferred_inner": 0, "shared_tracebacks": 0}
    for case in range(n_cases):
        big = os.environ.get("STRESS_BIG") == "1"             # fewer, larger cases: several parts per batch, longer chains
        lens = sorted(int(x) for x in rng.choice([2000, 3500, 5000, 8000] if big else [60, 150, 400, 900, 1600, 3000],
                                                 size=int(rng.integers(1, 4)), replace=False))
        n_refs = int(rng.integers(1, 5))
        refs = []
        for r in range(n_refs):
            # some references barely hold the longest read: every window there is clipped
            length = int(lens[-1] * 1.3) + int(rng.integers(8, 200)) if rng.random() < 0.4 else int(rng.integers(lens[-1] * 2, 400_000 if big else 60_000))
            ref = synthetic.random_reference(length, int(rng.integers(1 << 30)))
            if rng.random() < 0.3 and length > 5000:
                ref = synthetic.plant_repeats(ref, int(rng.integers(1 << 30)), families=3, unit=(200, 800), copies=(2, 5))
            refs.append(ref)
        ref_lens = [len(r) for r in refs]
        batches = []
        for L in lens:
            err = float(rng.choice([0.02, 0.05, 0.08, 0.12]))
            b = synthetic.make_batch(refs, int(rng.integers(10, 40)) if big else int(rng.integers(2, 9)), L, err, int(rng.integers(1 << 30)), g.pex_build,
                                     seed_errors=int(rng.integers(0, 4)), decoy_fraction=float(rng.choice([0.0, 0.3, 0.8])),
                                     bottom_up=bool(rng.integers(0, 2)))
            batches.append(b)
        spread = int(rng.choice([3, 20, 80, 300]))
        batch = merged_and_densified(batches, ref_lens, rng, spread, float(rng.choice([0.0, 0.5, 0.9])))
        cfg = VerifyConfig(interval_optimization=bool(rng.integers(0, 2)), without_cigar=bool(rng.random() < 0.25),
                           verification_kind=abi.KIND_DIRECT_FULL if rng.random() < 0.15 else abi.KIND_HIERARCHICAL,
                           extra_verification_ratio=float(rng.choice([0.0, 0.05, 0.5, 2.0])))
        if only is not None and case != only:
            continue
        print(f"case {case}: lens {lens} refs {ref_lens} reads {len(batch)} anchors {len(batch.anchors)} spread {spread} cfg {cfg}", flush=True)
        if only is not None:
            plain.set_references(refs)
            jb = plain.verify_reads(batch, cfg)
            print("  anchors", batch.anchors.tolist())
            print("  reads", batch.reads.tolist())
            print("  plain:", alignment_records(*jb.alignments()), jb.stats())
        fast.set_references(refs)
        plain.set_references(refs)
        fast.reset_counters()
        ja = fast.verify_reads(batch, cfg)
        jb = plain.verify_reads(batch, cfg)
        ra, rb = alignment_records(*ja.alignments()), alignment_records(*jb.alignments())
        sa, sb = ja.stats(), jb.stats()
        ctr = fast.counters()
        for k in totals:
            totals[k] += len(ra) if k == "alignments" else int(ctr[k])
        ok = ra == rb and sa == sb
        if ok and oracle_every and case % oracle_every == 0:
            from harness import oracle_verify_batch
            if oracle is None:
                from oracle import oracle as oracle_module
                oracle_module.lib()
                oracle = oracle_module
            want, want_stats = oracle_verify_batch(oracle, refs, batch, cfg)
            ok = ra == want and sa == want_stats
        print(f"   -> {len(ra)} alignments {'ok' if ok else 'MISMATCH'}", flush=True)
        if not ok:
            print("  stats fast ", sa)
            print("  stats plain", sb)
            only_a = [r for r in ra if r not in rb][:5]
            only_b = [r for r in rb if r not in ra][:5]
            print("  only fast ", only_a)
            print("  only plain", only_b)
            return 1
        ja.free()
        jb.free()
    print(f"{n_cases} cases identical in {time.time() - t0:.1f} s; shortcuts exercised: {totals}")
    fast.close()
    plain.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
