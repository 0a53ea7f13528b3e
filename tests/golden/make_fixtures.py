#!/usr/bin/env python
"""Regenerates tests/golden/oracle_cases.json: frozen outputs of the CPU oracle (oracle/floxer_oracle.c) on seeded inputs.

The oracle itself is pinned against the reference's own golden vectors (tests/golden_vectors.py, transcribed by hand from
test/*.cpp of the reference with file:line per entry; tests/test_oracle_golden.py).  The cases frozen here are larger,
seeded ones: they keep the oracle from drifting (tests/test_golden_fixtures.py, CPU) and give the CUDA path a committed
target (same file, -m gpu).  Run from the repository root:  python tests/golden/make_fixtures.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

from floxer_b200 import abi, synthetic  # noqa: E402
from floxer_b200.batch import VerifyConfig  # noqa: E402
from harness import oracle_align_tasks, oracle_verify_batch, random_align_tasks  # noqa: E402
from oracle import oracle  # noqa: E402

VERIFY_CASES = [
    dict(name="hierarchical", cfg=dict()),
    dict(name="interval_optimization", cfg=dict(interval_optimization=True)),
    dict(name="direct_without_cigar", cfg=dict(verification_kind=abi.KIND_DIRECT_FULL, without_cigar=True)),
]
ALIGN_CASES = [
    dict(name="small_cigar", seed=101, n=60, m_range=(1, 200), err_range=(0.0, 0.25), mode=abi.MODE_CIGAR),
    dict(name="medium_cigar", seed=102, n=20, m_range=(300, 1500), err_range=(0.02, 0.15), mode=abi.MODE_CIGAR),
    dict(name="medium_no_cigar", seed=103, n=20, m_range=(300, 1500), err_range=(0.02, 0.15), mode=abi.MODE_NO_CIGAR),
    dict(name="exists", seed=104, n=60, m_range=(1, 900), err_range=(0.0, 0.2), mode=abi.MODE_EXISTS),
]


def verify_inputs():
    refs = [synthetic.random_reference(60_000, 901), synthetic.plant_repeats(synthetic.random_reference(40_000, 902), 903, families=3, unit=(200, 600), copies=(3, 5))]
    batch = synthetic.make_batch(refs, 8, 700, 0.07, 904, oracle.pex_build, seed_errors=1, decoy_fraction=0.4)
    return refs, batch


def align_inputs(case):
    rng = np.random.default_rng(case["seed"])
    return random_align_tasks(rng, case["n"], case["m_range"], case["err_range"], case["mode"], ref_len=40_000)


def build():
    out = {"verify": {}, "align": {}}
    refs, batch = verify_inputs()
    for case in VERIFY_CASES:
        recs, stats = oracle_verify_batch(oracle, refs, batch, VerifyConfig(**case["cfg"]))
        out["verify"][case["name"]] = {"records": [list(r) for r in recs], "stats": stats}
    for case in ALIGN_CASES:
        ref, tasks, pool = align_inputs(case)
        out["align"][case["name"]] = [list(r) for r in oracle_align_tasks(oracle, ref, tasks, pool)]
    return out


if __name__ == "__main__":
    oracle.build()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_cases.json")
    with open(path, "w") as f:
        json.dump(build(), f, separators=(",", ":"))
    print(path, os.path.getsize(path), "bytes")
