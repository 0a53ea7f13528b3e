"""The q-gram seeder (floxer_b200/csrc/seeder.cpp, row N2) against the brute-force stand-in of tests/harness.py -- every
start position tried, min prefix edit distance, the reference's erase_useless_anchors -- and against the reference's own
known answer for erase_useless_anchors (test/search_test.cpp:138-184).  Host only."""
import numpy as np
import pytest

from floxer_b200 import abi, gpu, synthetic
from harness import brute_force_anchors, erase_useless_anchors


@pytest.mark.parametrize("seed,leaf_errors", [(1, 0), (2, 1), (3, 2)])
def test_seeder_finds_what_brute_force_finds(seed, leaf_errors):
    rng = np.random.default_rng(seed)
    refs = [rng.integers(1, 5, size=320, dtype=np.uint8), rng.integers(1, 5, size=150, dtype=np.uint8)]
    # a repeat, so that a leaf hits several loci, and a stretch of N
    refs[0][200:260] = refs[0][40:100]
    refs[1][20:26] = 5
    s = gpu.Seeder(refs, q=4)
    for trial in range(3):
        start = int(rng.integers(0, 200))
        read, _, _ = synthetic.simulate_read(rng, refs[0], start, 72, int(rng.integers(0, 6)))
        inner, leaves = gpu.pex_build(len(read), 5, leaf_errors, trial & 1)
        if any(int(l["query_index_to"] - l["query_index_from"] + 1) < 4 * (int(l["num_errors"]) + 1) for l in leaves):
            continue
        got = s.search(read, leaves, max_anchors_hard=10**9, max_anchors_soft=10**9, erase_useless=True)
        want = brute_force_anchors(read, leaves, refs, erase=True)
        assert [tuple(int(x) for x in a) for a in got] == sorted(want), (trial, len(got), len(want))
        raw = s.search(read, leaves, max_anchors_hard=10**9, max_anchors_soft=10**9, erase_useless=False)
        assert [tuple(int(x) for x in a) for a in raw] == sorted(brute_force_anchors(read, leaves, refs, erase=False))
    s.close()


def test_caps():
    ref = np.tile(np.array([1, 2, 3, 4, 1, 1, 2, 2, 3, 3, 4, 4], dtype=np.uint8), 100)          # every 12-mer occurs 99 times
    s = gpu.Seeder([ref], q=6)
    leaves = np.array([(abi.NULL_ID, 0, 11, 0), (abi.NULL_ID, 12, 23, 0)], dtype=abi.PEX_NODE_DTYPE)
    query = np.concatenate([ref[:12], np.array([4, 4, 4, 4, 3, 3, 3, 3, 2, 2, 2, 1], dtype=np.uint8)])
    a = s.search(query, leaves, max_anchors_hard=500, max_anchors_soft=50)
    assert len(a) == 50 and (a["pex_leaf_index"] == 0).all()                                   # soft cap; the second leaf has no hits
    assert len(s.search(query, leaves, max_anchors_hard=90, max_anchors_soft=50)) == 0          # more raw anchors than the hard cap: seed excluded
    s.close()


def test_erase_useless_known_answer():
    # test/search_test.cpp:138-184: positions/errors of one seed and reference; what survives
    anchors = [(100, 0), (101, 1), (102, 2), (200, 2), (201, 1), (300, 1), (302, 1), (400, 2), (401, 2)]
    kept = erase_useless_anchors(anchors)
    assert (100, 0) in kept and (101, 1) not in kept and (102, 2) not in kept
    assert (201, 1) in kept and (200, 2) not in kept
    assert (300, 1) in kept and (302, 1) in kept
