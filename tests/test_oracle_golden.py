"""Pins the CPU oracle (oracle/floxer_oracle.c) against every golden vector the reference's own tests
hold for the verification path.  CPU only."""
import itertools

import numpy as np
import pytest

import golden_vectors as G
from harness import brute_force_anchors, revcomp, to_ranks

STRATEGY = {"recursive": 0, "bottom_up": 1}


def test_alignment_small(oracle):
    # test/alignment_test.cpp:7-30
    v = G.ALIGNMENT_SMALL
    r = oracle.align(v["reference"], v["query"], v["max_errors"], oracle.MODE_CIGAR)
    assert r.exists and r.num_errors == v["num_errors"] and r.start == v["start"] and r.cigar == v["cigar"]
    assert oracle.align(v["reference"], v["query"], v["max_errors"], oracle.MODE_EXISTS).exists
    assert not oracle.align(v["reference"], v["query"], 0, oracle.MODE_EXISTS).exists


def test_math(oracle):
    # test/math_test.cpp:15-25
    L = oracle.lib()
    for (a, b), want in G.CEIL_DIV:
        assert L.fxo_ceil_div(a, b) == want
    for x, want in G.CEIL_EPS:
        assert L.fxo_ceil_eps(x) == want


def test_span_arithmetic(oracle):
    # test/verification_test.cpp:126-161
    s = G.SPAN
    for ratio, want in s["cases"]:
        assert oracle.compute_span(s["anchor_position"], s["node"], s["leaf_from"], s["ref_len"], ratio) == want


def test_interval_relationships(oracle):
    # test/intervals_test.cpp:35-65
    for a, b, rel in G.IVL_RELATIONS:
        assert oracle.interval_relationship(G.IVL[a], G.IVL[b]) == G.REL[rel], (a, b, rel)


def test_interval_trim(oracle):
    # test/intervals_test.cpp:67-89
    for base, amount, want in G.IVL_TRIM:
        assert oracle.interval_trim(base, amount) == want


def test_verified_intervals_do_not_merge(oracle):
    # test/intervals_test.cpp:91-157
    ivls = oracle.Intervals(True)
    for inserts, answers in G.IVL_CONTAINS_STEPS:
        for name in inserts:
            ivls.insert(*G.IVL[name])
        got = [ivls.contains(*G.IVL[p]) for p in G.IVL_PROBES]
        assert got == answers, inserts
    assert ivls.contains(*G.IVL["ivl1"]) and ivls.contains(*G.IVL["ivl2"])
    off = oracle.Intervals(False)
    off.insert(0, 100)
    assert not off.contains(5, 6) and len(off) == 0


def test_pex_trees(oracle):
    # test/pex_test.cpp:7-143
    for (total, errs, leaf_errs, strat), want in G.PEX_LEAVES:
        inner, leaves = oracle.pex_build(total, errs, leaf_errs, STRATEGY[strat])
        got = [(int(l["query_index_from"]), int(l["query_index_to"] - l["query_index_from"] + 1), int(l["num_errors"]))
               for l in leaves]
        assert got == want
        root = inner[0] if len(inner) else leaves[0]
        assert root["parent_id"] == oracle.NULL_ID and root["query_index_from"] == 0
        assert root["query_index_to"] == total - 1
        assert errs <= root["num_errors"] <= errs + leaf_errs          # pex.cpp:104-105
        for n in itertools.chain(inner[1:], leaves if len(inner) else []):
            p = inner[int(n["parent_id"])]
            assert p["query_index_from"] <= n["query_index_from"] and n["query_index_to"] <= p["query_index_to"]


def test_single_node(oracle):
    # test/verification_test.cpp:163-261
    n = G.NODE
    ref = np.array(G.NODE_REFERENCE, dtype=np.uint8)
    q = np.array(G.NODE_QUERY, dtype=np.uint8)
    window = ref[n["span_offset"]: n["span_offset"] + n["span_length"]]
    piece = q[n["node_from"]: n["node_to"] + 1]
    r = oracle.align(window, piece, n["num_errors"], oracle.MODE_CIGAR)
    assert r.exists and r.num_errors == n["expected"]["num_errors"]
    assert n["span_offset"] + r.start == n["expected"]["start"]
    assert oracle.align(window, piece, n["num_errors"], oracle.MODE_EXISTS).exists
    pos, val = n["extra_mismatch"]
    q2 = q.copy()
    q2[pos] = val
    assert not oracle.align(window, q2[n["node_from"]: n["node_to"] + 1], n["num_errors"], oracle.MODE_EXISTS).exists
    # through the verifier with a fake single-root tree: root hit is inserted, non-root is not
    leaves = np.array([(oracle.NULL_ID, n["node_from"], n["node_to"], n["num_errors"])], dtype=oracle.NODE_DTYPE)
    # (a root-only tree whose span is computed from an anchor; choose the anchor so the window is [50, 100))
    v = oracle.Verifier([ref], np.zeros(0, dtype=oracle.NODE_DTYPE), leaves, extra_verification_ratio=0.0)
    # window = anchor - (leaf.from - node.from) - k = anchor - 5 -> anchor 55; length 45 + 10 + 1 = 56 > 50, clipped at 100
    v.run(q, 0, [(0, 0, 55, 0)])
    al = v.alignments()
    assert len(al) == 1 and al[0][1] == 50 and al[0][2] == 5


def test_verify_hierarchical_and_direct(oracle):
    # test/verification_test.cpp:11-123
    V = G.VERIFY
    ref = np.array(G.VERIFY_REFERENCE, dtype=np.uint8)
    q = np.array(G.VERIFY_QUERY, dtype=np.uint8)
    t = V["tree"]
    inner, leaves = oracle.pex_build(t["total_len"], t["num_errors"], t["leaf_max_errors"], STRATEGY[t["strategy"]])
    a = V["anchor"]
    anchor = [(a["pex_leaf_index"], a["reference_id"], a["reference_position"], a["num_errors"])]
    v = oracle.Verifier([ref], inner, leaves, kind=oracle.KIND_HIERARCHICAL, interval_optimization=True,
                        extra_verification_ratio=V["extra_verification_ratio"])
    v.run(q, V["orientation"], anchor)
    e = V["expected"]
    assert v.alignments() == [(0, e["start"], e["num_errors"], e["orientation"], e["cigar"])]
    v.run(q, V["orientation"], anchor)                      # :84-87 no-op thanks to verified intervals
    assert len(v.alignments()) == 1
    assert v.stats()["n_avoided_root"] == 1
    d = oracle.Verifier([ref], inner, leaves, kind=oracle.KIND_DIRECT_FULL, interval_optimization=False,
                        extra_verification_ratio=V["extra_verification_ratio"])
    d.run(q, V["orientation"], anchor)                      # :94-112
    assert d.alignments() == v.alignments()
    q2 = q.copy()
    for pos, val in V["mutations"].items():
        q2[pos] = val
    d.run(q2, V["orientation"], anchor)                     # :114-122
    assert len(d.alignments()) == 1


def _whole_program(oracle, seed_errors, priority=None, rightmost=True):
    """Config 1: the reference's whole-program fixture, seeding replaced by the brute-force stand-in."""
    refs = [to_ranks(s) for s in G.WHOLE_REFERENCES.values()]
    F = G.WHOLE_FLAGS
    records = {}
    for qid, seq in G.WHOLE_QUERIES.items():
        fwd = to_ranks(seq)
        rc = revcomp(fwd)
        inner, leaves = oracle.pex_build(len(fwd), F["query_errors"], seed_errors, 0)
        v = oracle.Verifier(refs, inner, leaves, kind=oracle.KIND_HIERARCHICAL,
                            interval_optimization=F["interval_optimization"],
                            extra_verification_ratio=F["extra_verification_ratio"])
        v.run(fwd, 0, brute_force_anchors(fwd, leaves, refs))
        v.run(rc, 1, brute_force_anchors(rc, leaves, refs))
        records[qid] = v.alignments()
    return records


@pytest.mark.parametrize("seed_errors", G.WHOLE_FLAGS["seed_errors"])
def test_whole_program_fixture(oracle, seed_errors):
    # test/floxer_whole_program_via_cli_test.cpp:47-93,103-143
    records = _whole_program(oracle, seed_errors)
    for qid in G.WHOLE_UNMAPPED:
        assert records[qid] == []
    for (qid, rev), (lo, hi, nm, cigar) in G.WHOLE_EXPECT.items():
        recs = [r for r in records[qid] if bool(r[3]) == rev]
        assert recs, (qid, rev)
        for ref_id, start, errs, orient, cig in recs:
            assert ref_id == 0
            assert lo <= start <= hi, (qid, rev, start)
            assert errs == nm and cig == cigar, (qid, rev, cig)


def test_which_tie_breaks_the_golden_vectors_admit(oracle):
    """Documents what the reference's tests pin: rightmost end column and U before D.
    L's rank is not pinned: LUD, ULD and UDL all satisfy every expectation."""
    refs = [to_ranks(s) for s in G.WHOLE_REFERENCES.values()]
    ok = {}
    for prio in ("LUD", "LDU", "ULD", "UDL", "DLU", "DUL"):
        for rightmost in (True, False):
            good = True
            for (qid, rev), (lo, hi, nm, cigar) in G.WHOLE_EXPECT.items():
                q = to_ranks(G.WHOLE_QUERIES[qid])
                if rev:
                    q = revcomp(q)
                # the root window of the fixture covers the whole 71-bp reference (ratio 2)
                r = oracle.align(refs[0], q, 2, oracle.MODE_CIGAR, priority=prio, rightmost=rightmost)
                good &= r.exists and r.num_errors == nm and r.cigar == cigar and lo <= r.start <= hi
            ok[(prio, rightmost)] = good
    admitted = sorted(p for (p, rm), g in ok.items() if g and rm)
    assert admitted == ["LUD", "UDL", "ULD"]
    assert not any(g for (p, rm), g in ok.items() if not rm)


@pytest.mark.parametrize("seed_errors", G.WHOLE_FLAGS["seed_errors"])
def test_whole_program_fixture_sam_records(oracle, seed_errors):
    """The restatement of output.cpp:49-108 (oracle/sam_oracle.py) on the oracle's alignments gives the records that
    test/floxer_whole_program_via_cli_test.cpp:38-93 reads back: unmapped flag for query1/6, every query mentioned,
    one primary record per mapped query carrying SEQ, secondaries flagged 256 with SEQ '*'."""
    from oracle import sam_oracle
    records = _whole_program(oracle, seed_errors)
    ref_ids = list(G.WHOLE_REFERENCES)
    seen = set()
    for qid, seq in G.WHOLE_QUERIES.items():
        recs = sam_oracle.sam_records(qid, to_ranks(seq), "I" * len(seq), records[qid], ref_ids)
        seen.add(qid)
        if qid in G.WHOLE_UNMAPPED:
            assert [r[1] for r in recs] == [4] and recs[0][2] == "*" and recs[0][6] == seq.upper()
            continue
        assert sum(1 for r in recs if not r[1] & 256) == 1
        for qname, flag, rname, pos1, mapq, cigar, s, q, nm in recs:
            lo, hi, want_nm, want_cigar = G.WHOLE_EXPECT[(qid, bool(flag & 16))]
            assert not flag & 4 and mapq == 255 and rname == ref_ids[0]
            assert lo <= pos1 - 1 <= hi and nm == want_nm and cigar == want_cigar
            assert (s == "*") == bool(flag & 256)
    assert seen == set(G.WHOLE_QUERIES)
