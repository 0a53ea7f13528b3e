"""Parity of the CUDA path (through the C ABI) against the CPU oracle.  Needs a B200: run with -m gpu."""
import numpy as np
import pytest

import golden_vectors as G
from floxer_b200 import abi, synthetic
from floxer_b200.batch import BatchBuilder, VerifyConfig, alignment_records
from harness import (brute_force_anchors, oracle_align_tasks, oracle_verify_batch, random_align_tasks,
                     results_as_tuples, revcomp, to_ranks)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    from floxer_b200 import build, gpu as g
    build.build_native()
    g.lib()
    return g


@pytest.fixture(scope="module")
def ctx(gpu):
    c = gpu.Context(0)
    yield c
    c.close()


# ----------------------------------------------------------------------------- alignment::align

@pytest.mark.parametrize("mode", [abi.MODE_EXISTS, abi.MODE_NO_CIGAR, abi.MODE_CIGAR])
@pytest.mark.parametrize("m_range,err_range,n_tasks", [
    ((1, 40), (0.0, 0.3), 400),        # single word, k = 0, tiny queries
    ((30, 140), (0.0, 0.2), 400),      # 1..5 words: one lane per task
    ((120, 700), (0.0, 0.15), 300),    # rings of a few lanes
    ((600, 2500), (0.02, 0.15), 120),  # microbench sizes (config 5)
    ((2500, 6000), (0.03, 0.12), 24),  # root of a 5 kbp read
])
def test_align_matches_oracle(ctx, oracle, mode, m_range, err_range, n_tasks):
    rng = np.random.default_rng(4242 + 13 * mode + m_range[0])
    ref, tasks, pool = random_align_tasks(rng, n_tasks, m_range, err_range, mode, ref_len=60_000)
    ctx.set_references([ref])
    res, cig = ctx.align_batch(tasks, pool)
    got = results_as_tuples(res, cig, tasks)
    want = oracle_align_tasks(oracle, ref, tasks, pool)
    bad = [i for i, (g, w) in enumerate(zip(got, want)) if g != w]
    assert not bad, (bad[:5], [(got[i], want[i], tasks[i]) for i in bad[:2]])
    assert (res["orientation"] == tasks["orientation"]).all()


@pytest.mark.parametrize("mode", [abi.MODE_EXISTS, abi.MODE_NO_CIGAR, abi.MODE_CIGAR])
def test_align_band_of_one_diagonal(ctx, oracle, mode):
    """k = 0 and a window exactly as long as the query: the band is a single diagonal (queries of several blocks:
    every block starts where the block above has just ended), with and without one mismatch."""
    rng = np.random.default_rng(99 + mode)
    ref = rng.integers(1, 5, size=40_000, dtype=np.uint8)
    tasks, pool, off = [], [], 0
    for i, m in enumerate([33, 64, 65, 100, 129, 257, 300, 640, 1025, 1500, 2049, 3000, 5000, 9000]):
        for flip in (None, 0, m // 2, m - 1):
            at = int(rng.integers(0, len(ref) - m))
            q = ref[at:at + m].copy()
            if flip is not None:
                q[flip] = 1 + (q[flip] % 4)
            pool.append(q)
            tasks.append((at, at, off, m, m, 0, 0, mode, i & 1, (0,) * 6))
            off += m
    tasks = np.array(tasks, dtype=abi.ALIGN_TASK_DTYPE)
    pool = np.concatenate(pool)
    ctx.set_references([ref])
    res, cig = ctx.align_batch(tasks, pool)
    got = results_as_tuples(res, cig, tasks)
    want = oracle_align_tasks(oracle, ref, tasks, pool)
    assert got == want
    assert sum(g[0] for g in got) == len(got) // 4


def test_align_inline_reference_and_all_ranks(ctx, oracle):
    """Spans handed over as host bytes (the alignment::align shim) with ranks 0..5 incl. N == N matches."""
    rng = np.random.default_rng(7)
    v = G.ALIGNMENT_SMALL                                    # test/alignment_test.cpp:7-30
    refs_inline = [np.array(v["reference"], dtype=np.uint8)]
    queries = [np.array(v["query"], dtype=np.uint8)]
    ks = [v["max_errors"]]
    for _ in range(200):
        n = int(rng.integers(5, 400))
        r = rng.integers(0, 6, size=n, dtype=np.uint8)
        m = int(rng.integers(1, n + 10))
        q = rng.integers(0, 6, size=m, dtype=np.uint8)
        if rng.random() < 0.6 and m <= n:                    # plant a noisy copy
            at = int(rng.integers(0, n - m + 1))
            q = r[at:at + m].copy()
            flips = rng.integers(0, m, size=int(rng.integers(0, max(1, m // 8) + 1)))
            q[flips] = rng.integers(0, 6, size=len(flips), dtype=np.uint8)
        refs_inline.append(r); queries.append(q); ks.append(int(rng.integers(0, max(2, m // 5))))
    ipool = np.concatenate(refs_inline); qpool = np.concatenate(queries)
    roff = np.concatenate([[0], np.cumsum([len(r) for r in refs_inline])])
    qoff = np.concatenate([[0], np.cumsum([len(q) for q in queries])])
    for mode in (abi.MODE_EXISTS, abi.MODE_NO_CIGAR, abi.MODE_CIGAR):
        tasks = np.array([(roff[i], 1000 + i, qoff[i], len(refs_inline[i]), len(queries[i]), abi.REF_INLINE, ks[i], mode, i & 1, (0,) * 6)
                          for i in range(len(queries))], dtype=abi.ALIGN_TASK_DTYPE)
        res, cig = ctx.align_batch(tasks, qpool, ipool)
        got = results_as_tuples(res, cig, tasks)
        want = oracle_align_tasks(oracle, ipool, tasks, qpool)
        assert got == want
    # the golden vector itself
    assert got[0] == (True, v["num_errors"], 1000 + v["start"], v["cigar"])


def test_rejects_bad_input(ctx, gpu):
    ref = np.ones(100, dtype=np.uint8)
    ctx.set_references([ref])
    q = np.ones(10, dtype=np.uint8)
    bad_mode = np.array([(0, 0, 0, 50, 10, 0, 1, 3, 0, (0,) * 6)], dtype=abi.ALIGN_TASK_DTYPE)
    with pytest.raises(gpu.FloxerGpuError):
        ctx.align_batch(bad_mode, q)
    outside = np.array([(90, 90, 0, 50, 10, 0, 1, 0, 0, (0,) * 6)], dtype=abi.ALIGN_TASK_DTYPE)
    with pytest.raises(gpu.FloxerGpuError):
        ctx.align_batch(outside, q)
    q_bad = q.copy(); q_bad[3] = 9
    ok = np.array([(0, 0, 0, 50, 10, 0, 1, 0, 0, (0,) * 6)], dtype=abi.ALIGN_TASK_DTYPE)
    with pytest.raises(gpu.FloxerGpuError):
        ctx.align_batch(ok, q_bad)
    res, _ = ctx.align_batch(ok, q)
    assert res["exists"][0] == 1
    empty = np.zeros(0, dtype=abi.ALIGN_TASK_DTYPE)
    res, cig = ctx.align_batch(empty, q)
    assert len(res) == 0 and len(cig) == 0


# ----------------------------------------------------------------------------- query_verifier::verify

def _one_read_batch(fwd, rc, inner, leaves, af, ar):
    bb = BatchBuilder()
    bb.add(fwd, rc, inner, leaves, np.array(af, dtype=abi.ANCHOR_DTYPE), np.array(ar, dtype=abi.ANCHOR_DTYPE))
    return bb.build()


def test_golden_verify(ctx, gpu):
    # test/verification_test.cpp:11-123 through the GPU path
    V = G.VERIFY
    ref = np.array(G.VERIFY_REFERENCE, dtype=np.uint8)
    q = np.array(G.VERIFY_QUERY, dtype=np.uint8)
    t = V["tree"]
    inner, leaves = gpu.pex_build(t["total_len"], t["num_errors"], t["leaf_max_errors"], 1)
    a = V["anchor"]
    anchor = (a["pex_leaf_index"], a["reference_id"], a["reference_position"], a["num_errors"])
    ctx.set_references([ref])
    e = V["expected"]
    # the test verifies with orientation reverse_complement: put the query into the reverse pool
    batch = _one_read_batch(q, q, inner, leaves, [], [anchor, anchor])
    cfg = VerifyConfig(interval_optimization=True, extra_verification_ratio=V["extra_verification_ratio"])
    job = ctx.verify_reads(batch, cfg)
    al, cg = job.alignments()
    assert alignment_records(al, cg) == [(0, 0, e["start"], e["num_errors"], e["orientation"], e["cigar"])]
    assert job.stats()["n_avoided_root"] == 1              # second verify() is a no-op (:84-87)
    direct = VerifyConfig(verification_kind=abi.KIND_DIRECT_FULL, interval_optimization=False,
                          extra_verification_ratio=V["extra_verification_ratio"])
    al, cg = ctx.verify_reads(_one_read_batch(q, q, inner, leaves, [], [anchor]), direct).alignments()
    assert alignment_records(al, cg) == [(0, 0, e["start"], e["num_errors"], e["orientation"], e["cigar"])]
    q2 = q.copy()
    for pos, val in V["mutations"].items():
        q2[pos] = val
    al, cg = ctx.verify_reads(_one_read_batch(q2, q2, inner, leaves, [], [anchor]), direct).alignments()
    assert len(al) == 0


def test_golden_single_node(ctx):
    # test/verification_test.cpp:163-261
    n = G.NODE
    ref = np.array(G.NODE_REFERENCE, dtype=np.uint8)
    q = np.array(G.NODE_QUERY, dtype=np.uint8)
    ctx.set_references([ref])
    def task(mode, pool):
        return np.array([(n["span_offset"], n["span_offset"], n["node_from"], n["span_length"],
                          n["node_to"] - n["node_from"] + 1, 0, n["num_errors"], mode, 0, (0,) * 6)], dtype=abi.ALIGN_TASK_DTYPE)
    res, _ = ctx.align_batch(task(abi.MODE_CIGAR, q), q)
    assert res["exists"][0] and res["num_errors"][0] == n["expected"]["num_errors"]
    assert res["start_in_reference"][0] == n["expected"]["start"]
    res, _ = ctx.align_batch(task(abi.MODE_EXISTS, q), q)
    assert res["exists"][0] == 1
    pos, val = n["extra_mismatch"]
    q2 = q.copy(); q2[pos] = val
    res, _ = ctx.align_batch(task(abi.MODE_EXISTS, q2), q2)
    assert res["exists"][0] == 0


@pytest.mark.parametrize("seed_errors", G.WHOLE_FLAGS["seed_errors"])
def test_golden_whole_program_fixture(ctx, gpu, oracle, seed_errors):
    # config 1: test/floxer_whole_program_via_cli_test.cpp:47-93 (seeding replaced by the brute-force stand-in)
    refs = [to_ranks(s) for s in G.WHOLE_REFERENCES.values()]
    ctx.set_references(refs)
    F = G.WHOLE_FLAGS
    bb = BatchBuilder()
    names = list(G.WHOLE_QUERIES)
    for qid in names:
        fwd = to_ranks(G.WHOLE_QUERIES[qid]); rc = revcomp(fwd)
        inner, leaves = gpu.pex_build(len(fwd), F["query_errors"], seed_errors, 0)
        bb.add(fwd, rc, inner, leaves,
               np.array(brute_force_anchors(fwd, leaves, refs), dtype=abi.ANCHOR_DTYPE),
               np.array(brute_force_anchors(rc, leaves, refs), dtype=abi.ANCHOR_DTYPE))
    batch = bb.build()
    cfg = VerifyConfig(interval_optimization=F["interval_optimization"], extra_verification_ratio=F["extra_verification_ratio"])
    job = ctx.verify_reads(batch, cfg)
    al, cg = job.alignments()
    recs = alignment_records(al, cg)
    want, want_stats = oracle_verify_batch(oracle, refs, batch, cfg)
    assert recs == want and job.stats() == want_stats
    by_query = {q: [r for r in recs if names[r[0]] == q] for q in names}
    for qid in G.WHOLE_UNMAPPED:
        assert by_query[qid] == []
    for (qid, rev), (lo, hi, nm, cigar) in G.WHOLE_EXPECT.items():
        rs = [r for r in by_query[qid] if bool(r[4]) == rev]
        assert rs
        for _, ref_id, start, errs, _, cig in rs:
            assert ref_id == 0 and lo <= start <= hi and errs == nm and cig == cigar


CONFIGS = [
    VerifyConfig(interval_optimization=False),
    VerifyConfig(interval_optimization=True),
    VerifyConfig(interval_optimization=True, without_cigar=True),
    VerifyConfig(interval_optimization=False, without_cigar=True),
    VerifyConfig(verification_kind=abi.KIND_DIRECT_FULL, interval_optimization=False),
    VerifyConfig(verification_kind=abi.KIND_DIRECT_FULL, interval_optimization=True),
    VerifyConfig(interval_optimization=True, extra_verification_ratio=0.4),
    VerifyConfig(interval_optimization=True, extra_verification_ratio=0.0),
]


@pytest.mark.parametrize("cfg", CONFIGS)
def test_verify_reads_matches_oracle(ctx, gpu, oracle, cfg):
    refs = [synthetic.random_reference(80_000, 21), synthetic.random_reference(30_000, 22)]
    ctx.set_references(refs)
    batch = synthetic.make_batch(refs, 10, 900, 0.07, 123, gpu.pex_build, seed_errors=1, decoy_fraction=0.4)
    job = ctx.verify_reads(batch, cfg)
    al, cg = job.alignments()
    want, want_stats = oracle_verify_batch(oracle, refs, batch, cfg)
    assert alignment_records(al, cg) == want
    assert job.stats() == want_stats
    assert len(want) > 0


@pytest.mark.parametrize("bottom_up", [False, True])
def test_verify_reads_repeats_and_dense_anchors(ctx, gpu, oracle, bottom_up):
    """Repeat-seeded reference, many overlapping anchors per locus: stresses the interval-dependency scheduler."""
    ref = synthetic.plant_repeats(synthetic.random_reference(120_000, 31), 32, families=6, unit=(300, 900), copies=(3, 8))
    refs = [ref]
    ctx.set_references(refs)
    batch = synthetic.make_batch(refs, 8, 1200, 0.05, 77, gpu.pex_build, seed_errors=2, decoy_fraction=0.6, bottom_up=bottom_up)
    # add shifted duplicates of every anchor so that root windows nest in both directions
    rng = np.random.default_rng(3)
    bb = BatchBuilder()
    for R in batch.reads:
        no, ni, nl = int(R["node_offset"]), int(R["num_inner"]), int(R["num_leaves"])
        qo, ql = int(R["query_offset"]), int(R["query_len"])
        ao, af, ar = int(R["anchor_offset"]), int(R["num_anchors_forward"]), int(R["num_anchors_reverse"])
        def densify(a):
            out = []
            for x in a:
                out.append(tuple(int(v) for v in x))
                if rng.random() < 0.7:
                    shift = int(rng.integers(-60, 61))
                    out.append((int(x[0]), int(x[1]), max(0, min(len(ref) - 1, int(x[2]) + shift)), int(x[3])))
            return np.array(sorted(out), dtype=abi.ANCHOR_DTYPE)
        bb.add(batch.forward_pool[qo:qo + ql], batch.reverse_pool[qo:qo + ql], batch.nodes[no:no + ni], batch.nodes[no + ni:no + ni + nl],
               densify(batch.anchors[ao:ao + af]), densify(batch.anchors[ao + af:ao + af + ar]))
    dense = bb.build()
    for cfg in (VerifyConfig(interval_optimization=True), VerifyConfig(interval_optimization=False)):
        job = ctx.verify_reads(dense, cfg)
        al, cg = job.alignments()
        want, want_stats = oracle_verify_batch(oracle, refs, dense, cfg)
        assert alignment_records(al, cg) == want
        assert job.stats() == want_stats


def _densified(batch, ref_lens, rng, spread, prob):
    """every anchor followed (with probability prob) by a copy shifted by up to +-spread bases"""
    bb = BatchBuilder()
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
        bb.add(batch.forward_pool[qo:qo + ql], batch.reverse_pool[qo:qo + ql], batch.nodes[no:no + ni], batch.nodes[no + ni:no + ni + nl],
               densify(batch.anchors[ao:ao + af]), densify(batch.anchors[ao + af:ao + af + ar]))
    return bb.build()


@pytest.mark.parametrize("ratio", [0.0, 0.05, 0.3])
def test_shared_root_passes_and_rescoring(ctx, gpu, oracle, ratio):
    """Root windows of one locus are scored by one pass over their union; a window the shared pass cannot vouch for
    (tight windows, alignments that begin before it) is scored again on its own.  Same answers as window by window."""
    refs = [synthetic.random_reference(90_000, 71), synthetic.random_reference(40_000, 72)]
    ctx.set_references(refs)
    base = synthetic.make_batch(refs, 12, 1000, 0.08, 73, gpu.pex_build, seed_errors=1, decoy_fraction=0.3)
    dense = _densified(base, [90_000, 40_000], np.random.default_rng(5), 70, 0.8)
    for cfg in (VerifyConfig(extra_verification_ratio=ratio), VerifyConfig(extra_verification_ratio=ratio, verification_kind=abi.KIND_DIRECT_FULL)):
        ctx.reset_counters()
        job = ctx.verify_reads(dense, cfg)
        al, cg = job.alignments()
        want, want_stats = oracle_verify_batch(oracle, refs, dense, cfg)
        assert alignment_records(al, cg) == want and job.stats() == want_stats and len(want) > 20
        c = ctx.counters()
        assert c["shared_score_passes"] > 0
        if ratio == 0.0:
            assert c["rescored_roots"] > 0           # the tight windows make some members fall back
        job.free()


def test_sharing_knobs_do_not_change_results(gpu, monkeypatch):
    refs = [synthetic.random_reference(120_000, 81)]
    batch = _densified(synthetic.make_batch(refs, 10, 1500, 0.06, 83, gpu.pex_build, seed_errors=2, decoy_fraction=0.3),
                       [120_000], np.random.default_rng(7), 40, 0.7)
    results = []
    for share, infer, device, roots in (("1", "1", "1", "1"), ("0", "0", "1", "1"), ("1", "1", "0", "1"), ("0", "0", "0", "1"), ("1", "1", "1", "0"), ("0", "1", "1", "0")):
        monkeypatch.setenv("FXG_SHARE_ROOTS", share)
        monkeypatch.setenv("FXG_INFER_INNER", infer)
        monkeypatch.setenv("FXG_DEVICE_LEVELS", device)
        monkeypatch.setenv("FXG_DEVICE_ROOTS", roots)
        c2 = gpu.Context(0)
        try:
            c2.set_references(refs)
            for cfg in (VerifyConfig(), VerifyConfig(interval_optimization=True)):
                job = c2.verify_reads(batch, cfg)
                al, cg = job.alignments()
                results.append((alignment_records(al, cg), job.stats()))
                cnt = c2.counters()
                assert (cnt["shared_score_passes"] > 0) == (share == "1")
                job.free()
        finally:
            c2.close()
    for k in range(2, len(results)):
        assert results[k] == results[k % 2]
    assert len(results[0][0]) > 0


# ----------------------------------------------------------------------------- bands wider than one warp's ring

def test_wide_kernel_matches_oracle_at_small_sizes(gpu, oracle, monkeypatch):
    """FXG_FORCE_WIDE=1 sends every pass to the multi-warp kernel (dp_wide_kernel), so that it meets the oracle on many
    random tasks in all three modes (incl. clipped windows, k = 0, inline references with all six ranks) and on whole reads."""
    monkeypatch.setenv("FXG_FORCE_WIDE", "1")
    c2 = gpu.Context(0)
    try:
        for mode in (abi.MODE_EXISTS, abi.MODE_NO_CIGAR, abi.MODE_CIGAR):
            for m_range, err_range, n_tasks in (((1, 60), (0.0, 0.3), 120), ((100, 900), (0.0, 0.15), 80), ((900, 2600), (0.02, 0.12), 30)):
                rng = np.random.default_rng(977 + 17 * mode + m_range[0])
                ref, tasks, pool = random_align_tasks(rng, n_tasks, m_range, err_range, mode, ref_len=40_000)
                c2.set_references([ref])
                res, cig = c2.align_batch(tasks, pool)
                got = results_as_tuples(res, cig, tasks)
                want = oracle_align_tasks(oracle, ref, tasks, pool)
                bad = [i for i, (g, w) in enumerate(zip(got, want)) if g != w]
                assert not bad, (mode, m_range, bad[:5], [(got[i], want[i], tasks[i]) for i in bad[:2]])
        # bands of a single diagonal over several of the kernel's 1 024-row blocks
        rng = np.random.default_rng(31)
        ref = rng.integers(1, 5, size=20_000, dtype=np.uint8)
        c2.set_references([ref])
        for mode in (abi.MODE_EXISTS, abi.MODE_NO_CIGAR, abi.MODE_CIGAR):
            tasks, pool, off = [], [], 0
            for m in (1025, 2049, 3000, 5000):
                for flip in (None, 0, m // 2, m - 1):
                    at = int(rng.integers(0, len(ref) - m))
                    q = ref[at:at + m].copy()
                    if flip is not None:
                        q[flip] = 1 + (q[flip] % 4)
                    pool.append(q)
                    tasks.append((at, at, off, m, m, 0, 0, mode, 0, (0,) * 6))
                    off += m
            tasks = np.array(tasks, dtype=abi.ALIGN_TASK_DTYPE)
            pool = np.concatenate(pool)
            res, cig = c2.align_batch(tasks, pool)
            assert results_as_tuples(res, cig, tasks) == oracle_align_tasks(oracle, ref, tasks, pool)
        refs = [synthetic.random_reference(100_000, 5)]
        c2.set_references(refs)
        batch = synthetic.make_batch(refs, 5, 1300, 0.07, 41, gpu.pex_build, seed_errors=1, decoy_fraction=0.3)
        for cfg in (VerifyConfig(), VerifyConfig(interval_optimization=True), VerifyConfig(without_cigar=True)):
            job = c2.verify_reads(batch, cfg)
            al, cg = job.alignments()
            want, want_stats = oracle_verify_batch(oracle, refs, batch, cfg)
            assert alignment_records(al, cg) == want and job.stats() == want_stats
            assert len(want) > 0
            job.free()
    finally:
        c2.close()


@pytest.mark.parametrize("error_rate", [0.05, 0.10, 0.15])
def test_reads_at_the_length_limit(ctx, gpu, error_rate):
    """A 100 kbp read (input.hpp:42) at 5, 10 and 15 %: root bands of 31 000 .. 73 000 diagonals, which no ring of one
    warp holds (the multi-warp kernel takes them).  Alignment in all three modes and the whole read through verify_reads;
    the CPU port (bit-vector, checked against the plain-DP oracle in the CPU suite) is the checker at this size."""
    from oracle import cpu_baseline
    rng = np.random.default_rng(int(error_rate * 1000))
    m = 100_000
    ref = synthetic.random_reference(400_000, 31)
    k = int(np.ceil(m * error_rate))
    read, _, _ = synthetic.simulate_read(rng, ref, 150_000, m - 2000, int((m - 2000) * error_rate * 0.9))
    read = read[:m]
    extra = synthetic.ceil_eps((len(read) + 2 * k + 1) * 0.05)
    n = len(read) + 2 * k + 1 + 2 * extra
    at = 150_000 - k - extra
    tasks = np.array([(at, at, 0, n, len(read), 0, k, mode, 0, (0,) * 6) for mode in (abi.MODE_EXISTS, abi.MODE_NO_CIGAR, abi.MODE_CIGAR)], dtype=abi.ALIGN_TASK_DTYPE)
    ctx.set_references([ref])
    res, cig = ctx.align_batch(tasks, read)
    cres, ccig = cpu_baseline.align_batch([ref], tasks, read, threads=3)
    assert results_as_tuples(res, cig, tasks) == results_as_tuples(cres, ccig, tasks)
    assert res["exists"].all()
    # the same read with its tree and a handful of anchors
    inner, leaves = gpu.pex_build(len(read), k, 2, 0)
    ok = [li for li, lf in enumerate(leaves) if True][:: max(1, len(leaves) // 12)]
    # (ground truth is not tracked here: anchors at the diagonal of the simulated locus; most leaves carry too many edits,
    #  which is the common case for a real seeder as well -- the walks that fail at an inner level are part of the test)
    af = np.array([(li, 0, min(len(ref) - 1, 150_000 + int(leaves[li]["query_index_from"])), 0) for li in ok], dtype=abi.ANCHOR_DTYPE)
    bb = BatchBuilder()
    bb.add(read, revcomp(read), inner, leaves, af, np.zeros(0, dtype=abi.ANCHOR_DTYPE))
    batch = bb.build()
    cfg = VerifyConfig(interval_optimization=True)
    job = ctx.verify_reads(batch, cfg)
    al, cg = job.alignments()
    wal, wcg, wstats = cpu_baseline.verify_reads([ref], batch, cfg, threads=4)
    assert alignment_records(al, cg) == alignment_records(wal, wcg) and job.stats() == wstats
    job.free()


def test_config2_shape_against_cpu_port(ctx, gpu):
    """5 kbp reads at 5 % (config 2 shape, fewer reads): the bit-vector CPU port (itself checked against the
    oracle in the CPU suite) is the checker at this size."""
    from oracle import cpu_baseline
    refs = [synthetic.random_reference(2_000_000, 20240001)]
    ctx.set_references(refs)
    batch = synthetic.make_batch(refs, 6, 5000, 0.05, 9, gpu.pex_build, seed_errors=2, decoy_fraction=0.2)
    for cfg in (VerifyConfig(interval_optimization=False), VerifyConfig(interval_optimization=True)):
        job = ctx.verify_reads(batch, cfg)
        al, cg = job.alignments()
        wal, wcg, wstats = cpu_baseline.verify_reads(refs, batch, cfg, threads=8)
        assert alignment_records(al, cg) == alignment_records(wal, wcg)
        assert job.stats() == wstats
        assert len(al) > 0


def test_staged_run_is_repeatable(ctx, gpu):
    refs = [synthetic.random_reference(100_000, 41)]
    ctx.set_references(refs)
    batch = synthetic.make_batch(refs, 4, 1500, 0.06, 5, gpu.pex_build)
    job = ctx.stage_verify(batch, VerifyConfig())
    a1 = job.run().alignments()
    a2 = job.run().alignments()
    assert np.array_equal(a1[0], a2[0]) and np.array_equal(a1[1], a2[1]) and len(a1[0]) > 0
    c = ctx.counters()
    assert c["kernel_launches"] > 0 and c["dp_word_steps"] > 0


# ----------------------------------------------------------------------------- larger shapes, chunking, properties

def _thin_anchors(batch, keep_every: int):
    """Keeps every `keep_every`-th anchor of each orientation (root alignments of long reads are expensive for the CPU checker)."""
    bb = BatchBuilder()
    for R in batch.reads:
        no, ni, nl = int(R["node_offset"]), int(R["num_inner"]), int(R["num_leaves"])
        qo, ql = int(R["query_offset"]), int(R["query_len"])
        ao, af, ar = int(R["anchor_offset"]), int(R["num_anchors_forward"]), int(R["num_anchors_reverse"])
        bb.add(batch.forward_pool[qo:qo + ql], batch.reverse_pool[qo:qo + ql], batch.nodes[no:no + ni], batch.nodes[no + ni:no + ni + nl],
               batch.anchors[ao:ao + af][::keep_every], batch.anchors[ao + af:ao + af + ar][::keep_every])
    return bb.build()


@pytest.mark.parametrize("length,err,seed,keep", [(15000, 0.08, 11, 40), (20000, 0.10, 12, 60)])
def test_long_reads_config3_config4_shapes(ctx, gpu, length, err, seed, keep):
    """15 kbp at 8 % and 20 kbp at 10 % (configs 3 and 4): wide blocks (W = 8, 16), bands of thousands of diagonals,
    tracebacks over hundreds of checkpoint tiles; the bit-vector CPU port is the checker at this size."""
    from oracle import cpu_baseline
    ref = synthetic.plant_repeats(synthetic.random_reference(1_200_000, 20240002), 33, families=4, unit=(500, 3000), copies=(3, 6))
    refs = [ref]
    ctx.set_references(refs)
    batch = _thin_anchors(synthetic.make_batch(refs, 2, length, err, seed, gpu.pex_build, seed_errors=2, decoy_fraction=0.1), keep)
    for cfg in (VerifyConfig(), VerifyConfig(without_cigar=True), VerifyConfig(verification_kind=abi.KIND_DIRECT_FULL)):
        job = ctx.verify_reads(batch, cfg)
        al, cg = job.alignments()
        wal, wcg, wstats = cpu_baseline.verify_reads(refs, batch, cfg, threads=8)
        assert alignment_records(al, cg) == alignment_records(wal, wcg)
        assert job.stats() == wstats
        assert len(al) > 0
        job.free()


def test_checkpoint_chunks_and_small_budget(gpu, oracle, monkeypatch):
    """Root alignments cut into several chunks (checkpoint buffers reused while tracebacks are in flight) give the same answer."""
    monkeypatch.setenv("FXG_ROOT_CHUNKS", "7")
    monkeypatch.setenv("FXG_ROOT_CHUNK_MIN", "1")
    monkeypatch.setenv("FXG_WORKERS", "2")
    c2 = gpu.Context(0)
    try:
        refs = [synthetic.random_reference(150_000, 51)]
        c2.set_references(refs)
        batch = synthetic.make_batch(refs, 30, 1100, 0.06, 61, gpu.pex_build, seed_errors=1, decoy_fraction=0.3)
        for cfg in (VerifyConfig(), VerifyConfig(interval_optimization=True), VerifyConfig(verification_kind=abi.KIND_DIRECT_FULL)):
            job = c2.verify_reads(batch, cfg)
            al, cg = job.alignments()
            want, want_stats = oracle_verify_batch(oracle, refs, batch, cfg)
            assert alignment_records(al, cg) == want and job.stats() == want_stats and len(want) > 10
            job.free()
    finally:
        c2.close()


def _check_alignment_against_sequences(a, cigar_pool, batch, refs):
    """Size-independent property: the CIGAR, applied to the read and the reference, consumes the whole read, says '='
    exactly where the bases agree and 'X' where they differ, and its I/D/X total is the reported number of errors."""
    R = batch.reads[int(a["read_index"])]
    qo, ql = int(R["query_offset"]), int(R["query_len"])
    pool = batch.reverse_pool if int(a["orientation"]) else batch.forward_pool
    q = pool[qo:qo + ql]
    ops = cigar_pool[int(a["cigar_offset"]): int(a["cigar_offset"]) + int(a["cigar_len"])]
    lens, codes = (ops >> 4).astype(np.int64), (ops & 15).astype(np.int64)
    assert set(np.unique(codes)) <= {1, 2, 7, 8} and (lens > 0).all()
    assert (codes[1:] != codes[:-1]).all(), "adjacent runs of the same operation"
    per = np.repeat(codes, lens)
    q_use, r_use = per != 2, per != 1
    assert int(q_use.sum()) == ql
    qi = np.cumsum(q_use) - 1
    ri = int(a["start_in_reference"]) + np.cumsum(r_use) - 1
    ref = refs[int(a["reference_id"])]
    assert ri[-1] < len(ref) or not r_use[-1]
    both = q_use & r_use
    same = q[qi[both]] == ref[ri[both]]
    assert np.array_equal(same, per[both] == 7)
    assert int((per != 7).sum()) == int(a["num_errors"])


def test_bench_workload_properties_and_sample_parity(gpu):
    """BASELINE config 2 at full size (10 Mbp reference, 1 000 reads x 5 kbp at 5 %): every 7th alignment is checked against
    the sequences, the counts against the simulation's ground truth, and the first reads bit for bit against the CPU port."""
    import bench
    from oracle import cpu_baseline
    refs, batch, _ = bench.build_workload("config2", 0, gpu.pex_build, None, 8)
    c2 = gpu.Context(0)
    try:
        c2.set_references(refs)
        cfg = VerifyConfig()
        job = c2.stage_verify(batch, cfg)
        al, cg = job.run().alignments()
        k = int(np.ceil(5000 * 0.05)) + 14
        assert len(al) > 10 * len(batch) and (al["num_errors"] <= k).all()
        # every read is found on the strand and near the position it was simulated from
        for ri, (rid, start, on_reverse) in enumerate(batch.meta["truth"][:200]):
            mine = al[al["read_index"] == ri]
            assert len(mine) and (mine["orientation"] == int(on_reverse)).any()
            assert (np.abs(mine["start_in_reference"].astype(np.int64) - start) <= 600).any()
        for a in al[::7]:
            _check_alignment_against_sequences(a, cg, batch, refs)
        st = job.stats()
        assert st["n_aligned_root"] == len(al) + int(((al["num_errors"] < 0)).sum())     # every root alignment of this workload succeeds
        # a second run of the staged job is bit-identical
        al2, cg2 = job.run().alignments()
        assert np.array_equal(al[["start_in_reference", "num_errors", "read_index", "orientation", "cigar_len"]],
                              al2[["start_in_reference", "num_errors", "read_index", "orientation", "cigar_len"]])
        job.free()
        # bit-for-bit against the CPU port on the first reads
        sample = batch.slice(0, 12)
        job = c2.verify_reads(sample, cfg)
        sal, scg = job.alignments()
        wal, wcg, wstats = cpu_baseline.verify_reads(refs, sample, cfg, threads=8)
        assert alignment_records(sal, scg) == alignment_records(wal, wcg) and job.stats() == wstats
        job.free()
    finally:
        c2.close()


def test_two_batches_in_flight(ctx, gpu, oracle):
    """Two jobs run from two threads at the same time (one per worker group) give what they give one after the other."""
    import threading
    refs = [synthetic.random_reference(120_000, 81)]
    ctx.set_references(refs)
    batches = [synthetic.make_batch(refs, 10, 1000, 0.06, s, gpu.pex_build, seed_errors=1, decoy_fraction=0.3) for s in (91, 92)]
    cfgs = [VerifyConfig(), VerifyConfig(interval_optimization=True)]
    want = [oracle_verify_batch(oracle, refs, b, c) for b, c in zip(batches, cfgs)]
    staged = [ctx.stage_verify(b, c) for b, c in zip(batches, cfgs)]
    errors = []

    def lane(i):
        try:
            for _ in range(4):
                al, cg = staged[i].run().alignments()
                assert alignment_records(al, cg) == want[i][0] and staged[i].stats() == want[i][1]
                j = ctx.verify_reads(batches[i], cfgs[i])           # the one-call path as well
                al, cg = j.alignments()
                assert alignment_records(al, cg) == want[i][0] and j.stats() == want[i][1]
                j.free()
        except Exception as e:                                      # noqa: BLE001 -- reported by the main thread
            errors.append((i, repr(e)))

    threads = [threading.Thread(target=lane, args=(i,)) for i in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for s in staged:
        s.free()


@pytest.mark.parametrize("ivopt", [False, True])
def test_merged_batches_match_the_oracle(ctx, gpu, oracle, ivopt):
    """24 callers at once: the queue merges the jobs that wait into batches (floxer_gpu.cu: submit_and_wait) -- different
    read sets, read lengths and tree shapes in one launch sequence -- and every caller must still get exactly its own
    alignments, cigars and statistics.  Staged jobs and the one-call path, several rounds each."""
    import threading
    refs = [synthetic.random_reference(150_000, 181), synthetic.plant_repeats(synthetic.random_reference(60_000, 182), 183, families=3, unit=(300, 700), copies=(3, 5))]
    ctx.set_references(refs)
    cfg = VerifyConfig(interval_optimization=ivopt)
    shapes = [(6, 900, 0.06, 1), (3, 1500, 0.05, 2), (9, 400, 0.08, 1), (2, 2400, 0.07, 2), (5, 700, 0.04, 1), (1, 3000, 0.06, 2)]
    batches = [synthetic.make_batch(refs, n, L, e, 300 + i, gpu.pex_build, seed_errors=se, decoy_fraction=0.4) for i, (n, L, e, se) in enumerate(shapes)]
    want = [oracle_verify_batch(oracle, refs, b, cfg) for b in batches]
    n_lanes = 24
    staged = [ctx.stage_verify(batches[i % len(batches)], cfg) for i in range(n_lanes)]
    errors = []
    ctx.reset_counters()
    start = threading.Barrier(n_lanes)

    def lane(i):
        try:
            w = want[i % len(batches)]
            start.wait()
            for rnd in range(3):
                al, cg = staged[i].run().alignments()
                assert alignment_records(al, cg) == w[0] and staged[i].stats() == w[1], ("staged", i, rnd)
                j = ctx.verify_reads(batches[i % len(batches)], cfg)
                al, cg = j.alignments()
                assert alignment_records(al, cg) == w[0] and j.stats() == w[1], ("one call", i, rnd)
                j.free()
        except Exception as e:                                      # noqa: BLE001 -- reported by the main thread
            errors.append((i, repr(e)))

    threads = [threading.Thread(target=lane, args=(i,)) for i in range(n_lanes)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:3]
    c = ctx.counters()
    assert c["batch_jobs"] == n_lanes * 3 * 2
    assert c["batches"] < c["batch_jobs"], "no two jobs ever shared a batch"
    for s in staged:
        s.free()


def test_verify_with_seeder_anchors(ctx, gpu, oracle):
    """Anchors from the q-gram seeder (row N2) instead of the ground truth: several anchors per locus a few bases apart,
    hits in both orientations, repeat copies -- through verify_reads against the oracle, both settings of the interval
    optimisation."""
    from floxer_b200 import workloads as W
    refs = [synthetic.random_reference(90_000, 71), synthetic.plant_repeats(synthetic.random_reference(50_000, 72), 73, families=4, unit=(300, 900), copies=(3, 5))]
    seeder = gpu.Seeder(refs, q=8)
    batch = W.make_reads_seeded(refs, 10, 800, 0.06, 74, gpu.pex_build, seeder, seed_errors=1, threads=2)
    seeder.close()
    assert len(batch.anchors) > 50
    ctx.set_references(refs)
    for cfg in (VerifyConfig(), VerifyConfig(interval_optimization=True)):
        job = ctx.verify_reads(batch, cfg)
        al, cg = job.alignments()
        want, want_stats = oracle_verify_batch(oracle, refs, batch, cfg)
        assert alignment_records(al, cg) == want and job.stats() == want_stats
        assert len(want) > 0
        job.free()


def test_resident_references_with_n_and_sentinel(ctx, gpu, oracle):
    """Ranks 0 ($) and 5 (N / invalid) inside RESIDENT references and inside reads, through verify_reads: comparison is plain
    byte equality (src/lib/alignment.cpp:14), so N matches N and $ matches $."""
    rng = np.random.default_rng(5150)
    refs = [synthetic.random_reference(40_000, 61), synthetic.random_reference(25_000, 62)]
    for r in refs:
        for at in rng.integers(0, len(r) - 40, size=60):
            r[at: at + int(rng.integers(1, 30))] = 5                  # runs of N
        r[rng.integers(0, len(r), size=40)] = 0                       # stray sentinels
    batch = synthetic.make_batch(refs, 12, 900, 0.06, 63, gpu.pex_build, seed_errors=1, decoy_fraction=0.3)
    assert (batch.forward_pool == 5).any() and (batch.forward_pool == 0).any()      # the reads carry them too (simulated from the references)
    ctx.set_references(refs)
    for cfg in (VerifyConfig(), VerifyConfig(interval_optimization=True), VerifyConfig(without_cigar=True)):
        job = ctx.verify_reads(batch, cfg)
        al, cg = job.alignments()
        want, want_stats = oracle_verify_batch(oracle, refs, batch, cfg)
        assert alignment_records(al, cg) == want and job.stats() == want_stats
        assert len(want) > 0
        job.free()


@pytest.mark.parametrize("name,n_reads,read_len,error,fp,ivopt", [
    ("config3", 8, 15_000, 0.08, 1.0, True), ("config3", 4, 15_000, 0.08, 1.0, False),
    ("config4", 8, 20_000, 0.10, 3.0, True), ("config4", 3, 20_000, 0.10, 3.0, False),
])
def test_config3_and_4_shapes_unthinned(ctx, gpu, name, n_reads, read_len, error, fp, ivopt):
    """Read shapes of configs 3 and 4 with EVERY anchor of the stand-in seeder (true loci with jitter, hits at repeat copies,
    false positives) on a repeat-seeded multi-record reference, with and without the interval optimisation, bit for bit
    against the CPU port (alignments, cigars, order, statistics)."""
    from floxer_b200 import workloads as W
    from oracle import cpu_baseline
    refs = [W.random_reference(900_000, 301, 2), W.random_reference(500_000, 302, 2)]
    table = W.plant_repeats(refs, 303, families=12, unit=(500, 5000), copies=(4, 9))
    batch = W.make_reads(refs, n_reads, read_len, error, 304 + int(ivopt), gpu.pex_build, repeats=table, false_positives=fp)
    assert len(batch.anchors) > 400 * n_reads
    ctx.set_references(refs)
    cfg = VerifyConfig(interval_optimization=ivopt)
    job = ctx.verify_reads(batch, cfg)
    al, cg = job.alignments()
    wal, wcg, wstats = cpu_baseline.verify_reads(refs, batch, cfg, threads=16)
    assert alignment_records(al, cg) == alignment_records(wal, wcg)
    assert job.stats() == wstats
    assert len(al) >= n_reads
    job.free()


def test_two_contexts_in_one_process(gpu, oracle):
    """The shape floxer itself needs (src/main/floxer.cpp:141-171 is ONE process): several contexts driven from the threads of
    one process -- one per GPU where the box has several, two on the same GPU otherwise -- each verifying its shard of the
    reads (floxer_b200.sharding.shard), the records gathered in read order and equal to one context's result for all reads."""
    import threading
    import torch
    from floxer_b200 import sharding
    n_ctx = max(2, min(4, torch.cuda.device_count()))
    devices = [i % torch.cuda.device_count() for i in range(n_ctx)]
    refs = [synthetic.random_reference(150_000, 91)]
    batch = synthetic.make_batch(refs, 23, 1100, 0.06, 92, gpu.pex_build, seed_errors=1, decoy_fraction=0.3)
    cfg = VerifyConfig(interval_optimization=True)
    want, want_stats = oracle_verify_batch(oracle, refs, batch, cfg)
    ctxs = [gpu.Context(d) for d in devices]
    out, errors = [None] * n_ctx, []

    def lane(r):
        try:
            ctxs[r].set_references(refs)
            part, first = sharding.shard(batch, r, n_ctx)
            job = ctxs[r].verify_reads(part, cfg)
            al, cg = job.alignments()
            out[r] = (sharding.gather_records(alignment_records(al, cg), first), job.stats())
            job.free()
        except Exception as e:                                      # noqa: BLE001 -- reported by the main thread
            errors.append((r, repr(e)))
    threads = [threading.Thread(target=lane, args=(r,)) for r in range(n_ctx)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for c in ctxs:
        c.close()
    assert not errors, errors
    assert [rec for part in out for rec in part[0]] == want
    assert {k: sum(part[1][k] for part in out) for k in want_stats} == want_stats


def test_verify_reads_rejects_bad_ranks(ctx, gpu):
    """A rank above 5 in a query pool is an error of fxg_verify_reads (found on the device, reported with the call)."""
    refs = [synthetic.random_reference(50_000, 5)]
    ctx.set_references(refs)
    batch = synthetic.make_batch(refs, 3, 600, 0.05, 6, gpu.pex_build)
    batch.reverse_pool = batch.reverse_pool.copy()
    batch.reverse_pool[100] = 7
    with pytest.raises(gpu.FloxerGpuError):
        ctx.verify_reads(batch, VerifyConfig())
    with pytest.raises(gpu.FloxerGpuError):
        ctx.stage_verify(batch, VerifyConfig())
    batch.reverse_pool[100] = 1
    assert len(ctx.verify_reads(batch, VerifyConfig()).alignments()[0]) >= 0


def test_randomised_shortcuts_match_plain(gpu):
    """scripts/stress_parity.py, 14 random batches (tiny and clipped references, mixed read lengths, every mode): the
    queue with every shortcut on against the queue that computes every window on its own.  Seed 1 / case 8 is the batch
    that exposed a missing final checkpoint record (two root windows 3 bases apart sharing one pass)."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("stress_parity", os.path.join(os.path.dirname(__file__), "..", "scripts", "stress_parity.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.main(14, 1, None, 0) == 0
