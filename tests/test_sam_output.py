"""SAM records of a verified job (row N3 of SURVEY 8f): the host library's emitter against the CPU restatement of
src/lib/output.cpp:49-108 and against what test/floxer_whole_program_via_cli_test.cpp:38-93 expects of the output file."""
import numpy as np
import pytest

import golden_vectors as G
from floxer_b200 import abi, synthetic
from floxer_b200.batch import BatchBuilder, VerifyConfig, alignment_records
from harness import brute_force_anchors, revcomp, to_ranks
from oracle import sam_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    from floxer_b200 import build, gpu as g
    build.build_native()
    g.lib()
    return g


def _expected_records(batch, recs, query_ids, qualities, reference_ids):
    by_read = {}
    for r in recs:
        by_read.setdefault(r[0], []).append((r[1], r[2], r[3], r[4], r[5]))
    out = []
    for ri, R in enumerate(batch.reads):
        qo, ql = int(R["query_offset"]), int(R["query_len"])
        out += sam_oracle.sam_records(query_ids[ri], batch.forward_pool[qo:qo + ql], qualities[ri], by_read.get(ri, []), reference_ids)
    return out


@pytest.mark.parametrize("seed_errors", G.WHOLE_FLAGS["seed_errors"])
def test_whole_program_fixture_sam(gpu, seed_errors):
    refs = [to_ranks(s) for s in G.WHOLE_REFERENCES.values()]
    ref_ids = list(G.WHOLE_REFERENCES)
    ctx = gpu.Context(0)
    try:
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
        quals = ["I" * len(G.WHOLE_QUERIES[q]) for q in names]
        text = job.sam(batch, ref_ids, [len(r) for r in refs], names, quals)
        header, records = sam_oracle.parse_sam(text)
        assert header[0].startswith("@HD") and header[1:] == [f"@SQ\tSN:{n}\tLN:{len(r)}" for n, r in zip(ref_ids, refs)]
        # the emitter agrees with the restatement of output.cpp on the job's own alignments
        assert records == _expected_records(batch, alignment_records(*job.alignments()), names, quals, ref_ids)
        # ... and the records are what the reference's whole-program test expects to read back
        assert {r[0] for r in records} == set(names)
        for rec in records:
            qid, flag, rname, pos1, mapq, cigar, seq, qual, nm = rec
            if qid in G.WHOLE_UNMAPPED:
                assert flag == 4
                continue
            assert not flag & 4 and mapq == 255
            lo, hi, want_nm, want_cigar = G.WHOLE_EXPECT[(qid, bool(flag & 16))]
            assert lo <= pos1 - 1 <= hi and nm == want_nm and cigar == want_cigar
        for qid in names:
            mine = [r for r in records if r[0] == qid]
            if qid in G.WHOLE_UNMAPPED:
                assert len(mine) == 1
                continue
            primaries = [r for r in mine if not r[1] & 256]
            assert len(primaries) == 1 and primaries[0][6] == G.WHOLE_QUERIES[qid].upper() and primaries[0][8] == min(r[8] for r in mine)
            assert all(r[6] == "*" and r[7] == "*" for r in mine if r[1] & 256)
        job.free()
    finally:
        ctx.close()


def test_sam_on_a_simulated_batch(gpu):
    refs = [synthetic.random_reference(90_000, 71), synthetic.random_reference(40_000, 72)]
    ctx = gpu.Context(0)
    try:
        ctx.set_references(refs)
        batch = synthetic.make_batch(refs, 9, 800, 0.06, 19, gpu.pex_build, seed_errors=1, decoy_fraction=0.3)
        names = [f"read{i}" for i in range(len(batch))]
        quals = ["" if i % 3 == 0 else "F" * int(batch.reads[i]["query_len"]) for i in range(len(batch))]
        for cfg in (VerifyConfig(), VerifyConfig(interval_optimization=True), VerifyConfig(without_cigar=True)):
            job = ctx.verify_reads(batch, cfg)
            text = job.sam(batch, ["chrA", "chrB"], [len(r) for r in refs], names, quals, header=False)
            header, records = sam_oracle.parse_sam(text)
            assert not header
            assert records == _expected_records(batch, alignment_records(*job.alignments()), names, quals, ["chrA", "chrB"])
            assert sum(1 for r in records if not r[1] & (256 | 4)) == len({r[0] for r in records if not r[1] & 4})
            # the BAM image of the same job holds the same records
            from test_bam_output import parse_bam
            _, bam_refs, bam_records = parse_bam(job.bam(batch, ["chrA", "chrB"], [len(r) for r in refs], names, quals))
            assert bam_refs == [("chrA", len(refs[0])), ("chrB", len(refs[1]))]
            assert [(b["name"], b["flag"], "*" if b["ref_id"] < 0 else ["chrA", "chrB"][b["ref_id"]], b["pos"] + 1, b["mapq"], b["cigar"], b["seq"], b["qual"], b["nm"])
                    for b in bam_records] == records
            job.free()
    finally:
        ctx.close()
