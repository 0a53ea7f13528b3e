"""CPU restatement of output::alignment_output::write_alignments_for_query (src/lib/output.cpp:49-108 of the reference)
at the level of SAM record FIELDS.  Test infrastructure only: nothing in floxer_b200/ imports this.

The bytes of the file are SeqAn3's in the reference (seqan3::sam_file_output, not on disk here); what floxer itself decides --
record order, primary / secondary flags, MAPQ 255, NM, SEQ and QUAL only on the primary record, the unmapped record -- is
restated here and pinned by test/floxer_whole_program_via_cli_test.cpp:38-93."""
from __future__ import annotations

RANK_TO_CHAR = "$ACGTN"      # ivs::d_dna5 ranks as produced by src/lib/input.cpp:165-176
INT32_MAX = 2**31 - 1


def sam_records(query_id: str, forward_ranks, quality: str, alignments, reference_ids):
    """alignments: [(reference_id, start_in_reference, num_errors, orientation, cigar_string)] in insertion order
    (query_alignments::insert, src/lib/alignment.cpp:37-46).  Returns the record tuples
    (qname, flag, rname, pos1, mapq, cigar, seq, qual, nm_or_None) in output order."""
    seq = "".join(RANK_TO_CHAR[min(int(r), 5)] for r in forward_ranks) or "*"
    qual = quality or "*"
    if not alignments:
        return [(query_id, 4, "*", 0, 255, "*", seq, qual, None)]                       # output.cpp:95-107
    best = min(a[2] for a in alignments)                                               # query_alignments::best_num_errors
    out, primary_written = [], False
    for ref in sorted({a[0] for a in alignments}):                                     # output.cpp:57-59: reference by reference
        for a in alignments:
            if a[0] != ref:
                continue
            flag = 16 if a[3] else 0
            primary = (not primary_written) and a[2] == best                           # output.cpp:66-67
            if primary:
                primary_written = True
            else:
                flag |= 256
            out.append((query_id, flag, reference_ids[ref], min(int(a[1]), INT32_MAX) + 1, 255, a[4] or "*",
                        seq if primary else "*", qual if primary else "*", int(a[2])))
    return out


def parse_sam(text: str):
    """(header lines, record tuples as above) of SAM text."""
    header, records = [], []
    for line in text.splitlines():
        if line.startswith("@"):
            header.append(line)
            continue
        f = line.split("\t")
        nm = None
        for tag in f[11:]:
            if tag.startswith("NM:i:"):
                nm = int(tag[5:])
        assert f[6] == "*" and f[7] == "0" and f[8] == "0"
        records.append((f[0], int(f[1]), f[2], int(f[3]), int(f[4]), f[5], f[9], f[10], nm))
    return header, records
