"""BAM output (row N3 of SURVEY 8f) on the CPU: fxg_write_bam is host-only code, so the file image it produces for a set
of alignment records is decoded here (BGZF blocks -> BAM header and records, SAM specification sections 4.1 / 4.2) and
compared, field by field, with the CPU restatement of src/lib/output.cpp:49-108 (oracle/sam_oracle.py) -- the same record
contents the SAM emitter is checked against on the GPU box."""
import gzip
import struct

import numpy as np
import pytest

from floxer_b200 import abi
from floxer_b200.batch import BatchBuilder
from oracle import oracle as o
from oracle import sam_oracle

NIBBLES = "=ACMGRSVTWYHKDBN"
OPS = {1: "I", 2: "D", 3: "N", 4: "S", 7: "=", 8: "X"}


@pytest.fixture(scope="module")
def gpu():
    from floxer_b200 import build, gpu as g
    build.build_native()
    g.lib()
    return g


def bgzf_blocks(data: bytes):
    """Splits a BGZF stream into its blocks and checks every block header; returns the uncompressed payloads."""
    out, at = [], 0
    while at < len(data):
        assert data[at:at + 4] == b"\x1f\x8b\x08\x04" and data[at + 10:at + 12] == b"\x06\x00" and data[at + 12:at + 16] == b"BC\x02\x00"
        bsize = struct.unpack_from("<H", data, at + 16)[0] + 1
        block = data[at:at + bsize]
        payload = gzip.decompress(block)
        assert struct.unpack_from("<I", block, bsize - 4)[0] == len(payload) <= 0xff00
        out.append(payload)
        at += bsize
    assert at == len(data)
    return out


def parse_bam(data: bytes, with_header=True):
    blocks = bgzf_blocks(data)
    assert blocks[-1] == b"" and data.endswith(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))   # the EOF marker block
    raw = b"".join(blocks)
    at, refs, text = 0, [], ""
    if with_header:
        assert raw[:4] == b"BAM\x01"
        l_text = struct.unpack_from("<I", raw, 4)[0]
        text = raw[8:8 + l_text].decode()
        at = 8 + l_text
        n_ref = struct.unpack_from("<I", raw, at)[0]
        at += 4
        for _ in range(n_ref):
            l_name = struct.unpack_from("<I", raw, at)[0]
            name = raw[at + 4:at + 4 + l_name - 1].decode()
            assert raw[at + 4 + l_name - 1] == 0
            refs.append((name, struct.unpack_from("<I", raw, at + 4 + l_name)[0]))
            at += 8 + l_name
    records = []
    while at < len(raw):
        size = struct.unpack_from("<I", raw, at)[0]
        ref_id, pos, l_name, mapq, bin_, n_cigar, flag, l_seq, nref, npos, tlen = struct.unpack_from("<iiBBHHHIiii", raw, at + 4)
        p = at + 36
        name = raw[p:p + l_name - 1].decode()
        p += l_name
        cigar = "".join(f"{v >> 4}{OPS[v & 15]}" for v in struct.unpack_from(f"<{n_cigar}I", raw, p)) or "*"
        span = sum(v >> 4 for v in struct.unpack_from(f"<{n_cigar}I", raw, p) if (v & 15) in (2, 7, 8))
        p += 4 * n_cigar
        packed = raw[p:p + (l_seq + 1) // 2]
        seq = "".join(NIBBLES[(packed[i // 2] >> (4 if i % 2 == 0 else 0)) & 15] for i in range(l_seq)) or "*"
        p += (l_seq + 1) // 2
        q = raw[p:p + l_seq]
        qual = "*" if (l_seq == 0 or all(b == 0xff for b in q)) else "".join(chr(b + 33) for b in q)
        p += l_seq
        nm = None
        while p < at + 4 + size:
            tag, typ = raw[p:p + 2].decode(), chr(raw[p + 2])
            if typ == "B":                                                    # the real CIGAR of a record with more than 65 535 operations
                assert tag == "CG" and chr(raw[p + 3]) == "I"
                count = struct.unpack_from("<I", raw, p + 4)[0]
                real = struct.unpack_from(f"<{count}I", raw, p + 8)
                assert cigar == f"{sum(v >> 4 for v in real if (v & 15) in (1, 7, 8))}S{sum(v >> 4 for v in real if (v & 15) in (2, 7, 8))}N"
                cigar = "".join(f"{v >> 4}{OPS[v & 15]}" for v in real)
                span = sum(v >> 4 for v in real if (v & 15) in (2, 7, 8))
                p += 8 + 4 * count
                continue
            width = {"C": 1, "S": 2, "I": 4}[typ]
            val = int.from_bytes(raw[p + 3:p + 3 + width], "little")
            p += 3 + width
            if tag == "NM":
                nm = val
        assert p == at + 4 + size and (nref, npos, tlen) == (-1, -1, 0)
        records.append(dict(name=name, flag=flag, ref_id=ref_id, pos=pos, mapq=mapq, cigar=cigar, seq=seq, qual=qual, nm=nm,
                            bin=bin_, span=span))
        at += 4 + size
    return text, refs, records


def reg2bin(beg, end):
    end -= 1
    for shift, base in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        if beg >> shift == end >> shift:
            return base + (beg >> shift)
    return 0


def random_case(rng, n_reads, n_refs):
    """A batch (only its reads and forward pool matter here) and alignment records with consistent CIGARs."""
    bb = BatchBuilder()
    node = np.zeros(1, dtype=abi.PEX_NODE_DTYPE)
    node["parent_id"] = abi.NULL_ID
    alignments, cigars, per_read = [], [], []
    for ri in range(n_reads):
        m = int(rng.integers(1, 70))
        fwd = rng.integers(0, 6, size=m, dtype=np.uint8)                # every rank, '$' and N included
        node["query_index_to"] = m - 1
        bb.add(fwd, fwd[::-1].copy(), node[:0], node, np.zeros(0, dtype=abi.ANCHOR_DTYPE), np.zeros(0, dtype=abi.ANCHOR_DTYPE))
        mine = []
        for _ in range(int(rng.integers(0, 5))):
            runs, left = [], m
            while left:
                op = int(rng.choice([7, 7, 8, 1, 2]))
                n = int(rng.integers(1, 9))
                if op != 2:
                    n = min(n, left)
                    left -= n
                if runs and runs[-1][1] == op:
                    runs[-1][0] += n
                else:
                    runs.append([n, op])
            ops = [(n << 4) | op for n, op in runs]
            errors = sum(n for n, op in runs if op != 7)
            big = rng.random() < 0.1
            start = int(rng.integers(2**31 - 3, 2**33)) if big else int(rng.integers(0, 1 << 20))
            mine.append((int(rng.integers(0, n_refs)), start, errors if rng.random() < 0.8 else errors + int(rng.integers(250, 70000)),
                         int(rng.integers(0, 2)), ops))
        per_read.append(mine)
        for ref, start, errors, orient, ops in mine:
            alignments.append((start, len(cigars), len(ops), errors, ri, ref, orient, (0,) * 7))
            cigars += ops
    al = np.array(alignments, dtype=abi.ALIGNMENT_DTYPE) if alignments else np.zeros(0, dtype=abi.ALIGNMENT_DTYPE)
    return bb.build(), al, np.array(cigars, dtype=np.uint32), per_read


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_bam_records_match_the_restatement_of_output_cpp(gpu, seed):
    rng = np.random.default_rng(seed)
    n_refs = 3
    ref_ids = ["chr1", "second reference", "r3"]
    ref_lens = [1 << 21, 2**32 + 5, 77]
    batch, al, cg, per_read = random_case(rng, 40, n_refs)
    names = [f"read/{i}" for i in range(len(batch.reads))]
    quals = ["".join(chr(33 + int(x)) for x in rng.integers(0, 42, size=int(R["query_len"]))) if i % 3 else "" for i, R in enumerate(batch.reads)]
    data = gpu.write_bam(al, cg, batch, ref_ids, ref_lens, names, quals)
    text, refs, records = parse_bam(data)
    assert text.splitlines()[0].startswith("@HD") and text.splitlines()[1:] == [f"@SQ\tSN:{n}\tLN:{l}" for n, l in zip(ref_ids, ref_lens)]
    assert refs == [(n, min(l, 2**31 - 1)) for n, l in zip(ref_ids, ref_lens)]
    want = []
    for ri, R in enumerate(batch.reads):
        qo, ql = int(R["query_offset"]), int(R["query_len"])
        want += sam_oracle.sam_records(names[ri], batch.forward_pool[qo:qo + ql], quals[ri],
                                       [(ref, start, errors, orient, o.cigar_to_string(ops)) for ref, start, errors, orient, ops in per_read[ri]], ref_ids)
    assert len(records) == len(want)
    for got, (qname, flag, rname, pos1, mapq, cigar, seq, qual, nm) in zip(records, want):
        assert (got["name"], got["flag"], got["mapq"], got["cigar"], got["nm"]) == (qname, flag, mapq, cigar, nm)
        assert got["seq"] == seq.replace("$", "N") and got["qual"] == qual          # '$' has no BAM code: written as N
        if flag & 4:
            assert (got["ref_id"], got["pos"], got["bin"]) == (-1, -1, 4680)
        else:
            assert ref_ids[got["ref_id"]] == rname and got["pos"] + 1 == pos1
            end = got["pos"] + max(got["span"], 1)
            assert got["bin"] == (reg2bin(got["pos"], end) if end <= 1 << 29 else 4680)
    # without the header the same records follow directly
    _, _, again = parse_bam(gpu.write_bam(al, cg, batch, ref_ids, ref_lens, names, quals, header=False), with_header=False)
    assert again == records


def test_bam_many_blocks_and_bad_input(gpu):
    rng = np.random.default_rng(9)
    batch, al, cg, _ = random_case(rng, 4000, 2)                             # several BGZF blocks
    names = [f"q{i}" for i in range(len(batch.reads))]
    data = gpu.write_bam(al, cg, batch, ["a", "b"], [10, 20], names)
    assert len(bgzf_blocks(data)) > 3
    _, _, records = parse_bam(data)
    assert len(records) == sum(max(1, int((al["read_index"] == i).sum())) for i in range(len(batch.reads)))
    bad = al.copy()
    if len(bad) > 1:
        bad["read_index"][0], bad["read_index"][-1] = bad["read_index"][-1], bad["read_index"][0]     # not grouped by read any more
        with pytest.raises(gpu.FloxerGpuError):
            gpu.write_bam(bad, cg, batch, ["a", "b"], [10, 20], names)
    with pytest.raises(gpu.FloxerGpuError):
        gpu.write_bam(al, cg, batch, ["a"], [10], names)                     # an alignment names reference 1
    with pytest.raises(gpu.FloxerGpuError):
        gpu.write_bam(al, cg, batch, ["a", "b"], [10, 20], ["x" * 300] + names[1:])


def test_bam_cigar_with_more_than_65535_operations(gpu):
    """n_cigar_op is 16 bits wide: a longer CIGAR travels in the CG:B,I tag behind the placeholder <query>S<reference>N
    (SAM specification 4.2.2) -- a 100 kbp read at 10 % errors can have one."""
    m = 70_001
    bb = BatchBuilder()
    node = np.zeros(1, dtype=abi.PEX_NODE_DTYPE)
    node["parent_id"] = abi.NULL_ID
    node["query_index_to"] = m - 1
    fwd = (np.arange(m) % 4 + 1).astype(np.uint8)
    bb.add(fwd, fwd[::-1].copy(), node[:0], node, np.zeros(0, dtype=abi.ANCHOR_DTYPE), np.zeros(0, dtype=abi.ANCHOR_DTYPE))
    ops = np.array([(1 << 4) | (7 if i % 2 == 0 else 8) for i in range(m)], dtype=np.uint32)     # 1=1X1=1X...: 70 001 operations
    al = np.array([(1234, 0, len(ops), m // 2, 0, 0, 0, (0,) * 7), (99, 0, 3, 5, 0, 0, 1, (0,) * 7)], dtype=abi.ALIGNMENT_DTYPE)
    data = gpu.write_bam(al, ops, bb.build(), ["chr"], [1 << 20], ["long"])
    _, _, records = parse_bam(data)
    assert len(records) == 2
    assert records[0]["cigar"] == "".join(f"1{'=' if i % 2 == 0 else 'X'}" for i in range(m)) and records[0]["span"] == m
    assert records[0]["pos"] == 1234 and records[0]["bin"] == reg2bin(1234, 1234 + m) and records[0]["flag"] == 256 and records[0]["nm"] == m // 2
    assert records[1]["cigar"] == "1=1X1=" and records[1]["flag"] == 16          # fewer errors: the primary record
