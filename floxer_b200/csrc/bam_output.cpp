// bam_output.cpp -- BAM (BGZF-compressed) records for verified alignments: row N3 of SURVEY 8f in the container the
// reference actually writes (seqan3::sam_file_output on a .bam path, src/lib/output.cpp:197-212).
//
// Record contents and order are those of sam_output.cpp (output::alignment_output::write_alignments_for_query,
// src/lib/output.cpp:49-108): per query the alignments reference by reference in insertion order, the first one with the
// query's best number of errors primary (SEQ and QUAL present), the others secondary (flag | 256, no SEQ / QUAL), MAPQ 255,
// NM tag, one flag-4 record for a query without alignments.  The binary layout follows the SAM/BAM specification
// (section 4.2; BGZF section 4.1): the CIGAR operations of the C ABI are already BAM-encoded (len << 4 | op with
// I = 1, D = 2, '=' = 7, X = 8).  Byte-level parity with SeqAn3's writer is unpinned (its source is not on disk; optional
// choices such as the integer type of NM or the compression level are the specification's smallest / zlib's default).
//
// Host-only code: the array entry point needs neither a context nor a GPU, which is how tests/test_bam_output.py drives it.
#include "../../include/floxer_gpu.h"

#include <zlib.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

namespace {

void put_u16(std::vector<uint8_t>& b, uint32_t v) { b.push_back(uint8_t(v)); b.push_back(uint8_t(v >> 8)); }
void put_u32(std::vector<uint8_t>& b, uint32_t v) { for (int i = 0; i < 4; ++i) b.push_back(uint8_t(v >> (8 * i))); }
void put_i32(std::vector<uint8_t>& b, int32_t v) { put_u32(b, uint32_t(v)); }
void put_bytes(std::vector<uint8_t>& b, const void* p, size_t n) { auto q = static_cast<const uint8_t*>(p); b.insert(b.end(), q, q + n); }

// bin of the half-open, 0-based interval [beg, end) (SAM specification, section 5.3)
uint32_t reg2bin(int64_t beg, int64_t end) {
    --end;
    if (beg >> 14 == end >> 14) return uint32_t(((1 << 15) - 1) / 7 + (beg >> 14));
    if (beg >> 17 == end >> 17) return uint32_t(((1 << 12) - 1) / 7 + (beg >> 17));
    if (beg >> 20 == end >> 20) return uint32_t(((1 << 9) - 1) / 7 + (beg >> 20));
    if (beg >> 23 == end >> 23) return uint32_t(((1 << 6) - 1) / 7 + (beg >> 23));
    if (beg >> 26 == end >> 26) return uint32_t(((1 << 3) - 1) / 7 + (beg >> 26));
    return 0;
}

// ranks 0..5 = $ A C G T N (src/lib/input.cpp:165-176) as the nibbles of "=ACMGRSVTWYHKDBN"; '$' has no code: N
constexpr uint8_t kRankToNibble[6] = {15, 1, 2, 4, 8, 15};

// one BGZF block per <= 0xff00 bytes of input, then the empty end-of-file block
int bgzf_compress(std::vector<uint8_t> const& in, std::vector<uint8_t>& out) {
    size_t const kChunk = 0xff00;
    std::vector<uint8_t> buf(compressBound(kChunk) + 64);
    for (size_t at = 0; at < in.size(); at += kChunk) {
        size_t const n = std::min(kChunk, in.size() - at);
        z_stream zs{};
        if (deflateInit2(&zs, Z_DEFAULT_COMPRESSION, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return FXG_ERR_OUT_OF_MEMORY;
        zs.next_in = const_cast<Bytef*>(in.data() + at); zs.avail_in = uInt(n);
        zs.next_out = buf.data(); zs.avail_out = uInt(buf.size());
        int const rc = deflate(&zs, Z_FINISH);
        size_t const clen = buf.size() - zs.avail_out;
        deflateEnd(&zs);
        if (rc != Z_STREAM_END || clen + 26 > 65536) return FXG_ERR_OVERFLOW;
        static const uint8_t head[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
        put_bytes(out, head, 16);
        put_u16(out, uint32_t(clen + 25));                         // BSIZE: total block size - 1
        put_bytes(out, buf.data(), clen);
        put_u32(out, uint32_t(crc32(crc32(0L, Z_NULL, 0), in.data() + at, uInt(n))));
        put_u32(out, uint32_t(n));
    }
    static const uint8_t eof[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    put_bytes(out, eof, 28);
    return FXG_OK;
}

int append_record(std::vector<uint8_t>& out, const char* qname, int32_t ref_id, int32_t pos, uint32_t flag, const uint32_t* ops, uint32_t n_ops,
                  const uint8_t* seq_ranks, uint32_t l_seq, const char* qual, bool with_nm, uint32_t nm) {
    size_t const l_name = std::strlen(qname) + 1;
    if (l_name > 255) return FXG_ERR_INVALID_ARGUMENT;             // read names hold at most 254 characters
    int64_t ref_span = 0, query_span = 0;
    for (uint32_t i = 0; i < n_ops; ++i) {
        uint32_t const op = ops[i] & 15u;
        if (op == FXG_CIGAR_D || op == FXG_CIGAR_EQ || op == FXG_CIGAR_X) ref_span += ops[i] >> 4;
        if (op == FXG_CIGAR_I || op == FXG_CIGAR_EQ || op == FXG_CIGAR_X) query_span += ops[i] >> 4;
    }
    // n_cigar_op is 16 bits wide: a longer CIGAR (a 100 kbp read at 10 % errors can have one) goes into the CG:B,I tag and
    // the record carries the placeholder <query length>S<reference length>N (SAM specification, section 4.2.2)
    bool const long_cigar = n_ops > 65535;
    if (long_cigar && (query_span >= (int64_t(1) << 28) || ref_span >= (int64_t(1) << 28))) return FXG_ERR_OVERFLOW;
    uint32_t const placeholder[2] = {uint32_t(query_span) << 4 | 4u /* S */, uint32_t(ref_span) << 4 | 3u /* N */};
    uint32_t const n_written = long_cigar ? 2u : n_ops;
    const uint32_t* const written = long_cigar ? placeholder : ops;
    size_t const size_at = out.size();
    put_u32(out, 0);                                               // block_size, patched below
    put_i32(out, ref_id); put_i32(out, pos);
    out.push_back(uint8_t(l_name)); out.push_back(255);            // MAPQ 255: not available (output.cpp:76)
    // (the binning scheme covers coordinates below 2^29; beyond it -- and for unmapped records -- the bin of "no position")
    int64_t const end = int64_t(pos) + std::max<int64_t>(ref_span, 1);
    put_u16(out, pos < 0 || end > (int64_t(1) << 29) ? 4680u : reg2bin(pos, end));
    put_u16(out, n_written); put_u16(out, flag); put_u32(out, l_seq);
    put_i32(out, -1); put_i32(out, -1); put_i32(out, 0);           // no mate
    put_bytes(out, qname, l_name);
    for (uint32_t i = 0; i < n_written; ++i) put_u32(out, written[i]);
    for (uint32_t p = 0; p < l_seq; p += 2) {
        uint8_t const hi = kRankToNibble[std::min<uint8_t>(seq_ranks[p], 5)];
        uint8_t const lo = p + 1 < l_seq ? kRankToNibble[std::min<uint8_t>(seq_ranks[p + 1], 5)] : 0;
        out.push_back(uint8_t(hi << 4 | lo));
    }
    bool const have_qual = qual && qual[0] && std::strlen(qual) == l_seq;
    for (uint32_t p = 0; p < l_seq; ++p) out.push_back(have_qual ? uint8_t(qual[p] - 33) : uint8_t(0xff));
    if (with_nm) {
        out.push_back('N'); out.push_back('M');
        if (nm < 256) { out.push_back('C'); out.push_back(uint8_t(nm)); }
        else if (nm < 65536) { out.push_back('S'); put_u16(out, nm); }
        else { out.push_back('I'); put_u32(out, nm); }
    }
    if (long_cigar) {
        out.push_back('C'); out.push_back('G'); out.push_back('B'); out.push_back('I');
        put_u32(out, n_ops);
        for (uint32_t i = 0; i < n_ops; ++i) put_u32(out, ops[i]);
    }
    uint32_t const block_size = uint32_t(out.size() - size_at - 4);
    for (int i = 0; i < 4; ++i) out[size_at + size_t(i)] = uint8_t(block_size >> (8 * i));
    return FXG_OK;
}

}  // namespace

extern "C" {

int fxg_write_bam(const fxg_alignment* al, size_t n_al, const uint32_t* cigar_pool, size_t n_references, const char* const* reference_ids,
                  const uint64_t* reference_lengths, const fxg_read* reads, size_t n_reads, const uint8_t* forward_pool,
                  const fxg_sam_query* queries, int with_header, uint8_t** bytes, size_t* bytes_len) {
    if (!bytes || !bytes_len || (n_al && !al) || (n_reads && (!reads || !forward_pool || !queries)) || (n_references && (!reference_ids || !reference_lengths)))
        return FXG_ERR_INVALID_ARGUMENT;
    *bytes = nullptr; *bytes_len = 0;
    try {
    std::vector<uint8_t> raw;
    raw.reserve(size_t(1) << 20);
    if (with_header) {
        std::string text = "@HD\tVN:1.6\tSO:unknown\tGO:none\n";
        for (size_t r = 0; r < n_references; ++r) { text += "@SQ\tSN:"; text += reference_ids[r]; text += "\tLN:"; text += std::to_string(reference_lengths[r]); text += '\n'; }
        put_bytes(raw, "BAM\1", 4);
        put_u32(raw, uint32_t(text.size())); put_bytes(raw, text.data(), text.size());
        put_u32(raw, uint32_t(n_references));
        for (size_t r = 0; r < n_references; ++r) {
            size_t const l = std::strlen(reference_ids[r]) + 1;
            put_u32(raw, uint32_t(l)); put_bytes(raw, reference_ids[r], l);
            put_u32(raw, uint32_t(std::min<uint64_t>(reference_lengths[r], uint64_t(std::numeric_limits<int32_t>::max()))));
        }
    }
    std::vector<uint32_t> order;
    size_t a = 0;
    for (size_t ri = 0; ri < n_reads; ++ri) {
        size_t const a0 = a;
        while (a < n_al && al[a].read_index == ri) ++a;
        if (a < n_al && al[a].read_index < ri) return FXG_ERR_STATE;                 // alignments must be grouped by read
        const char* const qname = queries[ri].id ? queries[ri].id : "*";
        const uint8_t* const seq = forward_pool + reads[ri].query_offset;
        uint32_t const l_seq = reads[ri].query_len;
        if (a == a0) {                                                               // output.cpp:95-107: unmapped
            int const rc = append_record(raw, qname, -1, -1, 4, nullptr, 0, seq, l_seq, queries[ri].quality, false, 0);
            if (rc != FXG_OK) return rc;
            continue;
        }
        order.resize(a - a0);
        for (size_t k = 0; k < a - a0; ++k) order[k] = uint32_t(a0 + k);
        std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return al[x].reference_id < al[y].reference_id; });
        uint32_t best = std::numeric_limits<uint32_t>::max();
        for (size_t k = a0; k < a; ++k) best = std::min(best, al[k].num_errors);
        bool primary_written = false;
        for (uint32_t k : order) {
            fxg_alignment const& A = al[k];
            if (A.reference_id >= n_references) return FXG_ERR_INVALID_ARGUMENT;
            if (A.cigar_len && !cigar_pool) return FXG_ERR_INVALID_ARGUMENT;
            uint32_t flag = A.orientation == FXG_REVERSE_COMPLEMENT ? 16u : 0u;
            bool const primary = !primary_written && A.num_errors == best;
            if (primary) primary_written = true; else flag |= 256u;
            int32_t const pos = int32_t(std::min<uint64_t>(A.start_in_reference, uint64_t(std::numeric_limits<int32_t>::max())));   // math.hpp:10-16
            int const rc = append_record(raw, qname, int32_t(A.reference_id), pos, flag, A.cigar_len ? cigar_pool + A.cigar_offset : nullptr, A.cigar_len,
                                         seq, primary ? l_seq : 0, primary ? queries[ri].quality : nullptr, true, A.num_errors);
            if (rc != FXG_OK) return rc;
        }
    }
    if (a != n_al) return FXG_ERR_STATE;
    std::vector<uint8_t> packed;
    packed.reserve(raw.size() / 3 + 64);
    int const rc = bgzf_compress(raw, packed);
    if (rc != FXG_OK) return rc;
    uint8_t* buf = static_cast<uint8_t*>(std::malloc(packed.size() ? packed.size() : 1));
    if (!buf) return FXG_ERR_OUT_OF_MEMORY;
    std::memcpy(buf, packed.data(), packed.size());
    *bytes = buf; *bytes_len = packed.size();
    return FXG_OK;
    } catch (...) { return FXG_ERR_OUT_OF_MEMORY; }              // (std::bad_alloc / length_error of the buffers: nothing crosses the C ABI)
}

int fxg_job_write_bam(const fxg_job* job, size_t n_references, const char* const* reference_ids, const uint64_t* reference_lengths,
                      const fxg_read* reads, size_t n_reads, const uint8_t* forward_pool, const fxg_sam_query* queries,
                      int with_header, uint8_t** bytes, size_t* bytes_len) {
    if (!job) return FXG_ERR_INVALID_ARGUMENT;
    return fxg_write_bam(fxg_job_alignments(job), fxg_job_num_alignments(job), fxg_job_cigar_pool(job), n_references, reference_ids,
                         reference_lengths, reads, n_reads, forward_pool, queries, with_header, bytes, bytes_len);
}

}  // extern "C"
