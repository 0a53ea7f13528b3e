// root_kernels.cuh -- the root level of a batch on the device (query_verifier's root alignment, verification.cpp:52-64,
// 95-109; alignment.cpp:115-180): from the walks that verify their root (RootEntry list, dp_kernels.cuh) to finished
// alignment records, without the host touching a single walk.
//
//   root_prepare   window of every root walk (compute_reference_span_start_and_length with the extra length), statistics,
//                  sort key (read, strand | window start)
//   [sort]         windows of one read and strand next to each other, by start
//   root_cluster   windows of one query piece that nearly coincide form a UNIT: one checkpointed score pass over their
//                  union serves them all (the shared score passes of floxer_gpu.cu / DESIGN.md section 5)
//   root_units     one record per unit: union window, band, class, checkpoint words
//   root_tasks     the engine's tasks, per configuration class
//   [engine]       dp_kernel<W, true>  (or <W, false> on reversed views when no CIGAR is wanted)
//   root_results   per member: minimum of the last row over its own columns (read off the unit's records), accepted?,
//                  can the shared pass vouch for it?
//   [sort]         accepted alignments of one read and strand by end position
//   root_dedup     alignments that end at the same position with the same score share one traceback
//   root_walks     the traceback tasks of the representatives, cigar slots
//   [walk2]        tracebacks
//   root_finish    alignment records in anchor order
//
// A member whose result the shared pass cannot vouch for is counted; the host then runs the part's root level the slow
// way (run_root_passes), which scores such members again on their own -- rare (windows clipped by a reference's end).
#pragma once

#include "dp_kernels.cuh"

namespace fxg {

constexpr int kPosBits = 36;                                   // store positions and pool positions are below 2^36 on this path
constexpr uint64_t kPosMask = (uint64_t(1) << kPosBits) - 1;

struct RootClass { uint32_t W, G; };                           // the configuration classes of the context (block width, ring size)

struct UnitRec {
    uint64_t ws;                // store position of the union window's first base
    uint64_t qbase;             // pool position of the query piece
    uint64_t ck_base;           // first word of its checkpoint records
    uint32_t n, m, k;           // union window length, piece length, errors allowed
    uint32_t first, count;      // its members: sorted positions first .. first + count - 1
    uint32_t cls;               // configuration class
    uint32_t words;             // checkpoint words (a multiple of 4)
    uint32_t pair;              // read * 2 + strand (part-relative)
};

struct RootCtx {
    const RootEntry* entries; uint32_t n_roots;
    const ReadRec* reads; uint32_t n_reads;
    const uint64_t* ref_base; const uint64_t* ref_len;
    RootClass classes[kMaxLevelClasses];
    uint32_t want_cigar, share;
    // per root walk, original (anchor) order
    uint64_t* ws; uint32_t* len; uint32_t* read; uint8_t* orient;
    uint64_t* key; uint32_t* idx;                              // sort input: key, original index
    // per sorted position
    const uint64_t* key_s; const uint32_t* idx_s;
    uint32_t* unit_start;                                      // 1 = this member begins a unit
    const uint32_t* unit_of;                                   // inclusive scan of unit_start (unit id + 1)
    uint32_t* pos_of;                                          // original index -> sorted position
    int32_t* m_score; uint32_t* m_end; uint8_t* m_flag;        // member result: score, end column in unit coordinates, kMember* flags
    // units
    UnitRec* units; uint64_t* unit_words;                      // (unit_words: scan input)
    const uint64_t* unit_ck;                                   // exclusive scan of unit_words, in words
    DpTask* tasks; const DpResult* results;                    // engine tasks per class, results per unit
    const uint32_t* ck;                                        // checkpoint records
    // dedup / tracebacks
    uint64_t* key2; uint32_t* idx2; const uint64_t* key2_s; const uint32_t* idx2_s;
    uint32_t* rep;                                             // per sorted position: sorted position of the member whose traceback it shares
    uint64_t* tb_cap;                                          // per sorted position: cigar slots of its own traceback (0: none) -- scan input
    const uint64_t* tb_cig;                                    // exclusive scan of tb_cap
    uint32_t* tb_slot;                                         // per sorted position: index of its traceback task
    Walk2Task* wtasks; const WalkResult* wresults;
    // output
    uint32_t* hit;                                             // per original index: 1 = an alignment -- scan input
    const uint32_t* hit_at;                                    // exclusive scan of hit
    ::fxg_alignment* out;
    // counters (kCtr*): feasible roots, units, checkpoint words, members that need the slow way, accepted, tracebacks, errors,
    // units / tasks filled / word-steps per class, tracebacks / filled per block width, cigar slots, word-steps
    uint32_t* counters;
    unsigned long long* member_totals;
    uint32_t read0;                                            // first read of the part within the batch
};
enum : uint8_t { kMemberAccepted = 1, kMemberSafe = 2, kMemberMulti = 4 };
enum { kCtrFeasible = 0, kCtrUnits = 1, kCtrCkWords = 2 /* 64-bit */, kCtrSlow = 4, kCtrAccepted = 5, kCtrTracebacks = 6, kCtrErrors = 7,
       kCtrClassUnits = 8, kCtrClassFill = kCtrClassUnits + kMaxLevelClasses, kCtrWidthTb = kCtrClassFill + kMaxLevelClasses, kCtrWidthFill = kCtrWidthTb + 6,
       kCtrCigars = kCtrWidthFill + 6 /* 64-bit */, kCtrWordSteps = kCtrCigars + 2 /* 64-bit */, kCtrShared = kCtrWordSteps + 2,
       kCtrClassWs = kCtrShared + 4 /* kMaxLevelClasses x 64-bit */, kRootCounters = kCtrClassWs + 2 * kMaxLevelClasses };
static_assert(kCtrCkWords % 2 == 0 && kCtrCigars % 2 == 0 && kCtrWordSteps % 2 == 0 && kCtrClassWs % 2 == 0, "64-bit counters sit on 8-byte boundaries");

// ---- windows, statistics, sort keys ----
__global__ void root_prepare_kernel(RootCtx const C) {
    uint32_t const q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= C.n_roots) return;
    RootEntry const e = C.entries[q];
    uint32_t const r = read_of_walk(C.reads, C.n_reads, e.walk);
    ReadRec const R = C.reads[r];
    uint32_t const orient = e.walk - R.walk_begin >= R.n_forward ? 1u : 0u;
    uint64_t start, len;
    root_window(R, e.diag, C.ref_len[e.ref_id], start, len);
    uint64_t const ws = C.ref_base[e.ref_id] + start;
    C.ws[q] = ws; C.len[q] = uint32_t(len); C.read[q] = r; C.orient[q] = uint8_t(orient);
    // statistics of the root alignment (verification.cpp:238-242): counted whether or not a pass is needed
    warp_add_keyed(C.member_totals + 5, kMemberTotals, R.member, 1ull);
    warp_add_keyed(C.member_totals + 6, kMemberTotals, R.member, (unsigned long long)len);
    warp_add_keyed(C.member_totals + 7, kMemberTotals, R.member, (unsigned long long)R.root_m * len);
    bool const feasible = R.root_m != 0 && int64_t(R.root_m) - int64_t(len) <= int64_t(R.root_k);   // else more insertions needed than errors allowed
    C.key[q] = feasible ? ((uint64_t(r * 2 + orient) << kPosBits) | ws) : ~uint64_t(0);
    C.idx[q] = q;
    if (feasible) (void)warp_slot(C.counters, kCtrFeasible);
}

// ---- units: greedy clustering of a read and strand's windows, sorted by start ----
__device__ __forceinline__ double unit_cost(uint64_t n, uint64_t m, uint64_t k) { return double(n) * double(int64_t(n) - int64_t(m) + 2 * int64_t(k) + 256); }

__global__ void root_cluster_kernel(RootCtx const C) {
    uint32_t const n_feasible = C.counters[kCtrFeasible];
    uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_feasible) return;
    uint64_t const pair = C.key_s[i] >> kPosBits;
    C.pos_of[C.idx_s[i]] = i;
    if (i > 0 && (C.key_s[i - 1] >> kPosBits) == pair) return;       // not the first window of its read and strand
    ReadRec const R = C.reads[pair >> 1];
    uint64_t const m = R.root_m, k = R.root_k, n_cap = R.reserved1;
    bool const share = C.share && C.want_cigar;
    uint64_t u_start = 0, u_end = 0;
    for (uint32_t j = i; j < n_feasible && (C.key_s[j] >> kPosBits) == pair; ++j) {
        uint64_t const ws = C.key_s[j] & kPosMask, we = ws + C.len[C.idx_s[j]];
        bool joins = false;
        if (j > i && share && ws <= u_end) {
            uint64_t const new_end = we > u_end ? we : u_end;
            joins = new_end - u_start <= n_cap &&
                    unit_cost(new_end - u_start, m, k) <= unit_cost(u_end - u_start, m, k) + 0.75 * unit_cost(we - ws, m, k);
            if (joins) u_end = new_end;
        }
        if (!joins) { u_start = ws; u_end = we; }
        C.unit_start[j] = joins ? 0u : 1u;
    }
}

__global__ void root_units_kernel(RootCtx const C) {
    uint32_t const n_feasible = C.counters[kCtrFeasible];
    uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_feasible || !C.unit_start[i]) return;
    uint32_t const u = C.unit_of[i] - 1;
    uint64_t const pair = C.key_s[i] >> kPosBits;
    ReadRec const R = C.reads[pair >> 1];
    uint64_t const ws = C.key_s[i] & kPosMask;
    uint64_t u_end = ws + C.len[C.idx_s[i]];
    uint32_t count = 1;
    for (uint32_t j = i + 1; j < n_feasible && !C.unit_start[j] && (C.key_s[j] >> kPosBits) == pair; ++j) {
        uint64_t const we = (C.key_s[j] & kPosMask) + C.len[C.idx_s[j]];
        if (we > u_end) u_end = we;
        ++count;
    }
    UnitRec U;
    U.ws = ws; U.qbase = ((pair & 1) ? R.qoff_reverse : R.qoff_forward) + R.root_from; U.ck_base = 0;
    U.n = uint32_t(u_end - ws); U.m = R.root_m; U.k = R.root_k; U.first = i; U.count = count; U.cls = R.reserved0; U.pair = uint32_t(pair);
    uint32_t const W = C.classes[U.cls].W, rows = 32 * W, nb = (U.m + rows - 1) / rows;
    int64_t const band = int64_t(U.n) - int64_t(U.m) + 2 * int64_t(U.k) + 1;
    U.words = C.want_cigar ? uint32_t((uint64_t(nb) * ck_records_per_block(band, rows) * ck_record_words(W) + 3) & ~uint64_t(3)) : 0u;
    C.units[u] = U;
    C.unit_words[u] = U.words;
    atomicAdd(C.counters + kCtrUnits, 1u);
    atomicAdd(reinterpret_cast<unsigned long long*>(C.counters + kCtrCkWords), (unsigned long long)U.words);
    atomicAdd(C.counters + kCtrClassUnits + U.cls, 1u);
}

__global__ void root_tasks_kernel(RootCtx const C, uint32_t n_units) {
    uint32_t const u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_units) return;
    UnitRec U = C.units[u];
    U.ck_base = C.unit_ck[u];
    C.units[u].ck_base = U.ck_base;
    uint32_t const slot = class_offset(C.counters + kCtrClassUnits, U.cls) + atomicAdd(C.counters + kCtrClassFill + U.cls, 1u);
    DpTask t;
    t.ref_base = U.ws; t.query_base = U.qbase; t.trace_base = U.ck_base;
    t.n = U.n; t.m = U.m; t.dlo = -int32_t(U.k); t.dhi = int32_t(U.n) - int32_t(U.m) + int32_t(U.k);
    t.flags = C.want_cigar ? 0u : kFlagReverse; t.out = u;
    C.tasks[slot] = t;
    // word-steps the engine issues for it (host: word_steps_of)
    uint32_t const W = C.classes[U.cls].W, rows = 32 * W, nb = (U.m + rows - 1) / rows;
    int64_t const pad = int64_t(nb) * rows - U.m;
    unsigned long long ws = 0;
    for (uint32_t b = 0; b < nb; ++b) {
        int64_t lo = int64_t(rows) * b + 1 + t.dlo - pad, hi = int64_t(rows) * (b + 1) + t.dhi - pad;
        if (lo < 1) lo = 1;
        if (hi > int64_t(t.n)) hi = t.n;
        if (hi >= lo) ws += (unsigned long long)(hi - lo + 1) * W;
    }
    warp_add(reinterpret_cast<unsigned long long*>(C.counters + kCtrWordSteps), ws);
    warp_add_keyed(reinterpret_cast<unsigned long long*>(C.counters + kCtrClassWs), 1, U.cls, ws);      // per class: the launches are timed one by one
}

// minimum of the last row over columns col_from .. col_to of a checkpointed pass, and the rightmost column attaining it,
// from the records of the pass' last block (see range_min_kernel)
__device__ DpResult range_min_of(const uint32_t* __restrict__ ck_all, uint64_t ck_base, uint32_t n, uint32_t m, int32_t t_dlo, int32_t t_dhi, uint32_t W,
                                 uint32_t col_from, uint32_t col_to, DpResult const known) {
    uint32_t const ROWS = 32 * W, RECW = ck_record_words(W), BO = W == 1 ? 2 : 2 * W;
    uint32_t const nb = (m + ROWS - 1) / ROWS, lb = nb - 1;
    uint32_t const pad = nb * ROWS - m;
    int32_t const dlo = t_dlo - int32_t(pad), dhi = t_dhi - int32_t(pad);
    int32_t const lo = int32_t(ROWS) * int32_t(lb) + 1 + dlo, hi = int32_t(ROWS) * int32_t(nb) + dhi;
    int32_t const cs = lo < 1 ? 1 : lo, ce = hi > int32_t(n) ? int32_t(n) : hi;       // columns the last block worked on
    uint32_t const ck_per_block = ck_records_per_block(int64_t(t_dhi) - int64_t(t_dlo) + 1, ROWS);
    uint32_t const q_first = uint32_t(cs + int32_t(lb) + 31) >> 5;
    const uint32_t* const recs = ck_all + ck_base + uint64_t(lb) * ck_per_block * RECW + BO;
    auto rec_of = [&](uint32_t t) { return reinterpret_cast<const uint2*>(recs + uint64_t(((t + 31) >> 5) - q_first) * RECW); };
    auto prefix = [&](int32_t x) -> int32_t {                 // sum of the deltas of columns cs .. x (0 for x < cs)
        if (x < cs) return 0;
        uint32_t const t0 = uint32_t(cs) + lb, t1 = uint32_t(x) + lb;
        int32_t sum = 0;
        for (uint32_t q = (t0 + 31) >> 5; q <= (t1 + 31) >> 5; ++q) {
            uint2 const v = *rec_of(32 * (q - 1) + 1);
            uint32_t const first = 32 * (q - 1) + 1;                  // step of bit 0
            uint32_t mask = 0xffffffffu;
            if (t0 > first) mask &= 0xffffffffu << (t0 - first);
            if (t1 < first + 31) mask &= 0xffffffffu >> (first + 31 - t1);
            sum += __popc(v.x & mask) - __popc(v.y & mask);
        }
        return sum;
    };
    DpResult R; R.score = kNoScore; R.end_col = 0;
    int32_t const a = int32_t(col_from) > cs ? int32_t(col_from) : cs, b = int32_t(col_to) < ce ? int32_t(col_to) : ce;
    if (a <= b && known.score < kNoScore && int32_t(known.end_col) >= cs && int32_t(known.end_col) <= ce) {
        int32_t sc = known.score - prefix(int32_t(known.end_col)) + prefix(a - 1);        // row m at column a - 1
        uint2 v = make_uint2(0u, 0u);
        for (int32_t col = a; col <= b; ++col) {
            uint32_t const t = uint32_t(col) + lb, bit = (t - 1) & 31u;
            if (col == a || bit == 0) v = *rec_of(t);
            sc += int32_t((v.x >> bit) & 1u) - int32_t((v.y >> bit) & 1u);
            if (sc <= R.score) { R.score = sc; R.end_col = uint32_t(col); }
        }
    }
    return R;
}

// ---- per member: its result off the unit's pass; accepted? safe? ----
__global__ void root_results_kernel(RootCtx const C) {
    uint32_t const n_feasible = C.counters[kCtrFeasible];
    uint32_t const j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_feasible) return;
    uint32_t const u = C.unit_of[j] - 1;
    UnitRec const U = C.units[u];
    DpResult R = C.results[u];
    uint64_t const ws = C.key_s[j] & kPosMask;
    uint32_t const shift = uint32_t(ws - U.ws), len = C.len[C.idx_s[j]];
    uint8_t f = 0;
    if (U.count > 1) {
        f |= kMemberMulti;
        (void)warp_slot(C.counters, kCtrShared);
        R = range_min_of(C.ck, U.ck_base, U.n, U.m, -int32_t(U.k), int32_t(U.n) - int32_t(U.m) + int32_t(U.k), C.classes[U.cls].W, shift + 1, shift + len, R);
    }
    if (R.score == kPoisonScore) atomicAdd(C.counters + kCtrErrors, 1u);
    uint64_t key2 = ~uint64_t(0);
    if (R.score <= int32_t(U.k)) {
        // every alignment of this cost ending here begins at or after end - m - score: inside this member's window?
        bool const safe = int64_t(R.end_col) - int64_t(U.m) - int64_t(R.score) >= int64_t(shift);
        if ((f & kMemberMulti) && !safe) atomicAdd(C.counters + kCtrSlow, 1u);         // the shared pass cannot vouch for it
        else {
            f |= kMemberAccepted | (safe ? kMemberSafe : 0);
            (void)warp_slot(C.counters, kCtrAccepted);
            if (safe && C.want_cigar) key2 = (uint64_t(U.pair) << kPosBits) | (U.ws + R.end_col);
        }
    }
    C.m_score[j] = R.score; C.m_end[j] = R.end_col; C.m_flag[j] = f;
    C.key2[j] = key2; C.idx2[j] = j;
    C.rep[j] = j;
    C.tb_cap[j] = 0;
}

// ---- alignments of one read and strand that end at the same position with the same score share one traceback ----
// (sorted by (read, strand | end position); the first of equal ones -- in window order -- is traced back)
__global__ void root_dedup_kernel(RootCtx const C) {
    uint32_t const n_feasible = C.counters[kCtrFeasible];
    uint32_t const t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_feasible) return;
    uint64_t const key = C.key2_s[t];
    if (key == ~uint64_t(0)) return;                               // not accepted, or not safe: no sharing
    uint32_t const j = C.idx2_s[t];
    int32_t const score = C.m_score[j];
    uint32_t first = t;
    while (first > 0 && C.key2_s[first - 1] == key) --first;
    uint32_t r = j;
    for (uint32_t x = first; x < t; ++x) { uint32_t const jx = C.idx2_s[x]; if (C.m_score[jx] == score) { r = jx; break; } }
    C.rep[j] = r;
}

// cigar slots of the members that are traced back themselves (scan input), tracebacks per block width
__global__ void root_walk_count_kernel(RootCtx const C) {
    uint32_t const n_feasible = C.counters[kCtrFeasible];
    uint32_t const j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_feasible) return;
    if (!(C.m_flag[j] & kMemberAccepted) || C.rep[j] != j) return;
    uint32_t const cap = 2u * uint32_t(C.m_score[j]) + 3u;
    C.tb_cap[j] = cap;
    UnitRec const U = C.units[C.unit_of[j] - 1];
    uint32_t const W = C.classes[U.cls].W;
    uint32_t const widx = W == 1 ? 0 : W == 2 ? 1 : W == 4 ? 2 : W == 8 ? 3 : W == 16 ? 4 : 5;
    atomicAdd(C.counters + kCtrWidthTb + widx, 1u);
    atomicAdd(C.counters + kCtrTracebacks, 1u);
    atomicAdd(reinterpret_cast<unsigned long long*>(C.counters + kCtrCigars), (unsigned long long)cap);
}

__global__ void root_walks_kernel(RootCtx const C, uint64_t cigar_base) {
    uint32_t const n_feasible = C.counters[kCtrFeasible];
    uint32_t const j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_feasible || C.tb_cap[j] == 0) return;
    UnitRec const U = C.units[C.unit_of[j] - 1];
    uint32_t const W = C.classes[U.cls].W;
    uint32_t const widx = W == 1 ? 0 : W == 2 ? 1 : W == 4 ? 2 : W == 8 ? 3 : W == 16 ? 4 : 5;
    uint32_t const slot = class_offset(C.counters + kCtrWidthTb, widx) + atomicAdd(C.counters + kCtrWidthFill + widx, 1u);
    Walk2Task t;
    t.ck_base = U.ck_base; t.ref_base = U.ws; t.query_base = U.qbase;
    t.cigar_cap = uint32_t(C.tb_cap[j]); t.cigar_base = cigar_base + C.tb_cig[j];
    t.n = U.n; t.m = U.m; t.dlo = -int32_t(U.k); t.dhi = int32_t(U.n) - int32_t(U.m) + int32_t(U.k);
    t.end_col = C.m_end[j]; t.score = uint32_t(C.m_score[j]); t.flags = 0; t.out = slot; t.reserved = 0;
    C.wtasks[slot] = t;
    C.tb_slot[j] = slot;
}

// ---- alignment records in anchor order ----
__global__ void root_hits_kernel(RootCtx const C) {
    uint32_t const q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= C.n_roots) return;
    uint32_t h = 0;
    if (C.key[q] != ~uint64_t(0)) h = (C.m_flag[C.pos_of[q]] & kMemberAccepted) ? 1u : 0u;
    C.hit[q] = h;
}

__global__ void root_finish_kernel(RootCtx const C) {
    uint32_t const q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= C.n_roots || !C.hit[q]) return;
    uint32_t const j = C.pos_of[q];
    RootEntry const e = C.entries[q];
    ::fxg_alignment a;
    a.num_errors = uint32_t(C.m_score[j]); a.read_index = C.read0 + C.read[q]; a.reference_id = e.ref_id; a.orientation = C.orient[q];
    a.cigar_offset = 0; a.cigar_len = 0;
    for (int x = 0; x < 7; ++x) a.reserved[x] = 0;
    uint64_t const ws = C.ws[q];
    if (C.want_cigar) {
        uint32_t const r = C.rep[j];
        UnitRec const U = C.units[C.unit_of[r] - 1];
        uint32_t const slot = C.tb_slot[r];
        WalkResult const R = C.wresults[slot];
        Walk2Task const T = C.wtasks[slot];
        uint64_t const begin_abs = U.ws + R.begin_col;                               // store position of the alignment's first base
        if (R.cigar_len == 0xffffffffu || begin_abs < ws) { atomicAdd(C.counters + kCtrErrors, 1u); a.start_in_reference = 0; }
        else {
            a.start_in_reference = begin_abs - C.ref_base[e.ref_id];                 // alignment.cpp:175
            a.cigar_len = R.cigar_len; a.cigar_offset = T.cigar_base + T.cigar_cap - R.cigar_len;
        }
    } else {
        // reversed views: begin = window start + (window length - end column of the reversed pass), alignment.cpp:135-139
        a.start_in_reference = (ws - C.ref_base[e.ref_id]) + (C.len[q] - C.m_end[j]);
    }
    C.out[C.hit_at[q]] = a;
}

}  // namespace fxg
