/*
 * cpu_baseline.c -- multithreaded CPU port of the reference's verification path, in the
 * algorithmic shape of the code the reference actually runs (floxer 0.2.0 + SeqAn3 edit distance):
 *   - Myers/Hyyroe bit-vectors on 64-bit words, block-wise, with Ukkonen cut-off at k
 *     (seqan3 edit_distance_unbanded with align_cfg::min_score; call sites src/lib/alignment.cpp:101,122,160)
 *   - CIGAR mode keeps three trace bit-vectors (left / diagonal / up) per column for the whole active
 *     matrix and walks them back (seqan3 edit_distance_trace_matrix_full; alignment.cpp:156-178)
 *   - no-CIGAR mode runs on reversed views (alignment.cpp:115-145)
 *   - the per-anchor walk of src/lib/verification.cpp:8-245, reads distributed over threads
 *     (the reference distributes anchor packages over BS::thread_pool, parallelization.cpp:131-137)
 *
 * TEST / BENCH INFRASTRUCTURE ONLY (bench.py cpu_baseline and --impl reference, tests/): this is the
 * "port" CPU baseline because the reference binary cannot be built offline.  It is never
 * linked into the product library.  Its results are checked against floxer_oracle.c in tests/.
 */
#define _GNU_SOURCE
#include "../include/floxer_gpu.h"

#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>

#define WS 64
#define NSYM 6

/* ------------------------------------------------------------------ small helpers */

static uint64_t ceil_eps(double v) { return (uint64_t)(ceil(v - 1e-9) + 1e-9); }            /* math.hpp:22-27 */
static uint64_t nlen(const fxg_pex_node* n) { return n->query_index_to - n->query_index_from + 1; }
static int is_root(const fxg_pex_node* n) { return n->parent_id == FXG_NULL_ID; }

typedef struct { uint64_t offset, length, extra; } span_t;

/* verification.cpp:157-184 */
static span_t compute_span(uint64_t anchor_pos, const fxg_pex_node* node, uint64_t leaf_from, uint64_t ref_len, double ratio) {
    uint64_t const base = nlen(node) + 2 * node->num_errors + 1;
    uint64_t const extra = ceil_eps((double)base * ratio);
    int64_t const s = (int64_t)anchor_pos - (int64_t)(leaf_from - node->query_index_from) - (int64_t)node->num_errors - (int64_t)extra;
    span_t r;
    r.offset = s >= 0 ? (uint64_t)s : 0;
    uint64_t const want = base + 2 * extra, avail = ref_len - r.offset;
    r.length = want < avail ? want : avail;
    r.extra = extra;
    return r;
}

/* ------------------------------------------------------------------ per-thread scratch */

typedef struct {
    uint64_t* peq; size_t peq_cap;          /* [NSYM][nw] */
    uint64_t* pv; size_t pv_cap; uint64_t* mv; size_t mv_cap; int32_t* score; size_t score_cap;
    uint64_t* trace; size_t trace_cap;      /* 3 words per (column, active block) */
    size_t* col_off; size_t col_cap; uint32_t* col_last; size_t col_last_cap;
    uint8_t* ops; size_t ops_cap;
} scratch_t;

static int grow(void** p, size_t* cap, size_t need, size_t elt) {
    if (need <= *cap) return 0;
    size_t nc = *cap ? *cap : 64;
    while (nc < need) nc *= 2;
    void* np = realloc(*p, nc * elt);
    if (!np) return -1;
    *p = np; *cap = nc;
    return 0;
}

static void scratch_free(scratch_t* s) {
    free(s->peq); free(s->pv); free(s->mv); free(s->score); free(s->trace); free(s->col_off); free(s->col_last); free(s->ops);
}

/* ------------------------------------------------------------------ the aligner */

/* Myers block step with horizontal input hin in {-1,0,+1}; returns hout.  Optionally reports the
 * pre-shift horizontal-plus vector and the diagonal-zero vector for the trace matrix. */
static inline int block_step(uint64_t* Pv, uint64_t* Mv, uint64_t Eq, int hin, uint64_t top,
                             uint64_t* ph_out, uint64_t* d0_out) {
    uint64_t const pv = *Pv, mv = *Mv;
    uint64_t const hin_neg = (uint64_t)(hin < 0);
    uint64_t const xv = Eq | mv;
    uint64_t const eq = Eq | hin_neg;
    uint64_t const xh = (((eq & pv) + pv) ^ pv) | eq;
    uint64_t ph = mv | ~(xh | pv);
    uint64_t mh = pv & xh;
    int hout = 0;
    if (ph & top) hout = 1; else if (mh & top) hout = -1;
    if (ph_out) { *ph_out = ph; *d0_out = xh | mv; }
    ph = (ph << 1) | (uint64_t)(hin > 0);
    mh = (mh << 1) | hin_neg;
    *Pv = mh | ~(xv | ph);
    *Mv = ph & xv;
    return hout;
}

/*
 * Semi-global alignment of q[0..m) inside r[0..n).  rev walks both sequences backwards.
 * store_trace keeps (hp, d, vp) per active block per column.  On return: 1 if min over last row <= k
 * with *best / *best_col (1-based exclusive end column, rightmost minimum), else 0; -1 on OOM.
 */
static int myers_semiglobal(scratch_t* S, const uint8_t* r, size_t n, const uint8_t* q, size_t m, uint32_t k,
                            int rev, int store_trace, uint32_t* best_out, size_t* best_col_out) {
    if (m == 0) { *best_out = 0; *best_col_out = n; return 1; }
    size_t const nw = (m + WS - 1) / WS;
    if (grow((void**)&S->peq, &S->peq_cap, NSYM * nw, sizeof(uint64_t))) return -1;
    if (grow((void**)&S->pv, &S->pv_cap, nw, sizeof(uint64_t))) return -1;
    if (grow((void**)&S->mv, &S->mv_cap, nw, sizeof(uint64_t))) return -1;
    if (grow((void**)&S->score, &S->score_cap, nw, sizeof(int32_t))) return -1;
    memset(S->peq, 0, NSYM * nw * sizeof(uint64_t));
    for (size_t i = 0; i < m; ++i) {
        uint8_t const c = rev ? q[m - 1 - i] : q[i];
        S->peq[(size_t)c * nw + i / WS] |= 1ull << (i % WS);
    }
    if (store_trace) {
        if (grow((void**)&S->col_off, &S->col_cap, n + 2, sizeof(size_t))) return -1;
        if (grow((void**)&S->col_last, &S->col_last_cap, n + 2, sizeof(uint32_t))) return -1;
    }
    uint64_t const last_top = 1ull << ((m - 1) % WS);
    size_t last = 0;                           /* last active block */
    /* Ukkonen: initially rows 0..k are <= k */
    {
        size_t init_last = (k + 1 + WS - 1) / WS;
        if (init_last == 0) init_last = 1;
        if (init_last > nw) init_last = nw;
        last = init_last - 1;
    }
    for (size_t b = 0; b <= last; ++b) {
        S->pv[b] = ~0ull; S->mv[b] = 0;
        S->score[b] = (int32_t)((b == nw - 1) ? m : (b + 1) * WS);
    }
    uint32_t best = k + 1; size_t best_col = 0; int found = 0;
    if (last == nw - 1 && (uint32_t)S->score[last] <= k) { best = (uint32_t)S->score[last]; best_col = 0; found = 1; }
    size_t trace_used = 0;

    for (size_t j = 1; j <= n; ++j) {
        uint8_t const c = rev ? r[n - j] : r[j - 1];
        const uint64_t* peq = S->peq + (size_t)c * nw;
        int hin = 0;                           /* semi-global: first row is all zeros */
        int32_t prev_last_score = S->score[last];
        if (store_trace) {
            /* at most last + 2 blocks are written for this column */
            if (grow((void**)&S->trace, &S->trace_cap, trace_used + 3 * (last + 2), sizeof(uint64_t))) return -1;
            S->col_off[j] = trace_used;
        }
        for (size_t b = 0; b <= last; ++b) {
            uint64_t const top = (b == nw - 1) ? last_top : (1ull << 63);
            uint64_t ph = 0, d0 = 0;
            int const hout = block_step(&S->pv[b], &S->mv[b], peq[b], hin, top, store_trace ? &ph : NULL, &d0);
            if (store_trace) {
                uint64_t* t = S->trace + trace_used; trace_used += 3;
                t[0] = ph; t[1] = ~(peq[b] ^ d0); t[2] = S->pv[b];
            }
            S->score[b] += hout;
            hin = hout;
        }
        /* extend the active region if the block below could hold a cell <= k (see SeqAn3 / Myers cut-off) */
        if (last + 1 < nw && (prev_last_score <= (int32_t)k || S->score[last] <= (int32_t)k)) {
            size_t const b = ++last;
            uint64_t const top = (b == nw - 1) ? last_top : (1ull << 63);
            S->pv[b] = ~0ull; S->mv[b] = 0;
            int32_t const rows = (int32_t)((b == nw - 1) ? (m - b * WS) : WS);
            uint64_t ph = 0, d0 = 0;
            int const hout = block_step(&S->pv[b], &S->mv[b], peq[b], hin, top, store_trace ? &ph : NULL, &d0);
            if (store_trace) {
                uint64_t* t = S->trace + trace_used; trace_used += 3;
                t[0] = ph; t[1] = ~(peq[b] ^ d0); t[2] = S->pv[b];
            }
            /* previous column of the new block is (bottom of block above, previous column) + 1..rows */
            S->score[b] = (S->score[b - 1] - hin) + rows + hout;
        }
        if (store_trace) S->col_last[j] = (uint32_t)last;
        if (last == nw - 1 && S->score[last] >= 0 && (uint32_t)S->score[last] <= best && (uint32_t)S->score[last] <= k) {
            best = (uint32_t)S->score[last]; best_col = j; found = 1;
        }
        /* shrink: a block whose bottom value is >= k + (rows in block) holds no cell <= k */
        while (last > 0) {
            int32_t const rows = (int32_t)((last == nw - 1) ? (m - last * WS) : WS);
            if (S->score[last] >= (int32_t)k + rows) --last; else break;
        }
    }
    if (!found) return 0;
    *best_out = best; *best_col_out = best_col;
    return 1;
}

/* alignment::align (alignment.cpp:83-181).  cigar is written to *cig (grown as needed). */
typedef struct { uint32_t* v; size_t n, cap; } cigar_buf;

static int cpu_align(scratch_t* S, const uint8_t* r, size_t n, const uint8_t* q, size_t m, uint32_t k, int mode,
                     uint32_t* num_errors, uint64_t* start, cigar_buf* cig, uint32_t* cig_len) {
    uint32_t best; size_t col;
    *cig_len = 0;
    if (mode == FXG_MODE_EXISTS) return myers_semiglobal(S, r, n, q, m, k, 0, 0, &best, &col);
    if (mode == FXG_MODE_NO_CIGAR) {
        int const rc = myers_semiglobal(S, r, n, q, m, k, 1, 0, &best, &col);
        if (rc == 1) { *num_errors = best; *start = n - col; }
        return rc;
    }
    int const rc = myers_semiglobal(S, r, n, q, m, k, 0, 1, &best, &col);
    if (rc != 1) return rc;
    if (grow((void**)&S->ops, &S->ops_cap, m + n + 1, 1)) return -1;
    size_t n_ops = 0, i = m, j = col;
    while (i > 0) {
        if (j == 0) { S->ops[n_ops++] = FXG_CIGAR_I; --i; continue; }       /* column 0: only "up" */
        size_t const b = (i - 1) / WS; uint64_t const bit = 1ull << ((i - 1) % WS);
        if (b > S->col_last[j]) return -1;                                  /* outside the active matrix */
        const uint64_t* t = S->trace + S->col_off[j] + 3 * b;
        /* trace priority left > up > diagonal (see oracle/floxer_oracle.h, FXO_TRACE_PRIORITY) */
        if (t[0] & bit) { S->ops[n_ops++] = FXG_CIGAR_D; --j; }
        else if (t[2] & bit) { S->ops[n_ops++] = FXG_CIGAR_I; --i; }
        else if (t[1] & bit) { S->ops[n_ops++] = (q[i - 1] == r[j - 1]) ? FXG_CIGAR_EQ : FXG_CIGAR_X; --i; --j; }
        else return -1;
    }
    *num_errors = best; *start = j;
    uint32_t len = 0;
    for (size_t p = n_ops; p > 0;) {
        uint8_t const op = S->ops[p - 1]; uint32_t run = 0;
        while (p > 0 && S->ops[p - 1] == op) { ++run; --p; }
        if (grow((void**)&cig->v, &cig->cap, cig->n + 1, sizeof(uint32_t))) return -1;
        cig->v[cig->n++] = (run << 4) | op; ++len;
    }
    *cig_len = len;
    return 1;
}

/* ------------------------------------------------------------------ verified intervals (intervals.cpp:84-127) */

typedef struct { uint64_t* se; size_t n, cap; } ivset;   /* pairs start,end */

static int iv_contains(const ivset* s, uint64_t a, uint64_t b) {
    for (size_t i = 0; i < s->n; ++i) if (s->se[2 * i] <= a && s->se[2 * i + 1] >= b) return 1;
    return 0;
}
static int iv_insert(ivset* s, uint64_t a, uint64_t b) {
    if (iv_contains(s, a, b)) return 0;
    if (grow((void**)&s->se, &s->cap, 2 * (s->n + 1), sizeof(uint64_t))) return -1;
    s->se[2 * s->n] = a; s->se[2 * s->n + 1] = b; s->n++;
    return 0;
}

/* ------------------------------------------------------------------ per-read verification */

typedef struct {
    fxg_alignment* alns; size_t n_alns, cap_alns;
    cigar_buf cig;
    fxg_stats stats;
} read_out;

typedef struct {
    size_t n_refs; const uint8_t* const* refs; const uint64_t* ref_lens;
    const fxg_verify_config* cfg;
    const fxg_read* reads; size_t n_reads;
    const uint8_t* fwd; const uint8_t* rc;
    const fxg_pex_node* nodes; const fxg_anchor* anchors;
    read_out* outs;
    atomic_size_t next;
    atomic_int failed;
} job_t;

static int verify_read(job_t* J, scratch_t* S, size_t ri) {
    const fxg_read* R = &J->reads[ri];
    read_out* O = &J->outs[ri];
    const fxg_pex_node* inner = J->nodes + R->node_offset;
    const fxg_pex_node* leaves = inner + R->num_inner;
    const fxg_pex_node* root = R->num_inner ? &inner[0] : &leaves[0];
    double const ratio = J->cfg->extra_verification_ratio;
    ivset* sets = (ivset*)calloc(J->n_refs ? J->n_refs : 1, sizeof(ivset));
    if (!sets) return -1;
    int rc_all = 0;
    for (int orient = 0; orient < 2 && !rc_all; ++orient) {
        const uint8_t* query = (orient == 0 ? J->fwd : J->rc) + R->query_offset;
        const fxg_anchor* A = J->anchors + R->anchor_offset + (orient == 0 ? 0 : R->num_anchors_forward);
        size_t const nA = orient == 0 ? R->num_anchors_forward : R->num_anchors_reverse;
        for (size_t r = 0; r < J->n_refs; ++r) sets[r].n = 0;
        for (size_t ai = 0; ai < nA && !rc_all; ++ai) {
            const fxg_anchor* a = &A[ai];
            if (a->pex_leaf_index >= R->num_leaves || a->reference_id >= J->n_refs) { rc_all = -1; break; }
            const fxg_pex_node* leaf = &leaves[a->pex_leaf_index];
            uint64_t const ref_len = J->ref_lens[a->reference_id];
            ivset* ivs = &sets[a->reference_id];
            span_t const rs = compute_span(a->reference_position, root, leaf->query_index_from, ref_len, ratio);
            if (J->cfg->interval_optimization) {                      /* root_was_already_verified, verification.cpp:119-136 */
                uint64_t const s0 = rs.offset, e0 = rs.offset + rs.length;
                uint64_t const e1t = rs.extra > e0 ? 0 : e0 - rs.extra;
                uint64_t const ne = (s0 + 1 > e1t) ? s0 + 1 : e1t;
                uint64_t const ns = (ne - 1 < s0 + rs.extra) ? ne - 1 : s0 + rs.extra;
                if (iv_contains(ivs, ns, ne)) { O->stats.n_avoided_root++; O->stats.sum_avoided_root += rs.length; continue; }
            }
            const fxg_pex_node* curr;
            if (J->cfg->verification_kind == FXG_KIND_DIRECT_FULL || is_root(leaf)) curr = root;
            else if (J->cfg->verification_kind == FXG_KIND_HIERARCHICAL) curr = &inner[leaf->parent_id];
            else { rc_all = -1; break; }
            for (;;) {
                int const root_now = is_root(curr);
                span_t const sp = root_now ? rs : compute_span(a->reference_position, curr, leaf->query_index_from, ref_len, 0.0);
                int const mode = !root_now ? FXG_MODE_EXISTS : (J->cfg->without_cigar ? FXG_MODE_NO_CIGAR : FXG_MODE_CIGAR);
                uint32_t errs = 0, clen = 0; uint64_t start = 0;
                size_t const cig_before = O->cig.n;
                size_t const m = (size_t)nlen(curr);
                int const res = cpu_align(S, J->refs[a->reference_id] + sp.offset, (size_t)sp.length,
                                          query + curr->query_index_from, m, (uint32_t)curr->num_errors, mode,
                                          &errs, &start, &O->cig, &clen);
                if (res < 0) { rc_all = -1; break; }
                if (root_now) { O->stats.n_aligned_root++; O->stats.sum_aligned_root += sp.length; O->stats.cells_root += (uint64_t)m * sp.length; }
                else { O->stats.n_aligned_inner++; O->stats.sum_aligned_inner += sp.length; O->stats.cells_inner += (uint64_t)m * sp.length; }
                if (res == 1 && root_now) {
                    if (grow((void**)&O->alns, &O->cap_alns, O->n_alns + 1, sizeof(fxg_alignment))) { rc_all = -1; break; }
                    fxg_alignment* al = &O->alns[O->n_alns++];
                    memset(al, 0, sizeof *al);
                    al->start_in_reference = sp.offset + start;
                    al->cigar_offset = cig_before; al->cigar_len = clen;
                    al->num_errors = errs; al->read_index = (uint32_t)ri;
                    al->reference_id = (uint32_t)a->reference_id; al->orientation = (uint8_t)orient;
                }
                if (root_now && J->cfg->interval_optimization) {
                    if (iv_insert(ivs, sp.offset, sp.offset + sp.length)) { rc_all = -1; break; }
                }
                if (res == 0 || root_now) break;
                curr = &inner[curr->parent_id];
            }
        }
    }
    for (size_t r = 0; r < J->n_refs; ++r) free(sets[r].se);
    free(sets);
    return rc_all;
}

static void* worker(void* arg) {
    job_t* J = (job_t*)arg;
    scratch_t S; memset(&S, 0, sizeof S);
    for (;;) {
        size_t const ri = atomic_fetch_add(&J->next, 1);
        if (ri >= J->n_reads || atomic_load(&J->failed)) break;
        if (verify_read(J, &S, ri) != 0) atomic_store(&J->failed, 1);
    }
    scratch_free(&S);
    return NULL;
}

/* ------------------------------------------------------------------ exported API */

typedef struct fxc_result {
    fxg_alignment* alns; size_t n_alns;
    uint32_t* cigars; size_t n_cigars;
    fxg_stats stats;
} fxc_result;

void fxc_result_free(fxc_result* r) { if (r) { free(r->alns); free(r->cigars); free(r); } }
size_t fxc_result_num_alignments(const fxc_result* r) { return r->n_alns; }
const fxg_alignment* fxc_result_alignments(const fxc_result* r) { return r->alns; }
size_t fxc_result_cigar_len(const fxc_result* r) { return r->n_cigars; }
const uint32_t* fxc_result_cigar_pool(const fxc_result* r) { return r->cigars; }
const fxg_stats* fxc_result_stats(const fxc_result* r) { return &r->stats; }

int fxc_verify_reads(size_t n_refs, const uint8_t* const* refs, const uint64_t* ref_lens,
                     const fxg_verify_config* cfg, const fxg_read* reads, size_t n_reads,
                     const uint8_t* fwd, const uint8_t* rc, const fxg_pex_node* nodes, const fxg_anchor* anchors,
                     int threads, fxc_result** out) {
    job_t J; memset(&J, 0, sizeof J);
    J.n_refs = n_refs; J.refs = refs; J.ref_lens = ref_lens; J.cfg = cfg; J.reads = reads; J.n_reads = n_reads;
    J.fwd = fwd; J.rc = rc; J.nodes = nodes; J.anchors = anchors;
    J.outs = (read_out*)calloc(n_reads ? n_reads : 1, sizeof(read_out));
    if (!J.outs) return FXG_ERR_OUT_OF_MEMORY;
    atomic_init(&J.next, 0); atomic_init(&J.failed, 0);
    if (threads < 1) threads = 1;
    pthread_t* th = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
    int started = 0;
    for (int t = 0; t < threads - 1; ++t) { if (pthread_create(&th[started], NULL, worker, &J) == 0) ++started; }
    worker(&J);
    for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
    free(th);
    int rc_all = atomic_load(&J.failed) ? FXG_ERR_INVALID_ARGUMENT : FXG_OK;
    fxc_result* R = (fxc_result*)calloc(1, sizeof *R);
    if (!R) rc_all = FXG_ERR_OUT_OF_MEMORY;
    if (rc_all == FXG_OK) {
        size_t na = 0, nc = 0;
        for (size_t i = 0; i < n_reads; ++i) { na += J.outs[i].n_alns; nc += J.outs[i].cig.n; }
        R->alns = (fxg_alignment*)malloc((na ? na : 1) * sizeof(fxg_alignment));
        R->cigars = (uint32_t*)malloc((nc ? nc : 1) * sizeof(uint32_t));
        if (!R->alns || !R->cigars) rc_all = FXG_ERR_OUT_OF_MEMORY;
        else {
            for (size_t i = 0; i < n_reads; ++i) {
                read_out* O = &J.outs[i];
                for (size_t a = 0; a < O->n_alns; ++a) {
                    fxg_alignment al = O->alns[a];
                    al.cigar_offset += R->n_cigars;
                    R->alns[R->n_alns++] = al;
                }
                if (O->cig.n) memcpy(R->cigars + R->n_cigars, O->cig.v, O->cig.n * sizeof(uint32_t));
                R->n_cigars += O->cig.n;
                uint64_t* d = (uint64_t*)&R->stats; const uint64_t* s = (const uint64_t*)&O->stats;
                for (size_t f = 0; f < sizeof(fxg_stats) / sizeof(uint64_t); ++f) d[f] += s[f];
            }
        }
    }
    for (size_t i = 0; i < n_reads; ++i) { free(J.outs[i].alns); free(J.outs[i].cig.v); }
    free(J.outs);
    if (rc_all != FXG_OK) { fxc_result_free(R); return rc_all; }
    *out = R;
    return FXG_OK;
}

/* batched alignment::align on host spans, same task format as fxg_align_batch */
typedef struct {
    size_t n_refs; const uint8_t* const* refs; const uint64_t* ref_lens;
    const fxg_align_task* tasks; size_t n_tasks;
    const uint8_t* qpool; const uint8_t* ipool;
    fxg_align_result* results; cigar_buf* cigs;   /* one cigar_buf per thread-chunk */
    atomic_size_t next; atomic_int failed;
    uint32_t** task_cig; /* per task cigar copy */
} ajob_t;

static void* aworker(void* arg) {
    ajob_t* J = (ajob_t*)arg;
    scratch_t S; memset(&S, 0, sizeof S);
    cigar_buf cb; memset(&cb, 0, sizeof cb);
    for (;;) {
        size_t const ti = atomic_fetch_add(&J->next, 1);
        if (ti >= J->n_tasks || atomic_load(&J->failed)) break;
        const fxg_align_task* T = &J->tasks[ti];
        const uint8_t* r = (T->ref_id == FXG_REF_INLINE ? J->ipool : J->refs[T->ref_id]) + T->ref_offset;
        uint32_t errs = 0, clen = 0; uint64_t start = 0;
        cb.n = 0;
        int const res = cpu_align(&S, r, T->ref_len, J->qpool + T->query_offset, T->query_len, T->max_errors, T->mode,
                                  &errs, &start, &cb, &clen);
        if (res < 0) { atomic_store(&J->failed, 1); break; }
        fxg_align_result* R = &J->results[ti];
        memset(R, 0, sizeof *R);
        R->exists = (uint8_t)res; R->orientation = T->orientation;
        if (res == 1 && T->mode != FXG_MODE_EXISTS) {
            R->num_errors = errs; R->start_in_reference = T->reference_span_offset + start; R->cigar_len = clen;
            if (clen) {
                J->task_cig[ti] = (uint32_t*)malloc(clen * sizeof(uint32_t));
                if (!J->task_cig[ti]) { atomic_store(&J->failed, 1); break; }
                memcpy(J->task_cig[ti], cb.v, clen * sizeof(uint32_t));
            }
        }
    }
    free(cb.v);
    scratch_free(&S);
    return NULL;
}

int fxc_align_batch(size_t n_refs, const uint8_t* const* refs, const uint64_t* ref_lens,
                    const fxg_align_task* tasks, size_t n_tasks, const uint8_t* query_pool,
                    const uint8_t* inline_ref_pool, int threads,
                    fxg_align_result* results, uint32_t* cigar_pool, size_t cigar_capacity, size_t* cigar_used) {
    ajob_t J; memset(&J, 0, sizeof J);
    J.n_refs = n_refs; J.refs = refs; J.ref_lens = ref_lens; J.tasks = tasks; J.n_tasks = n_tasks;
    J.qpool = query_pool; J.ipool = inline_ref_pool; J.results = results;
    J.task_cig = (uint32_t**)calloc(n_tasks ? n_tasks : 1, sizeof(uint32_t*));
    if (!J.task_cig) return FXG_ERR_OUT_OF_MEMORY;
    atomic_init(&J.next, 0); atomic_init(&J.failed, 0);
    if (threads < 1) threads = 1;
    pthread_t* th = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
    int started = 0;
    for (int t = 0; t < threads - 1; ++t) { if (pthread_create(&th[started], NULL, aworker, &J) == 0) ++started; }
    aworker(&J);
    for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
    free(th);
    int rc = atomic_load(&J.failed) ? FXG_ERR_INVALID_ARGUMENT : FXG_OK;
    size_t used = 0;
    for (size_t i = 0; i < n_tasks && rc == FXG_OK; ++i) {
        if (results[i].cigar_len) {
            if (used + results[i].cigar_len > cigar_capacity) { rc = FXG_ERR_OVERFLOW; }
            else { memcpy(cigar_pool + used, J.task_cig[i], results[i].cigar_len * sizeof(uint32_t)); results[i].cigar_offset = used; }
            used += results[i].cigar_len;
        }
    }
    for (size_t i = 0; i < n_tasks; ++i) free(J.task_cig[i]);
    free(J.task_cig);
    if (cigar_used) *cigar_used = used;
    return rc;
}
