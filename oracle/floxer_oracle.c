/*
 * floxer_oracle.c -- see floxer_oracle.h.  TEST INFRASTRUCTURE ONLY (never linked into the product).
 * Plain C99, no dependencies.  Every function cites the reference lines it restates.
 */
#include "floxer_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ math.hpp */

/* include/math.hpp:18-20 */
uint64_t fxo_ceil_div(uint64_t a, uint64_t b) { return (a % b) ? a / b + 1 : a / b; }

/* include/math.hpp:22-27: ceil(value - eps) + eps, truncated to size_t */
uint64_t fxo_ceil_eps(double value) {
    static const double epsilon = 0.000000001;
    return (uint64_t)(ceil(value - epsilon) + epsilon);
}

/* ------------------------------------------------------------------ verification.cpp:157-184 */

static uint64_t node_len(const fxo_node* n) { return n->query_index_to - n->query_index_from + 1; } /* pex.cpp:52-54 */
static int node_is_root(const fxo_node* n) { return n->parent_id == FXO_NULL_ID; }                  /* pex.cpp:56-58 */

fxo_span fxo_compute_span(uint64_t anchor_reference_position, const fxo_node* node,
                          uint64_t leaf_query_index_from, uint64_t full_reference_length,
                          double extra_verification_ratio) {
    uint64_t const base_length = node_len(node) + 2 * node->num_errors + 1;
    uint64_t const extra = fxo_ceil_eps((double)base_length * extra_verification_ratio);

    int64_t const start_signed = (int64_t)anchor_reference_position
        - (int64_t)(leaf_query_index_from - node->query_index_from)
        - (int64_t)node->num_errors
        - (int64_t)extra;

    uint64_t const start = start_signed >= 0 ? (uint64_t)start_signed : 0;
    uint64_t const want = base_length + 2 * extra;
    uint64_t const avail = full_reference_length - start;
    fxo_span s;
    s.offset = start;
    s.length = want < avail ? want : avail;
    s.extra = extra;
    return s;
}

/* ------------------------------------------------------------------ intervals.cpp */

/* intervals.cpp:26-46 */
int fxo_interval_relationship(fxo_interval a, fxo_interval b) {
    if (a.start > b.end) return FXO_REL_COMPLETELY_ABOVE;
    if (a.end < b.start) return FXO_REL_COMPLETELY_BELOW;
    if (a.start == b.start && a.end == b.end) return FXO_REL_EQUAL;
    if (a.start <= b.start && a.end >= b.end) return FXO_REL_CONTAINS;
    if (a.start >= b.start && a.end <= b.end) return FXO_REL_INSIDE;
    if (a.start > b.start && a.start <= b.end) return FXO_REL_OVERLAP_ABOVE;
    return FXO_REL_OVERLAP_BELOW;
}

/* intervals.cpp:48-58 */
fxo_interval fxo_interval_trim(fxo_interval a, uint64_t amount) {
    uint64_t const e = amount > a.end ? 0 : a.end - amount;
    uint64_t const new_end = (a.start + 1 > e) ? a.start + 1 : e;
    uint64_t const new_start = (new_end - 1 < a.start + amount) ? new_end - 1 : a.start + amount;
    fxo_interval r = { new_start, new_end };
    return r;
}

/* intervals.cpp:84-127.  The reference keeps closed intervals in an interval tree and asks for all
 * intervals overlapping the target; any stored interval that contains the target necessarily
 * overlaps it, so a linear scan over all stored intervals decides the same predicate. */
struct fxo_intervals {
    int active;
    fxo_interval* v;
    size_t n, cap;
};

fxo_intervals* fxo_intervals_new(int active) {
    fxo_intervals* s = (fxo_intervals*)calloc(1, sizeof *s);
    if (s) s->active = active;
    return s;
}
void fxo_intervals_free(fxo_intervals* s) { if (s) { free(s->v); free(s); } }
void fxo_intervals_configure(fxo_intervals* s, int active) { s->active = active; }
size_t fxo_intervals_size(const fxo_intervals* s) { return s->n; }

int fxo_intervals_contains(const fxo_intervals* s, fxo_interval t) {
    if (!s->active) return 0;                                    /* intervals.cpp:97-99 */
    for (size_t i = 0; i < s->n; ++i) {
        int const rel = fxo_interval_relationship(s->v[i], t);   /* existing.relationship_with(target) */
        if (rel == FXO_REL_EQUAL || rel == FXO_REL_CONTAINS) return 1;
    }
    return 0;
}

void fxo_intervals_insert(fxo_intervals* s, fxo_interval iv) {
    if (!s->active || fxo_intervals_contains(s, iv)) return;     /* intervals.cpp:84-88 */
    if (s->n == s->cap) {
        size_t const nc = s->cap ? 2 * s->cap : 16;
        fxo_interval* nv = (fxo_interval*)realloc(s->v, nc * sizeof *nv);
        if (!nv) return;
        s->v = nv; s->cap = nc;
    }
    s->v[s->n++] = iv;
}

/* ------------------------------------------------------------------ pex.cpp: builders */

typedef struct {
    fxo_node* inner; size_t n_inner, cap_inner;
    fxo_node* leaves; size_t n_leaves, cap_leaves;
    uint64_t no_error_seed_length, leaf_max_num_errors;
    int oom;
} pex_builder;

static void push_node(fxo_node** v, size_t* n, size_t* cap, fxo_node x, int* oom) {
    if (*n == *cap) {
        size_t const nc = *cap ? 2 * *cap : 64;
        fxo_node* nv = (fxo_node*)realloc(*v, nc * sizeof *nv);
        if (!nv) { *oom = 1; return; }
        *v = nv; *cap = nc;
    }
    (*v)[(*n)++] = x;
}

/* pex.cpp:110-156, 1-based indices as in the reference */
static void add_nodes_recursive(pex_builder* b, uint64_t from, uint64_t to, uint64_t num_errors, uint64_t parent_id) {
    uint64_t const num_leafs_left = fxo_ceil_div(num_errors + 1, 2);
    fxo_node const curr = { parent_id, from - 1, to - 1, num_errors };
    if (b->oom) return;
    if (num_errors <= b->leaf_max_num_errors) {
        push_node(&b->leaves, &b->n_leaves, &b->cap_leaves, curr, &b->oom);
    } else {
        uint64_t const id = b->n_inner;
        push_node(&b->inner, &b->n_inner, &b->cap_inner, curr, &b->oom);
        uint64_t const split = from + num_leafs_left * b->no_error_seed_length;
        uint64_t const e_left = (num_leafs_left * num_errors) / (num_errors + 1);
        uint64_t const e_right = ((num_errors + 1 - num_leafs_left) * num_errors) / (num_errors + 1);
        add_nodes_recursive(b, from, split - 1, e_left, id);
        add_nodes_recursive(b, split, to, e_right, id);
    }
}

/* pex.cpp:242-256: sets child.parent_id, returns the parent */
static fxo_node create_parent_node(fxo_node* children, size_t n_children, uint64_t parent_id) {
    uint64_t errs = 0;
    for (size_t i = 0; i < n_children; ++i) { children[i].parent_id = parent_id; errs += children[i].num_errors; }
    fxo_node p = { 0, children[0].query_index_from, children[n_children - 1].query_index_to, errs + n_children - 1 };
    return p;
}

/* pex.cpp:158-240 */
static void add_nodes_bottom_up(pex_builder* b, uint64_t total_len, uint64_t query_num_errors, uint64_t leaf_max) {
    uint64_t const base_leaf_weight = leaf_max + 1;
    uint64_t const num_desired_leaves = fxo_ceil_div(query_num_errors + 1, base_leaf_weight);
    if (num_desired_leaves == 1) {
        fxo_node const root = { FXO_NULL_ID, 0, total_len - 1, query_num_errors };
        push_node(&b->leaves, &b->n_leaves, &b->cap_leaves, root, &b->oom);
        return;
    }
    /* create_leaves, pex.cpp:214-240 */
    uint64_t const base_seed_length = total_len / num_desired_leaves;
    uint64_t const remainder = total_len % num_desired_leaves;
    uint64_t start = 0;
    for (uint64_t i = 0; i < num_desired_leaves; ++i) {
        uint64_t const len = i < remainder ? base_seed_length + 1 : base_seed_length;
        fxo_node const leaf = { 0, start, start + len - 1, leaf_max };
        push_node(&b->leaves, &b->n_leaves, &b->cap_leaves, leaf, &b->oom);
        start += len;
    }
    if (b->oom) return;
    /* reserve so that pointers into inner stay valid while we append (pex.cpp:176-178) */
    b->inner = (fxo_node*)malloc((size_t)(num_desired_leaves + 1) * sizeof(fxo_node));
    if (!b->inner) { b->oom = 1; return; }
    b->cap_inner = (size_t)num_desired_leaves + 1;
    memset(&b->inner[0], 0, sizeof(fxo_node));
    b->n_inner = 1;                                     /* slot for the root, must be index 0 */

    fxo_node* level = b->leaves;
    size_t level_n = b->n_leaves;
    while (level_n > 3) {
        for (size_t i = 0; i < level_n; i += 2) {
            size_t const remaining = level_n - i;
            if (remaining == 1) break;
            size_t const n_children = (remaining == 3) ? 3 : 2;
            uint64_t const new_parent_id = b->n_inner;
            b->inner[b->n_inner] = create_parent_node(level + i, n_children, new_parent_id);
            b->n_inner++;
            /* NOTE: when remaining == 3 the loop continues with i += 2 -> remaining == 1 -> break */
        }
        size_t const next_n = level_n / 2;
        level = b->inner + (b->n_inner - next_n);       /* std::span(inner_nodes).last(size / 2) */
        level_n = next_n;
    }
    b->inner[0] = create_parent_node(level, level_n, 0);
    b->inner[0].parent_id = FXO_NULL_ID;
}

int fxo_pex_build(uint64_t total_query_length, uint64_t query_num_errors, uint64_t leaf_max_num_errors,
                  int build_strategy, fxo_node** inner, size_t* n_inner, fxo_node** leaves, size_t* n_leaves) {
    pex_builder b;
    memset(&b, 0, sizeof b);
    b.no_error_seed_length = total_query_length / (query_num_errors + 1);   /* pex.cpp:85 */
    b.leaf_max_num_errors = leaf_max_num_errors;
    if (build_strategy == FXO_BUILD_RECURSIVE) {
        add_nodes_recursive(&b, 1, total_query_length, query_num_errors, FXO_NULL_ID);
    } else if (build_strategy == FXO_BUILD_BOTTOM_UP) {
        add_nodes_bottom_up(&b, total_query_length, query_num_errors, leaf_max_num_errors);
    } else {
        return -1;
    }
    if (b.oom) { free(b.inner); free(b.leaves); return -1; }
    *inner = b.inner; *n_inner = b.n_inner; *leaves = b.leaves; *n_leaves = b.n_leaves;
    return 0;
}

void fxo_free(void* p) { free(p); }

/* ------------------------------------------------------------------ alignment.cpp:83-181 */

enum { TB_L = 1, TB_U = 2, TB_D = 4 };

/* last DP row of semi-global edit distance (reference ends free, query consumed entirely;
 * alignment.cpp:89-94).  rev != 0 walks both sequences backwards (alignment.cpp:118-125). */
static int last_row(const uint8_t* r, size_t n, const uint8_t* q, size_t m, int rev, uint32_t* row /* n+1 */) {
    for (size_t j = 0; j <= n; ++j) row[j] = 0;
    for (size_t i = 1; i <= m; ++i) {
        uint8_t const qc = rev ? q[m - i] : q[i - 1];
        uint32_t diag = row[0];
        row[0] = (uint32_t)i;
        for (size_t j = 1; j <= n; ++j) {
            uint8_t const rc = rev ? r[n - j] : r[j - 1];
            uint32_t const up = row[j];
            uint32_t best = diag + (qc != rc);
            if (up + 1 < best) best = up + 1;
            if (row[j - 1] + 1 < best) best = row[j - 1] + 1;
            diag = up;
            row[j] = best;
        }
    }
    return 0;
}

static void pick_end(const uint32_t* row, size_t n, int rightmost, uint32_t* best, size_t* best_col) {
    *best = row[0]; *best_col = 0;
    for (size_t j = 1; j <= n; ++j) {
        if (rightmost ? row[j] <= *best : row[j] < *best) { *best = row[j]; *best_col = j; }
    }
}

int fxo_align_ex(const uint8_t* reference, size_t n, const uint8_t* query, size_t m, size_t max_errors,
                 int mode, const char* priority, int rightmost,
                 uint64_t* num_errors, uint64_t* start_in_window,
                 uint32_t* cigar, size_t cigar_cap, size_t* cigar_len) {
    if (cigar_len) *cigar_len = 0;
    if (mode == FXO_MODE_EXISTS || mode == FXO_MODE_NO_CIGAR) {
        uint32_t* row = (uint32_t*)malloc((n + 1) * sizeof *row);
        if (!row) return -1;
        int const rev = (mode == FXO_MODE_NO_CIGAR);
        last_row(reference, n, query, m, rev, row);
        uint32_t best; size_t col;
        pick_end(row, n, rightmost, &best, &col);
        free(row);
        if (best > max_errors) return 0;                       /* min_score{-k}: score == infinite */
        if (mode == FXO_MODE_NO_CIGAR) {
            if (num_errors) *num_errors = best;
            /* reference_begin_position = reference.size() - sequence1_end_position (alignment.cpp:135) */
            if (start_in_window) *start_in_window = n - col;
        }
        return 1;
    }
    if (mode != FXO_MODE_CIGAR) return -1;

    /* full matrix with the set of valid predecessor moves per cell; this is what SeqAn3's
     * edit_distance_trace_matrix_full stores as three bit-vectors per column */
    size_t const W = n + 1;
    uint8_t* tb = (uint8_t*)malloc((m + 1) * W);
    uint32_t* row = (uint32_t*)malloc(W * sizeof *row);
    if (!tb || !row) { free(tb); free(row); return -1; }
    for (size_t j = 0; j <= n; ++j) { row[j] = 0; tb[j] = 0; }
    for (size_t i = 1; i <= m; ++i) {
        uint8_t const qc = query[i - 1];
        uint32_t diag = row[0];
        row[0] = (uint32_t)i;
        tb[i * W] = TB_U;
        for (size_t j = 1; j <= n; ++j) {
            uint32_t const up = row[j];
            uint32_t const dcost = diag + (qc != reference[j - 1]);
            uint32_t best = dcost;
            if (up + 1 < best) best = up + 1;
            if (row[j - 1] + 1 < best) best = row[j - 1] + 1;
            uint8_t f = 0;
            if (row[j - 1] + 1 == best) f |= TB_L;
            if (up + 1 == best) f |= TB_U;
            if (dcost == best) f |= TB_D;
            tb[i * W + j] = f;
            diag = up;
            row[j] = best;
        }
    }
    uint32_t best; size_t col;
    pick_end(row, n, rightmost, &best, &col);
    free(row);
    if (best > max_errors) { free(tb); return 0; }

    /* traceback; ops collected back-to-front */
    uint8_t* ops = (uint8_t*)malloc(m + n + 1);
    if (!ops) { free(tb); return -1; }
    size_t n_ops = 0, i = m, j = col;
    while (i > 0) {                                           /* row 0 == trace_directions::none (semi-global) */
        uint8_t const f = tb[i * W + j];
        char take = 0;
        for (const char* p = priority; *p && !take; ++p) {
            if (*p == 'L' && (f & TB_L)) take = 'L';
            else if (*p == 'U' && (f & TB_U)) take = 'U';
            else if (*p == 'D' && (f & TB_D)) take = 'D';
        }
        if (take == 'L') { ops[n_ops++] = FXO_CIGAR_D; --j; }
        else if (take == 'U') { ops[n_ops++] = FXO_CIGAR_I; --i; }
        else if (take == 'D') { ops[n_ops++] = (query[i - 1] == reference[j - 1]) ? FXO_CIGAR_EQ : FXO_CIGAR_X; --i; --j; }
        else { free(tb); free(ops); return -1; }              /* impossible in an unbanded matrix */
    }
    free(tb);
    if (num_errors) *num_errors = best;
    if (start_in_window) *start_in_window = j;                /* sequence1_begin_position (alignment.cpp:175) */

    /* run-length encode front-to-back: cigar_from_alignment(aln, {}, extended = true) (alignment.cpp:178) */
    size_t out = 0;
    for (size_t p = n_ops; p > 0;) {
        uint8_t const op = ops[p - 1];
        uint32_t run = 0;
        while (p > 0 && ops[p - 1] == op) { ++run; --p; }
        if (out >= cigar_cap) { free(ops); return -1; }
        cigar[out++] = (run << 4) | op;
    }
    free(ops);
    if (cigar_len) *cigar_len = out;
    return 1;
}

int fxo_align(const uint8_t* reference, size_t n, const uint8_t* query, size_t m, size_t max_errors,
              int mode, uint64_t* num_errors, uint64_t* start_in_window,
              uint32_t* cigar, size_t cigar_cap, size_t* cigar_len) {
    return fxo_align_ex(reference, n, query, m, max_errors, mode, FXO_TRACE_PRIORITY, 1,
                        num_errors, start_in_window, cigar, cigar_cap, cigar_len);
}

/* ------------------------------------------------------------------ verification.cpp:8-245 */

struct fxo_verifier {
    size_t n_references;
    const uint8_t* const* ref;
    const uint64_t* ref_len;
    const fxo_node* inner; size_t n_inner;
    const fxo_node* leaves; size_t n_leaves;
    int kind, without_cigar;
    double ratio;
    fxo_intervals** iv[2];            /* [orientation][reference]  (parallelization.hpp:47-48) */
    fxo_alignment* alns; size_t n_alns, cap_alns;
    uint32_t* cigars; size_t n_cig, cap_cig;
    fxo_stats stats;
};

fxo_verifier* fxo_verifier_new(size_t n_references, const uint8_t* const* reference_ranks,
                               const uint64_t* reference_lengths,
                               const fxo_node* inner, size_t n_inner, const fxo_node* leaves, size_t n_leaves,
                               int kind, int interval_optimization, double extra_verification_ratio,
                               int without_cigar) {
    fxo_verifier* v = (fxo_verifier*)calloc(1, sizeof *v);
    if (!v) return NULL;
    v->n_references = n_references; v->ref = reference_ranks; v->ref_len = reference_lengths;
    v->inner = inner; v->n_inner = n_inner; v->leaves = leaves; v->n_leaves = n_leaves;
    v->kind = kind; v->without_cigar = without_cigar; v->ratio = extra_verification_ratio;
    for (int o = 0; o < 2; ++o) {
        v->iv[o] = (fxo_intervals**)calloc(n_references ? n_references : 1, sizeof(fxo_intervals*));
        for (size_t r = 0; r < n_references; ++r) v->iv[o][r] = fxo_intervals_new(interval_optimization);
    }
    return v;
}

void fxo_verifier_free(fxo_verifier* v) {
    if (!v) return;
    for (int o = 0; o < 2; ++o) {
        for (size_t r = 0; r < v->n_references; ++r) fxo_intervals_free(v->iv[o][r]);
        free(v->iv[o]);
    }
    free(v->alns); free(v->cigars); free(v);
}

void fxo_verifier_configure_intervals(fxo_verifier* v, int active) {
    for (int o = 0; o < 2; ++o)
        for (size_t r = 0; r < v->n_references; ++r) fxo_intervals_configure(v->iv[o][r], active);
}

size_t fxo_verifier_num_alignments(const fxo_verifier* v) { return v->n_alns; }
const fxo_alignment* fxo_verifier_alignments(const fxo_verifier* v) { return v->alns; }
const uint32_t* fxo_verifier_cigar_pool(const fxo_verifier* v) { return v->cigars; }
const fxo_stats* fxo_verifier_stats(const fxo_verifier* v) { return &v->stats; }

static const fxo_node* tree_root(const fxo_verifier* v) {             /* pex.cpp:64-68 */
    return v->n_inner == 0 ? &v->leaves[0] : &v->inner[0];
}

/* verification.cpp:186-245.  Returns 1 exists, 0 none, -1 error. */
static int try_to_align(fxo_verifier* v, const fxo_node* node, uint64_t reference_id, fxo_span span,
                        const uint8_t* query, int orientation) {
    const uint8_t* q = query + node->query_index_from;
    size_t const m = (size_t)node_len(node);
    const uint8_t* r = v->ref[reference_id] + span.offset;
    size_t const n = (size_t)span.length;
    int const is_root = node_is_root(node);
    int mode = FXO_MODE_EXISTS;
    if (is_root) mode = v->without_cigar ? FXO_MODE_NO_CIGAR : FXO_MODE_CIGAR;

    uint64_t errs = 0, start = 0;
    size_t clen = 0;
    size_t const need = m + n + 2;
    if (mode == FXO_MODE_CIGAR && v->n_cig + need > v->cap_cig) {
        size_t nc = v->cap_cig ? v->cap_cig : 1024;
        while (nc < v->n_cig + need) nc *= 2;
        uint32_t* np = (uint32_t*)realloc(v->cigars, nc * sizeof *np);
        if (!np) return -1;
        v->cigars = np; v->cap_cig = nc;
    }
    int const rc = fxo_align(r, n, q, m, (size_t)node->num_errors, mode, &errs, &start,
                             mode == FXO_MODE_CIGAR ? v->cigars + v->n_cig : NULL,
                             mode == FXO_MODE_CIGAR ? need : 0, &clen);
    if (rc < 0) return -1;
    if (rc == 1 && mode != FXO_MODE_EXISTS) {                       /* alignments.insert(...), :228-236 */
        if (v->n_alns == v->cap_alns) {
            size_t const nc = v->cap_alns ? 2 * v->cap_alns : 16;
            fxo_alignment* na = (fxo_alignment*)realloc(v->alns, nc * sizeof *na);
            if (!na) return -1;
            v->alns = na; v->cap_alns = nc;
        }
        fxo_alignment a;
        a.reference_id = reference_id;
        a.start_in_reference = span.offset + start;                  /* alignment.cpp:139,175 */
        a.num_errors = errs;
        a.orientation = (uint32_t)orientation;
        a.cigar_len = (uint32_t)clen;
        a.cigar_offset = v->n_cig;
        v->n_cig += clen;
        v->alns[v->n_alns++] = a;
    }
    if (is_root) { v->stats.n_aligned_root++; v->stats.sum_aligned_root += span.length; v->stats.cells_root += (uint64_t)m * n; }
    else { v->stats.n_aligned_inner++; v->stats.sum_aligned_inner += span.length; v->stats.cells_inner += (uint64_t)m * n; }
    return rc;
}

static fxo_interval span_interval(fxo_span s) { fxo_interval i = { s.offset, s.offset + s.length }; return i; } /* :150-155 */

/* verification.cpp:119-136 */
static int root_was_already_verified(fxo_verifier* v, const fxo_anchor* a, const fxo_node* leaf, int orientation) {
    fxo_span const rs = fxo_compute_span(a->reference_position, tree_root(v), leaf->query_index_from,
                                         v->ref_len[a->reference_id], v->ratio);
    fxo_interval const t = fxo_interval_trim(span_interval(rs), rs.extra);
    if (fxo_intervals_contains(v->iv[orientation][a->reference_id], t)) {
        v->stats.n_avoided_root++; v->stats.sum_avoided_root += rs.length;
        return 1;
    }
    return 0;
}

static int verify_one(fxo_verifier* v, const uint8_t* query, int orientation, const fxo_anchor* a) {
    const fxo_node* leaf = &v->leaves[a->pex_leaf_index];
    const fxo_node* root = tree_root(v);
    fxo_intervals* ivs = v->iv[orientation][a->reference_id];
    uint64_t const ref_len = v->ref_len[a->reference_id];

    if (v->kind == FXO_KIND_DIRECT_FULL) {                            /* verification.cpp:23-42 */
        if (root_was_already_verified(v, a, leaf, orientation)) return 0;
        fxo_span const rs = fxo_compute_span(a->reference_position, root, leaf->query_index_from, ref_len, v->ratio);
        if (try_to_align(v, root, a->reference_id, rs, query, orientation) < 0) return -1;
        fxo_intervals_insert(ivs, span_interval(rs));
        return 0;
    }
    if (v->kind != FXO_KIND_HIERARCHICAL) return -1;                  /* verification.cpp:19 throws */

    /* verification.cpp:44-117 */
    if (root_was_already_verified(v, a, leaf, orientation)) return 0;
    fxo_span const rs = fxo_compute_span(a->reference_position, root, leaf->query_index_from, ref_len, v->ratio);
    if (node_is_root(leaf)) {
        if (try_to_align(v, leaf, a->reference_id, rs, query, orientation) < 0) return -1;
        fxo_intervals_insert(ivs, span_interval(rs));
        return 0;
    }
    uint64_t const seed_from = leaf->query_index_from;
    const fxo_node* curr = &v->inner[leaf->parent_id];
    for (;;) {
        int const is_root = node_is_root(curr);
        fxo_span const sp = fxo_compute_span(a->reference_position, curr, seed_from, ref_len, is_root ? v->ratio : 0.0);
        if (sp.length > 512 && root_was_already_verified(v, a, leaf, orientation)) return 0;   /* :85-93 */
        int const outcome = try_to_align(v, curr, a->reference_id, sp, query, orientation);
        if (outcome < 0) return -1;
        if (is_root) fxo_intervals_insert(ivs, span_interval(sp));
        if (outcome == 0 || is_root) break;
        curr = &v->inner[curr->parent_id];
    }
    return 0;
}

int fxo_verifier_run(fxo_verifier* v, const uint8_t* query, size_t query_len, int orientation,
                     const fxo_anchor* anchors, size_t n_anchors) {
    (void)query_len;
    for (size_t i = 0; i < n_anchors; ++i) {
        if (anchors[i].pex_leaf_index >= v->n_leaves || anchors[i].reference_id >= v->n_references) return -1;
        if (verify_one(v, query, orientation, &anchors[i]) < 0) return -1;
    }
    return 0;
}
