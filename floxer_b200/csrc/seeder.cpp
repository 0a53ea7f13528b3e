// seeder.cpp -- a q-gram seeder behind the anchor interface of the verification path (row N2 of SURVEY 8f).
//
// The reference finds its anchors with approximate FM-index search of every PEX leaf under optimum search schemes
// (search::searcher::search_seeds, src/lib/search.cpp:143-324; fmindex-collection, not vendored and not buildable offline).
// This is a host-side STAND-IN with the same observable contract, built the simple way: a q-gram index of the references
// (counting sort of all q-grams over A, C, G, T) and, per leaf with error budget e, the pigeonhole rule -- an occurrence
// with at most e edits contains one of e + 1 pieces of the leaf unchanged, shifted by at most e -- followed by a banded
// edit-distance check of every candidate start.  What it reports per leaf is what the reference's search reports after
// locating: every reference position p where the leaf matches some prefix of reference[p..] with d <= e edits, as
// search::anchor_t {pex_leaf_index, reference_id, reference_position, num_errors = d} (include/search.hpp:27-31), then
//   * max_num_anchors_hard: a seed with more raw anchors is dropped altogether (search.cpp:186-199),
//   * max_num_anchors_soft: at most that many are kept -- fewest errors first here; the reference takes them round-robin
//     from its FM-index cursors, an order that depends on the suffix array (search.cpp:226-275),
//   * erase_useless_anchors: an anchor with a better neighbour within the difference of their error counts goes
//     (search.cpp:352-389, anchor_t::is_better_than :39-45), per seed and reference, in position order,
// in the order seed -> reference -> position (search.cpp:78-100).
// Different from the reference by construction: the order in which the soft cap picks among too many anchors.
#include "../../include/floxer_gpu.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>
#include <thread>
#include <vector>

struct fxg_seeder {
    uint32_t q = 0;
    std::vector<const uint8_t*> refs;
    std::vector<uint64_t> lens, base;          // base[r] = position of reference r in the concatenation
    std::vector<uint32_t> bucket;              // 4^q + 1 offsets into `pos`
    std::vector<uint32_t> pos;                 // concatenated positions of the q-grams, by code
    std::vector<std::vector<uint8_t>> copy;    // the references (the index outlives the caller's arrays)
};

namespace {

// code of the q-gram at p, or false if it holds a rank outside A, C, G, T (1..4)
inline bool gram_code(const uint8_t* s, uint32_t q, uint32_t& code) {
    uint32_t c = 0;
    for (uint32_t i = 0; i < q; ++i) {
        uint32_t const r = s[i];
        if (r < 1 || r > 4) return false;
        c = (c << 2) | (r - 1);
    }
    code = c;
    return true;
}

// min over prefixes t of `text` (at most text_len bases) of the edit distance of `seed` and t, if <= e; else e + 1.
// Banded: cells with |i - j| > e cannot lie on an alignment with <= e edits.
inline uint32_t prefix_distance(const uint8_t* seed, uint32_t m, const uint8_t* text, uint32_t text_len, uint32_t e) {
    uint32_t const inf = e + 1;
    uint32_t const width = 2 * e + 1;
    // row i holds columns j = i - e .. i + e at index j - i + e
    uint32_t prev[2 * 8 + 3], cur[2 * 8 + 3];
    if (e > 8) return inf;
    for (uint32_t x = 0; x < width; ++x) { int64_t const j = int64_t(x) - e; prev[x] = j >= 0 && j <= int64_t(e) ? uint32_t(j) : inf; }   // row 0: D[0][j] = j
    for (uint32_t i = 1; i <= m; ++i) {
        uint32_t row_min = inf;
        for (uint32_t x = 0; x < width; ++x) {
            int64_t const j = int64_t(i) + int64_t(x) - e;
            uint32_t v = inf;
            if (j >= 0 && j <= int64_t(text_len)) {
                if (j == 0) v = i <= e ? i : inf;
                else {
                    uint32_t const diag = prev[x] + (seed[i - 1] != text[j - 1] ? 1u : 0u);        // (i-1, j-1) sits at the same index of the row above
                    uint32_t const up = x + 1 < width ? prev[x + 1] + 1 : inf;                   // (i-1, j)
                    uint32_t const left = x > 0 ? cur[x - 1] + 1 : inf;                          // (i, j-1)
                    v = std::min(std::min(diag, up), left);
                    if (v > inf) v = inf;
                }
            }
            cur[x] = v;
            row_min = std::min(row_min, v);
        }
        if (row_min >= inf) return inf;
        std::memcpy(prev, cur, width * sizeof(uint32_t));
    }
    uint32_t best = inf;
    for (uint32_t x = 0; x < width; ++x) best = std::min(best, prev[x]);
    return best;
}

struct Raw { uint64_t position; uint32_t reference, errors; };

// search.cpp:352-389 on the anchors of one seed and reference, sorted by position
void erase_useless(std::vector<Raw>& a, size_t lo, size_t hi) {
    constexpr uint32_t kErase = 0xffffffffu;
    auto better = [](Raw const& x, Raw const& y) {                 // anchor_t::is_better_than with the reference's size_t arithmetic
        uint64_t const dist = x.position < y.position ? y.position - x.position : x.position - y.position;
        return uint64_t(x.errors) <= uint64_t(y.errors) && dist <= uint64_t(y.errors) - uint64_t(x.errors);
    };
    if (hi - lo < 2) return;
    for (size_t cur = lo; cur < hi - 1;) {
        size_t other = cur + 1;
        while (other < hi && better(a[cur], a[other])) { a[other].errors = kErase; ++other; }
        if (other < hi && better(a[other], a[cur])) a[cur].errors = kErase;
        cur = other;
    }
}

}  // namespace

extern "C" {

int fxg_seeder_create(size_t n_refs, const uint8_t* const* ranks, const uint64_t* lens, uint32_t q, fxg_seeder** out) {
    if (!out || (n_refs && (!ranks || !lens)) || q < 4 || q > 14) return FXG_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    fxg_seeder* s = new (std::nothrow) fxg_seeder();
    if (!s) return FXG_ERR_OUT_OF_MEMORY;
    try {
        s->q = q;
        uint64_t total = 0;
        for (size_t r = 0; r < n_refs; ++r) {
            s->copy.emplace_back(ranks[r], ranks[r] + lens[r]);
            s->refs.push_back(s->copy.back().data()); s->lens.push_back(lens[r]); s->base.push_back(total);
            total += lens[r];
        }
        if (total >= (uint64_t(1) << 32)) { delete s; return FXG_ERR_INVALID_ARGUMENT; }      // positions are 32-bit: index references below 4 Gbp
        size_t const n_codes = size_t(1) << (2 * q);
        s->bucket.assign(n_codes + 1, 0);
        for (int pass = 0; pass < 2; ++pass) {
            for (size_t r = 0; r < n_refs; ++r) {
                if (lens[r] < q) continue;
                const uint8_t* t = s->refs[r];
                for (uint64_t p = 0; p + q <= lens[r]; ++p) {
                    uint32_t code;
                    if (!gram_code(t + p, q, code)) continue;
                    if (pass == 0) s->bucket[code + 1]++;
                    else s->pos[s->bucket[code]++] = uint32_t(s->base[r] + p);
                }
            }
            if (pass == 0) {
                for (size_t c = 0; c < n_codes; ++c) s->bucket[c + 1] += s->bucket[c];
                s->pos.resize(s->bucket[n_codes]);
            } else {
                for (size_t c = n_codes; c > 0; --c) s->bucket[c] = s->bucket[c - 1];        // the fill moved every offset to its bucket's end
                s->bucket[0] = 0;
            }
        }
    } catch (...) { delete s; return FXG_ERR_OUT_OF_MEMORY; }
    *out = s;
    return FXG_OK;
}

void fxg_seeder_free(fxg_seeder* s) { delete s; }

int fxg_seeder_search(const fxg_seeder* s, const uint8_t* query, size_t query_len, const fxg_pex_node* leaves, size_t n_leaves,
                      uint64_t max_anchors_hard, uint64_t max_anchors_soft, int erase_useless_anchors, fxg_anchor** anchors, size_t* n_anchors) {
    if (!s || !anchors || !n_anchors || (n_leaves && !leaves) || (query_len && !query)) return FXG_ERR_INVALID_ARGUMENT;
    *anchors = nullptr; *n_anchors = 0;
    try {
        std::vector<fxg_anchor> out;
        std::vector<Raw> raw;
        std::vector<uint64_t> cand;
        uint32_t const q = s->q;
        for (size_t li = 0; li < n_leaves; ++li) {
            fxg_pex_node const& leaf = leaves[li];
            if (leaf.query_index_to < leaf.query_index_from || leaf.query_index_to >= query_len) return FXG_ERR_INVALID_ARGUMENT;
            uint32_t const m = uint32_t(leaf.query_index_to - leaf.query_index_from + 1), e = uint32_t(leaf.num_errors);
            if (e > 8 || m / (e + 1) < q) return FXG_ERR_INVALID_ARGUMENT;        // pieces shorter than the index's q-grams: build the index with a smaller q
            const uint8_t* seed = query + leaf.query_index_from;
            // ---- candidate starts (positions in the concatenation): piece k of e + 1 found unchanged, shifted by at most e ----
            cand.clear();
            for (uint32_t k = 0; k <= e; ++k) {
                uint32_t const off = uint32_t(uint64_t(m) * k / (e + 1));
                uint32_t code;
                if (!gram_code(seed + off, q, code)) continue;
                for (uint32_t x = s->bucket[code]; x < s->bucket[code + 1]; ++x) {
                    int64_t const nominal = int64_t(s->pos[x]) - int64_t(off);
                    for (int64_t d = -int64_t(e); d <= int64_t(e); ++d) if (nominal + d >= 0) cand.push_back(uint64_t(nominal + d));
                }
            }
            std::sort(cand.begin(), cand.end());
            cand.erase(std::unique(cand.begin(), cand.end()), cand.end());
            // ---- every candidate start against the leaf ----
            raw.clear();
            for (uint64_t g : cand) {
                size_t r = size_t(std::upper_bound(s->base.begin(), s->base.end(), g) - s->base.begin()) - 1;
                uint64_t const p = g - s->base[r];
                if (p >= s->lens[r]) continue;
                uint32_t const room = uint32_t(std::min<uint64_t>(s->lens[r] - p, uint64_t(m) + e));
                uint32_t const d = prefix_distance(seed, m, s->refs[r] + p, room, e);
                if (d <= e) raw.push_back(Raw{p, uint32_t(r), d});
            }
            if (raw.size() > max_anchors_hard) continue;                      // the seed is fully excluded (search.cpp:186-199)
            if (raw.size() > max_anchors_soft) {
                std::stable_sort(raw.begin(), raw.end(), [](Raw const& a, Raw const& b) { return a.errors < b.errors; });
                raw.resize(size_t(max_anchors_soft));
            }
            std::sort(raw.begin(), raw.end(), [](Raw const& a, Raw const& b) { return a.reference != b.reference ? a.reference < b.reference : a.position < b.position; });
            if (erase_useless_anchors) {
                size_t lo = 0;
                while (lo < raw.size()) {
                    size_t hi = lo + 1;
                    while (hi < raw.size() && raw[hi].reference == raw[lo].reference) ++hi;
                    erase_useless(raw, lo, hi);
                    lo = hi;
                }
            }
            for (Raw const& a : raw) if (a.errors != 0xffffffffu) out.push_back(fxg_anchor{uint64_t(li), uint64_t(a.reference), a.position, uint64_t(a.errors)});
        }
        if (!out.empty()) {
            fxg_anchor* p = static_cast<fxg_anchor*>(std::malloc(out.size() * sizeof(fxg_anchor)));
            if (!p) return FXG_ERR_OUT_OF_MEMORY;
            std::memcpy(p, out.data(), out.size() * sizeof(fxg_anchor));
            *anchors = p; *n_anchors = out.size();
        }
    } catch (...) { return FXG_ERR_OUT_OF_MEMORY; }
    return FXG_OK;
}

}  // extern "C"
