// The reference-side binding of the B200 path: definitions of alignment::align and verification::query_verifier::verify
// with the reference's signatures (include/alignment.hpp:73-77, include/verification.hpp:22-48) on top of the C ABI in
// include/floxer_gpu.h.  This is the code INTEGRATION.md sections 3 and 4 show, compiled.
//
// verify() is DEFERRED: the reference calls it once per anchor inside the anchor loop of a verification task
// (src/lib/parallelization.cpp:230-249) and reads its results (alignments, statistics) only after the loop, so the shim
// queues the anchor and floxer_gpu::flush_package(), called right behind the loop, verifies the package in one
// fxg_verify_reads call -- hierarchical walk, interval optimisation and all -- and scatters the results into the
// `alignments` / `stats` objects the verifiers were given.  Errors become std::runtime_error, the reference's convention
// (src/lib/verification.cpp:19, src/lib/parallelization.cpp:282-289).
#pragma once
#include "floxer_gpu.h"
#include "reference_stubs.hpp"

#include <algorithm>

static_assert(sizeof(fxg_pex_node) == sizeof(pex::pex_tree::node) && sizeof(fxg_anchor) == sizeof(search::anchor_t), "layouts (pex.hpp:59-70, search.hpp:27-31)");

namespace floxer_gpu {

inline fxg_ctx*& context() { static fxg_ctx* c = nullptr; return c; }

inline void check(int rc) { if (rc != FXG_OK) throw std::runtime_error(fxg_last_error(context())); }

inline alignment::query_alignment to_query_alignment(uint64_t start, uint32_t errors, uint8_t orientation, const uint32_t* ops, uint32_t n_ops) {
    return alignment::query_alignment{size_t(start), size_t(errors),
                                      orientation == FXG_FORWARD ? alignment::query_orientation::forward : alignment::query_orientation::reverse_complement,
                                      std::vector<alignment::cigar_op>(ops, ops + n_ops)};
}

// the anchors of one verification task (one query, one orientation, one or more references), in call order
struct package {
    const verification::query_verifier* first = nullptr;     // configuration and targets are those of the first verifier
    std::vector<fxg_anchor> anchors;
    std::vector<const input::reference_record*> records;
};
inline package& current_package() { thread_local package p; return p; }

// behind the anchor loop (parallelization.cpp:249): verifies what the loop queued
inline void flush_package() {
    package& P = current_package();
    if (!P.first) return;
    verification::query_verifier const& v = *P.first;
    fxg_verify_config cfg{};
    cfg.extra_verification_ratio = v.extra_verification_ratio;
    cfg.verification_kind = v.kind == pex::verification_kind_t::direct_full ? FXG_KIND_DIRECT_FULL : FXG_KIND_HIERARCHICAL;
    cfg.interval_optimization = v.interval_optimization; cfg.without_cigar = v.without_cigar;
    auto const& inner = v.pex_tree.get_inner_nodes(); auto const& leaves = v.pex_tree.get_leaves();
    std::vector<fxg_pex_node> nodes;
    auto put = [&](auto const& ns) { auto p = reinterpret_cast<const fxg_pex_node*>(ns.data()); nodes.insert(nodes.end(), p, p + ns.size()); };
    put(inner); put(leaves);
    bool const forward = v.orientation == alignment::query_orientation::forward;
    fxg_read r{};
    r.query_len = uint32_t(v.query.size()); r.num_inner = uint32_t(inner.size()); r.num_leaves = uint32_t(leaves.size());
    (forward ? r.num_anchors_forward : r.num_anchors_reverse) = uint32_t(P.anchors.size());
    // the package's query is the pool of its own orientation; the other pool is not read (no anchors there)
    const uint8_t* pool = v.query.data();
    fxg_job* job = nullptr;
    int const rc = fxg_verify_reads(context(), &cfg, &r, 1, pool, pool, v.query.size(), nodes.data(), nodes.size(), P.anchors.data(), P.anchors.size(), &job);
    P = package{};
    check(rc);
    const fxg_alignment* al = fxg_job_alignments(job); const uint32_t* ops = fxg_job_cigar_pool(job);
    for (size_t a = 0; a < fxg_job_num_alignments(job); ++a)
        v.alignments.insert(to_query_alignment(al[a].start_in_reference, al[a].num_errors, al[a].orientation, ops + al[a].cigar_offset, al[a].cigar_len), al[a].reference_id);
    // the histograms receive counts and sums (the GPU path does not keep every single span length)
    const fxg_stats* s = fxg_job_stats(job);
    for (uint64_t i = 0; i < s->n_aligned_inner; ++i) v.stats.add_reference_span_size_aligned_inner_node(i == 0 ? s->sum_aligned_inner - (s->n_aligned_inner - 1) * (s->sum_aligned_inner / s->n_aligned_inner) : s->sum_aligned_inner / s->n_aligned_inner);
    for (uint64_t i = 0; i < s->n_aligned_root; ++i) v.stats.add_reference_span_size_aligned_root(i == 0 ? s->sum_aligned_root - (s->n_aligned_root - 1) * (s->sum_aligned_root / s->n_aligned_root) : s->sum_aligned_root / s->n_aligned_root);
    for (uint64_t i = 0; i < s->n_avoided_root; ++i) v.stats.add_reference_span_size_avoided_root(i == 0 ? s->sum_avoided_root - (s->n_avoided_root - 1) * (s->sum_avoided_root / s->n_avoided_root) : s->sum_avoided_root / s->n_avoided_root);
    fxg_job_free(context(), job);
}

}  // namespace floxer_gpu

// include/verification.hpp:22-48 / src/lib/verification.cpp:8-21
inline void verification::query_verifier::verify() {
    if (kind != pex::verification_kind_t::direct_full && kind != pex::verification_kind_t::hierarchical)
        throw std::runtime_error("Internal bug in verification kind (should not happen)");             // verification.cpp:19
    floxer_gpu::package& P = floxer_gpu::current_package();
    if (!P.first) P.first = this;
    // the anchor as the GPU path wants it: the reference id is the record's internal id (verification.cpp:228-236)
    P.anchors.push_back(fxg_anchor{anchor.pex_leaf_index, reference.internal_id, anchor.reference_position, anchor.num_errors});
}

// include/alignment.hpp:73-77 / src/lib/alignment.cpp:83-181
inline alignment::alignment_result alignment::align(std::span<const uint8_t> reference, std::span<const uint8_t> query, alignment_config const& config) {
    fxg_align_task t{};
    t.ref_offset = 0; t.reference_span_offset = config.reference_span_offset; t.query_offset = 0;
    t.ref_len = uint32_t(reference.size()); t.query_len = uint32_t(query.size());
    t.ref_id = FXG_REF_INLINE;                      // the span is handed over as bytes
    t.max_errors = uint32_t(config.num_allowed_errors);
    t.mode = config.mode == alignment_mode::only_verify_existance ? FXG_MODE_EXISTS
           : config.mode == alignment_mode::verify_and_return_alignment_without_cigar ? FXG_MODE_NO_CIGAR : FXG_MODE_CIGAR;
    t.orientation = config.orientation == query_orientation::forward ? FXG_FORWARD : FXG_REVERSE_COMPLEMENT;
    fxg_align_result r{};
    std::vector<uint32_t> ops(2 * config.num_allowed_errors + 3);
    size_t used = 0;
    floxer_gpu::check(fxg_align_batch(floxer_gpu::context(), &t, 1, query.data(), query.size(), reference.data(), reference.size(), &r, ops.data(), ops.size(), &used));
    if (!r.exists) return {.outcome = alignment_outcome::no_adequate_alignment_exists};
    if (t.mode == FXG_MODE_EXISTS) return {.outcome = alignment_outcome::alignment_exists};
    return {.outcome = alignment_outcome::alignment_exists,
            .alignment = floxer_gpu::to_query_alignment(r.start_in_reference, r.num_errors, r.orientation, ops.data() + r.cigar_offset, r.cigar_len)};
}
