// Minimal stand-ins for the reference's types on the hot-path boundary (floxer 0.2.0), with the reference's names, member
// names and meaning, so that the shim in floxer_shim.hpp is the code a floxer maintainer would add.  TEST INFRASTRUCTURE:
// the real headers need SeqAn3 and the other un-vendored dependencies; a BAM-encoded uint32_t stands in for
// seqan3::cigar.  Citations are file:line into the reference tree.
#pragma once
#include <cstddef>
#include <cstdint>
#include <limits>
#include <optional>
#include <span>
#include <stdexcept>
#include <string>
#include <vector>

namespace alignment {                                   // include/alignment.hpp
enum class query_orientation { forward, reverse_complement };                                       // :14-16
using cigar_op = uint32_t;                                                                            // seqan3::cigar stand-in: len << 4 | op
struct query_alignment { size_t start_in_reference; size_t num_errors; query_orientation orientation; std::vector<cigar_op> cigar; };   // :18-23
class query_alignments {                                                                              // :28-51, src/lib/alignment.cpp:37-79
public:
    explicit query_alignments(size_t num_references) : alignments_per_reference(num_references) {}
    void insert(query_alignment alignment, size_t reference_id) {
        if (!best || alignment.num_errors < *best) best = alignment.num_errors;
        alignments_per_reference[reference_id].emplace_back(std::move(alignment));
    }
    std::vector<query_alignment> const& to_reference(size_t reference_id) const { return alignments_per_reference[reference_id]; }
    std::optional<size_t> best_num_errors() const { return best; }
    size_t size() const { size_t n = 0; for (auto const& v : alignments_per_reference) n += v.size(); return n; }
private:
    std::optional<size_t> best;
    std::vector<std::vector<query_alignment>> alignments_per_reference;
};
enum class alignment_mode { only_verify_existance, verify_and_return_alignment_with_cigar, verify_and_return_alignment_without_cigar };   // :53-55
struct alignment_config { size_t reference_span_offset; size_t num_allowed_errors; query_orientation orientation; alignment_mode mode; };  // :57-62
enum class alignment_outcome { alignment_exists, no_adequate_alignment_exists };                    // :64-66
struct alignment_result { alignment_outcome outcome; std::optional<query_alignment> alignment = std::nullopt; };                         // :68-71
alignment_result align(std::span<const uint8_t> reference, std::span<const uint8_t> query, alignment_config const& config);             // :73-77
}  // namespace alignment

namespace search {                                      // include/search.hpp:27-38
struct anchor_t { size_t pex_leaf_index; size_t reference_id; size_t reference_position; size_t num_errors; };
}

namespace input {                                       // include/input.hpp:16-20
struct reference_record { std::string id; std::vector<uint8_t> rank_sequence; size_t internal_id; };
}

namespace pex {                                         // include/pex.hpp
enum class verification_kind_t { direct_full, hierarchical };                                        // :43-45
class pex_tree {
public:
    struct node {                                                                                     // :59-76
        static constexpr size_t null_id = std::numeric_limits<size_t>::max();
        size_t parent_id; size_t query_index_from; size_t query_index_to; size_t num_errors;
        size_t length_of_query_span() const { return query_index_to - query_index_from + 1; }
        bool is_root() const { return parent_id == null_id; }
    };
    pex_tree(std::vector<node> inner, std::vector<node> leaf_nodes) : inner_nodes(std::move(inner)), leaves(std::move(leaf_nodes)) {}
    node const& root() const { return inner_nodes.empty() ? leaves.at(0) : inner_nodes.at(0); }      // src/lib/pex.cpp:64-68
    std::vector<node> const& get_leaves() const { return leaves; }                                    // :78-80
    std::vector<node> const& get_inner_nodes() const { return inner_nodes; }                          // (the accessor INTEGRATION.md asks for)
private:
    std::vector<node> inner_nodes, leaves;
};
}  // namespace pex

namespace statistics {                                  // the three histograms the hot path feeds (src/lib/verification.cpp:130,239,241)
struct search_and_alignment_statistics {
    std::vector<size_t> spans_inner, spans_root, spans_avoided;
    void add_reference_span_size_aligned_inner_node(size_t v) { spans_inner.push_back(v); }
    void add_reference_span_size_aligned_root(size_t v) { spans_root.push_back(v); }
    void add_reference_span_size_avoided_root(size_t v) { spans_avoided.push_back(v); }
};
}

namespace intervals { class verified_intervals {}; }    // include/intervals.hpp:60-81 (kept by the GPU path internally per call)

namespace verification {                                // include/verification.hpp:22-48
struct query_verifier {
    pex::pex_tree const& pex_tree;
    search::anchor_t const& anchor;
    pex::pex_tree::node const& pex_leaf_node;
    std::span<const uint8_t> const query;
    alignment::query_orientation const orientation;
    input::reference_record const& reference;
    intervals::verified_intervals& already_verified_intervals;
    double const extra_verification_ratio;
    bool const without_cigar;
    pex::verification_kind_t const kind;
    bool const interval_optimization;
    alignment::query_alignments& alignments;
    statistics::search_and_alignment_statistics& stats;
    void verify();                                       // "should only be called once on each instance" (:23)
};
}  // namespace verification
