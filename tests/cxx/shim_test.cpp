// Compiles the reference-side binding (floxer_shim.hpp) against include/floxer_gpu.h, links libfloxer_gpu.so and runs the
// reference's own known-answer cases through it: test/alignment_test.cpp:7-30 and test/verification_test.cpp:11-123.
//   shim_test --link-only   checks that the library loads and every call resolves (no GPU needed)
//   shim_test               runs the cases on device 0; exit code 0 = all as the reference expects
#include "floxer_shim.hpp"

#include <cstdio>
#include <cstring>

static std::string cigar_string(std::vector<alignment::cigar_op> const& ops) {
    std::string s;
    for (uint32_t op : ops) { s += std::to_string(op >> 4); s += (op & 15) == FXG_CIGAR_I ? 'I' : (op & 15) == FXG_CIGAR_D ? 'D' : (op & 15) == FXG_CIGAR_EQ ? '=' : 'X'; }
    return s;
}
#define EXPECT(cond) do { if (!(cond)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); ++failures; } } while (0)

int main(int argc, char** argv) {
    int failures = 0;
    if (argc > 1 && std::strcmp(argv[1], "--link-only") == 0) {
        std::printf("%s\n", fxg_version());
        // (taking their addresses is enough to make the linker resolve them)
        void* used[] = {(void*)&fxg_create, (void*)&fxg_destroy, (void*)&fxg_set_references, (void*)&fxg_align_batch, (void*)&fxg_verify_reads,
                        (void*)&fxg_job_alignments, (void*)&fxg_job_cigar_pool, (void*)&fxg_job_stats, (void*)&fxg_job_free, (void*)&fxg_pex_build};
        return used[0] ? 0 : 1;
    }
    if (fxg_create(0, &floxer_gpu::context()) != FXG_OK) { std::fprintf(stderr, "no CUDA device\n"); return 2; }

    // ---- test/alignment_test.cpp:7-30: align in with-CIGAR mode ----
    {
        std::vector<uint8_t> reference{0, 0, 1, 2, 1, 3, 0, 2, 2, 3, 0, 1}, query{1, 2, 1, 3, 1, 2, 2};
        alignment::alignment_config config{.reference_span_offset = 0, .num_allowed_errors = 2, .orientation = alignment::query_orientation::forward,
                                           .mode = alignment::alignment_mode::verify_and_return_alignment_with_cigar};
        auto const result = alignment::align(reference, query, config);
        EXPECT(result.outcome == alignment::alignment_outcome::alignment_exists);
        EXPECT(result.alignment && result.alignment->num_errors == 1 && result.alignment->start_in_reference == 2);
        EXPECT(result.alignment && cigar_string(result.alignment->cigar) == "4=1X2=");
        config.mode = alignment::alignment_mode::only_verify_existance; config.num_allowed_errors = 0;
        EXPECT(alignment::align(reference, query, config).outcome == alignment::alignment_outcome::no_adequate_alignment_exists);
    }
    // ---- test/verification_test.cpp:11-123: query_verifier::verify, hierarchical, then direct_full ----
    {
        std::vector<uint8_t> ref{4, 2, 3, 4, 3, 4, 4, 4, 3, 2, 4, 3, 3, 2, 2, 3, 4, 4, 3, 3, 4, 3, 2, 2, 1, 4, 3, 3, 4, 2, 4, 4, 4, 3, 3, 2, 1, 1, 1, 2,
                                 3, 4, 4, 3, 2, 4, 4, 2, 1, 4, 4, 3, 4, 4, 4, 4, 3, 3, 2, 1, 2, 3, 4, 3, 2, 1, 2, 3, 4, 3, 1, 4, 2, 1, 4, 4, 2, 2, 3, 4,
                                 3, 3, 2, 1, 4, 4, 1, 1, 1, 2, 4, 3, 2, 1, 2, 2, 2, 3, 3, 1};
        std::vector<uint8_t> query{4, 3, 4, 4, 4, 4, 3, 3, 2, 1, 4, 2, 3, 4, 3, 2, 1, 2, 3, 4, 1, 4, 2, 1, 4, 4, 2, 2, 3, 4};
        input::reference_record reference{"name", ref, 0};
        const uint8_t* ptrs[1] = {ref.data()}; uint64_t lens[1] = {ref.size()};
        floxer_gpu::check(fxg_set_references(floxer_gpu::context(), 1, ptrs, lens));
        // the bottom-up tree of (30, 5, 1), by the library's builder (checked against test/pex_test.cpp in the Python suite)
        fxg_pex_node *inner = nullptr, *leaves = nullptr; size_t n_inner = 0, n_leaves = 0;
        floxer_gpu::check(fxg_pex_build(30, 5, 1, 1, &inner, &n_inner, &leaves, &n_leaves));
        auto as_nodes = [](const fxg_pex_node* p, size_t n) { std::vector<pex::pex_tree::node> v(n); std::memcpy(v.data(), p, n * sizeof(fxg_pex_node)); return v; };
        pex::pex_tree tree(as_nodes(inner, n_inner), as_nodes(leaves, n_leaves));
        fxg_pex_free(inner); fxg_pex_free(leaves);
        search::anchor_t anchor{.pex_leaf_index = 0, .reference_id = 0, .reference_position = 50, .num_errors = 0};
        intervals::verified_intervals ivls;
        alignment::query_alignments alignments(1);
        statistics::search_and_alignment_statistics stats;
        for (auto kind : {pex::verification_kind_t::hierarchical, pex::verification_kind_t::direct_full}) {
            verification::query_verifier verifier{.pex_tree = tree, .anchor = anchor, .pex_leaf_node = tree.get_leaves().at(0), .query = query,
                                                  .orientation = alignment::query_orientation::reverse_complement, .reference = reference,
                                                  .already_verified_intervals = ivls, .extra_verification_ratio = 0.1, .without_cigar = false, .kind = kind,
                                                  .interval_optimization = false, .alignments = alignments, .stats = stats};
            verifier.verify();
            floxer_gpu::flush_package();
        }
        EXPECT(alignments.size() == 2);                                       // the hierarchical one and the identical direct one (:94-112)
        for (auto const& a : alignments.to_reference(0)) {
            EXPECT(cigar_string(a.cigar) == "10=1I9=1D10=");
            EXPECT(a.num_errors == 2 && a.start_in_reference == 50);
            EXPECT(a.orientation == alignment::query_orientation::reverse_complement);
        }
        EXPECT(alignments.best_num_errors() && *alignments.best_num_errors() == 2);
        EXPECT(stats.spans_root.size() == 2);
        // after four more mutations nothing is added (:114-122)
        query[5] = 1; query[6] = 1; query[11] = 3; query[20] = 2;
        verification::query_verifier verifier{.pex_tree = tree, .anchor = anchor, .pex_leaf_node = tree.get_leaves().at(0), .query = query,
                                              .orientation = alignment::query_orientation::reverse_complement, .reference = reference,
                                              .already_verified_intervals = ivls, .extra_verification_ratio = 0.1, .without_cigar = false,
                                              .kind = pex::verification_kind_t::hierarchical, .interval_optimization = false, .alignments = alignments, .stats = stats};
        verifier.verify();
        floxer_gpu::flush_package();
        EXPECT(alignments.size() == 2);
    }
    fxg_destroy(floxer_gpu::context());
    if (failures) { std::fprintf(stderr, "%d expectation(s) failed\n", failures); return 1; }
    std::printf("shim_test: all expectations met\n");
    return 0;
}
