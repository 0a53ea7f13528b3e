// pex_tree.cpp -- host-side PEX tree construction behind fxg_pex_build (include/floxer_gpu.h).
// Follows the published construction (Navarro & Raffinot, "Flexible Pattern Matching in Strings", ch. 6.5.1)
// with the reference's adjustment for leaves that tolerate errors (src/lib/pex.cpp:84-256).
#include "../../include/floxer_gpu.h"

#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

using node = fxg_pex_node;

uint64_t ceil_div(uint64_t a, uint64_t b) { return (a + b - 1) / b; }

// top-down splitting, pex.cpp:110-156; an explicit work stack visits nodes in the same pre-order as the recursion
void build_top_down(uint64_t length, uint64_t errors, uint64_t leaf_errors, std::vector<node>& inner, std::vector<node>& leaves) {
    uint64_t const piece = length / (errors + 1);          // no_error_seed_length, pex.cpp:85
    struct Item { uint64_t from, to, k, parent; };          // 1-based inclusive range, as in the book
    std::vector<Item> stack{{1, length, errors, FXG_NULL_ID}};
    while (!stack.empty()) {
        Item const it = stack.back();
        stack.pop_back();
        node const n{it.parent, it.from - 1, it.to - 1, it.k};
        if (it.k <= leaf_errors) { leaves.push_back(n); continue; }
        uint64_t const id = inner.size();
        inner.push_back(n);
        uint64_t const left = ceil_div(it.k + 1, 2);        // number of error-free pieces that go left
        uint64_t const split = it.from + left * piece;
        uint64_t const k_left = (left * it.k) / (it.k + 1);
        uint64_t const k_right = ((it.k + 1 - left) * it.k) / (it.k + 1);
        stack.push_back({split, it.to, k_right, id});       // right child is visited after the whole left subtree
        stack.push_back({it.from, split - 1, k_left, id});
    }
}

// bottom-up merging, pex.cpp:158-256
void build_bottom_up(uint64_t length, uint64_t errors, uint64_t leaf_errors, std::vector<node>& inner, std::vector<node>& leaves) {
    uint64_t const n_leaves = ceil_div(errors + 1, leaf_errors + 1);
    if (n_leaves == 1) { leaves.push_back(node{FXG_NULL_ID, 0, length - 1, errors}); return; }
    uint64_t const base = length / n_leaves, rem = length % n_leaves;
    uint64_t at = 0;
    for (uint64_t i = 0; i < n_leaves; ++i) {
        uint64_t const len = base + (i < rem ? 1 : 0);
        leaves.push_back(node{0, at, at + len - 1, leaf_errors});
        at += len;
    }
    inner.push_back(node{});                                // index 0 is reserved for the root
    auto merge = [](node* first, size_t count, uint64_t parent_id) {
        uint64_t k = count - 1;
        for (size_t i = 0; i < count; ++i) { first[i].parent_id = parent_id; k += first[i].num_errors; }
        return node{0, first[0].query_index_from, first[count - 1].query_index_to, k};
    };
    // `level` names the nodes being merged: the leaves first, afterwards the tail of `inner`
    bool level_is_leaves = true;
    size_t level_begin = 0, level_size = leaves.size();
    while (level_size > 3) {
        size_t const produced_from = inner.size();
        for (size_t i = 0; i + 1 < level_size;) {
            size_t const remaining = level_size - i;
            size_t const take = remaining == 3 ? 3 : 2;      // an odd level ends with a triple
            node* first = (level_is_leaves ? leaves.data() : inner.data()) + level_begin + i;
            node const parent = merge(first, take, inner.size());
            inner.push_back(parent);                         // may reallocate: `first` is not used afterwards
            i += take;
        }
        level_is_leaves = false;
        level_begin = produced_from;
        level_size = inner.size() - produced_from;
    }
    node* first = (level_is_leaves ? leaves.data() : inner.data()) + level_begin;
    node root = merge(first, level_size, 0);
    root.parent_id = FXG_NULL_ID;
    inner[0] = root;
}

node* to_c_array(std::vector<node> const& v) {
    node* p = static_cast<node*>(std::malloc((v.empty() ? 1 : v.size()) * sizeof(node)));
    if (p && !v.empty()) std::memcpy(p, v.data(), v.size() * sizeof(node));
    return p;
}

}  // namespace

extern "C" int fxg_pex_build(uint64_t total_query_length, uint64_t query_num_errors, uint64_t leaf_max_num_errors,
                             int build_strategy, fxg_pex_node** inner, size_t* n_inner, fxg_pex_node** leaves, size_t* n_leaves) {
    if (!inner || !n_inner || !leaves || !n_leaves || total_query_length == 0) return FXG_ERR_INVALID_ARGUMENT;
    if (build_strategy != 0 && build_strategy != 1) return FXG_ERR_INVALID_ARGUMENT;   // pex.cpp:100-101 throws
    try {
        std::vector<node> in, lv;
        if (build_strategy == 0) build_top_down(total_query_length, query_num_errors, leaf_max_num_errors, in, lv);
        else build_bottom_up(total_query_length, query_num_errors, leaf_max_num_errors, in, lv);
        *inner = to_c_array(in); *leaves = to_c_array(lv);
        if (!*inner || !*leaves) { std::free(*inner); std::free(*leaves); return FXG_ERR_OUT_OF_MEMORY; }
        *n_inner = in.size(); *n_leaves = lv.size();
        return FXG_OK;
    } catch (...) {
        return FXG_ERR_OUT_OF_MEMORY;
    }
}

extern "C" void fxg_pex_free(fxg_pex_node* nodes) { std::free(nodes); }
