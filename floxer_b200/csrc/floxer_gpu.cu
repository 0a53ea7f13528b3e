// floxer_gpu.cu -- C ABI (include/floxer_gpu.h) and host-side batch queue of the B200 PEX verification path.
//
// Host responsibilities (the work of parallelization.cpp:193-293 and verification.cpp:8-245 of the
// reference, re-organised for a GPU): turn align calls into DP passes, pick a (words-per-lane, ring size)
// configuration per pass, bucket passes by configuration and length, launch the engine, and drive the
// leaf -> root walk of every anchor level-synchronously ("waves").  No CPU fallback exists: every
// alignment result comes from the kernels in dp_kernels.cuh.
#include "../../include/floxer_gpu.h"
#include "dp_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

using namespace fxg;

// ------------------------------------------------------------------------------------------------ utilities

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t const want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { e = cudaMalloc(&p, bytes); if (e != cudaSuccess) { p = nullptr; return e; } cap = bytes; return e; }
        cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct RefStore {
    DevBuf packed;                       // 4 bits per base, every reference starts at a multiple of 32 bases
    std::vector<uint64_t> base, len;
    uint64_t total = 0;
};

// one staged pool of query bytes with its Peq planes (+ optional inline reference pool)
struct Pool {
    DevBuf bytes, peq, inline_packed;
    uint64_t len = 0, plane_words = 0, inline_len = 0;
    void release() { bytes.release(); peq.release(); inline_packed.release(); }
};

struct Pass {                            // one DP pass of the engine
    uint64_t ref_base;                   // position in the packed store / inline pool
    uint64_t query_base;                 // position in the pool
    uint32_t n, m;
    int32_t dlo, dhi;
    uint32_t flags;
};

struct TraceReq {                        // traceback request for an accepted CIGAR-mode alignment
    Pass pass;                           // sub-window pass (band around the end diagonal)
    uint32_t s_star;                     // score found by the score pass
    uint32_t col0;                       // sub-window start inside the original window
};

struct TraceOut { uint32_t begin_col; uint64_t cigar_offset; uint32_t cigar_len; };

constexpr int kWidths[6] = {1, 2, 4, 8, 16, 32};

}  // namespace

struct fxg_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;
    RefStore refs;
    bool have_refs = false;
    fxg_counters ctr{};
    int num_sms = 0;
    size_t smem_limit = 0;
    DevBuf d_tasks, d_results, d_trace, d_wtasks, d_wresults, d_cig_scratch, d_cig_pool, d_cursor, d_tmp;
    std::vector<uint32_t> h_cigar_pool;  // cigars of the last trace run
    std::mutex mu;
};

struct fxg_batch {
    std::vector<fxg_align_task> tasks;
    Pool pool;
    std::vector<fxg_align_result> results;
    std::vector<uint32_t> cigars;
    bool ran = false;
};

struct fxg_job {
    fxg_verify_config cfg{};
    std::vector<fxg_read> reads;
    std::vector<fxg_pex_node> nodes;
    std::vector<fxg_anchor> anchors;
    uint64_t pool_len = 0;
    Pool pool;                           // forward pool followed by the reverse-complement pool
    std::vector<fxg_alignment> alignments;
    std::vector<uint32_t> cigars;
    fxg_stats stats{};
    bool ran = false;
};

namespace {

int fail(fxg_ctx* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (c) c->err = buf;
    return code;
}

#define CUDA_TRY(ctx, expr)                                                                          \
    do {                                                                                             \
        cudaError_t e__ = (expr);                                                                    \
        if (e__ != cudaSuccess) return fail(ctx, FXG_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
    } while (0)

// ------------------------------------------------------------------------------------------------ kernel dispatch

template <int W, bool TR>
cudaError_t launch_one(DpLaunch const& L, uint32_t grid, size_t smem, cudaStream_t s) {
    dp_kernel<W, TR><<<grid, 32, smem, s>>>(L);
    return cudaGetLastError();
}

cudaError_t launch_dp(int widx, bool trace, DpLaunch const& L, uint32_t grid, size_t smem, cudaStream_t s) {
    switch (widx * 2 + (trace ? 1 : 0)) {
        case 0: return launch_one<1, false>(L, grid, smem, s);
        case 1: return launch_one<1, true>(L, grid, smem, s);
        case 2: return launch_one<2, false>(L, grid, smem, s);
        case 3: return launch_one<2, true>(L, grid, smem, s);
        case 4: return launch_one<4, false>(L, grid, smem, s);
        case 5: return launch_one<4, true>(L, grid, smem, s);
        case 6: return launch_one<8, false>(L, grid, smem, s);
        case 7: return launch_one<8, true>(L, grid, smem, s);
        case 8: return launch_one<16, false>(L, grid, smem, s);
        case 9: return launch_one<16, true>(L, grid, smem, s);
        case 10: return launch_one<32, false>(L, grid, smem, s);
        case 11: return launch_one<32, true>(L, grid, smem, s);
    }
    return cudaErrorInvalidValue;
}

template <int W, bool TR>
cudaError_t set_smem_attr(size_t bytes) {
    return cudaFuncSetAttribute(dp_kernel<W, TR>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes));
}

cudaError_t set_all_smem_attrs(size_t bytes) {
    cudaError_t e;
#define FXG_SET(W)                                                        \
    if ((e = set_smem_attr<W, false>(bytes)) != cudaSuccess) return e;    \
    if ((e = set_smem_attr<W, true>(bytes)) != cudaSuccess) return e;
    FXG_SET(1) FXG_SET(2) FXG_SET(4) FXG_SET(8) FXG_SET(16) FXG_SET(32)
#undef FXG_SET
    return cudaSuccess;
}

// ------------------------------------------------------------------------------------------------ configuration choice

struct Config { int widx; uint32_t G; uint64_t steps; uint32_t nb; };

inline uint32_t pow2_ceil(uint32_t x) { uint32_t p = 1; while (p < x) p <<= 1; return p; }
inline size_t win_stride_for(uint32_t n) { return (size_t(n + 63) / 32 + 1) * 32; }

// Picks words-per-lane W and ring size G for one pass.  The ring must be long enough that a lane is idle
// (and publishes the +1 boundary) whenever the block below still needs a boundary from it:
//   G > (B - 4) / (32 W + 1) + 2,  B = number of diagonals in the band  (derivation in DESIGN.md).
bool choose_config(Pass const& p, size_t smem_limit, Config& out) {
    uint32_t const nw = (p.m + 31) / 32;
    int64_t const B = int64_t(p.dhi) - int64_t(p.dlo) + 1;
    double best_cost = 1e300;
    bool found = false;
    for (int wi = 0; wi < 6; ++wi) {
        uint32_t const W = uint32_t(kWidths[wi]);
        uint32_t const nb = (nw + W - 1) / W;
        uint32_t G;
        if (nb == 1) G = 1;
        else {
            int64_t const need = (B > 4 ? (B - 4) / (32 * int64_t(W) + 1) : 0) + 3;
            int64_t const g = std::min<int64_t>(need, nb);
            if (g > 32) continue;
            G = pow2_ceil(uint32_t(std::max<int64_t>(g, 2)));
        }
        size_t const smem = size_t(32 / G) * (win_stride_for(p.n) + size_t(kNumSymbols) * nb * W * 4);
        if (smem > smem_limit) continue;
        uint64_t const steps = uint64_t(p.n) + nb - 1;
        double const cost = double(G) * double(steps) * (10.0 * W + 16.0);
        if (cost < best_cost) { best_cost = cost; out = Config{wi, G, steps, nb}; found = true; }
    }
    return found;
}

// ------------------------------------------------------------------------------------------------ running passes

struct Planned { uint32_t idx; Config cfg; };

// Runs `passes` (score passes, or trace passes when trace_bases != nullptr).  results[i] belongs to passes[i].
int run_passes(fxg_ctx* c, Pool const& pool, std::vector<Pass> const& passes, const uint64_t* trace_bases,
               std::vector<Config>* cfg_out, std::vector<DpResult>& results) {
    size_t const N = passes.size();
    results.assign(N, DpResult{kNoScore, 0});
    if (cfg_out) cfg_out->resize(N);
    if (N == 0) return FXG_OK;
    bool const trace = trace_bases != nullptr;

    std::vector<Planned> plan(N);
    for (size_t i = 0; i < N; ++i) {
        plan[i].idx = uint32_t(i);
        if (!choose_config(passes[i], c->smem_limit, plan[i].cfg))
            return fail(c, FXG_ERR_INVALID_ARGUMENT, "alignment of query length %u against window %u with band %d..%d exceeds the supported size",
                        passes[i].m, passes[i].n, passes[i].dlo, passes[i].dhi);
        if (cfg_out) (*cfg_out)[i] = plan[i].cfg;
    }
    std::sort(plan.begin(), plan.end(), [](Planned const& a, Planned const& b) {
        if (a.cfg.widx != b.cfg.widx) return a.cfg.widx > b.cfg.widx;
        if (a.cfg.G != b.cfg.G) return a.cfg.G > b.cfg.G;
        if (a.cfg.steps != b.cfg.steps) return a.cfg.steps > b.cfg.steps;
        return a.idx < b.idx;
    });

    std::vector<DpTask> tasks(N);
    for (size_t i = 0; i < N; ++i) {
        Pass const& p = passes[plan[i].idx];
        DpTask& t = tasks[i];
        t.ref_base = p.ref_base; t.query_base = p.query_base;
        t.trace_base = trace ? trace_bases[plan[i].idx] : 0;
        t.n = p.n; t.m = p.m; t.dlo = p.dlo; t.dhi = p.dhi; t.flags = p.flags; t.out = plan[i].idx;
    }
    CUDA_TRY(c, c->d_tasks.ensure(N * sizeof(DpTask)));
    CUDA_TRY(c, c->d_results.ensure(N * sizeof(DpResult)));
    CUDA_TRY(c, cudaMemcpyAsync(c->d_tasks.p, tasks.data(), N * sizeof(DpTask), cudaMemcpyHostToDevice, c->stream));
    c->ctr.h2d_bytes += N * sizeof(DpTask);
    CUDA_TRY(c, cudaMemsetAsync(c->d_results.p, 0x3f, N * sizeof(DpResult), c->stream));

    CUDA_TRY(c, cudaEventRecord(c->ev0, c->stream));
    size_t i = 0;
    while (i < N) {
        // one launch: same (W, G); window lengths within a factor of two so that shared memory is not wasted
        int const widx = plan[i].cfg.widx; uint32_t const G = plan[i].cfg.G;
        uint32_t const W = uint32_t(kWidths[widx]);
        size_t j = i;
        uint32_t max_n = 0, max_words = 0;
        uint32_t const first_n = passes[plan[i].idx].n;
        while (j < N && plan[j].cfg.widx == widx && plan[j].cfg.G == G) {
            uint32_t const n = passes[plan[j].idx].n;
            if (j > i && size_t(n) * 2 < first_n && (j - i) % (32 / G) == 0 && j - i >= 4096) break;
            max_n = std::max(max_n, n);
            max_words = std::max(max_words, plan[j].cfg.nb * W);
            ++j;
        }
        uint32_t const tpw = 32 / G;
        DpLaunch L{};
        L.tasks = c->d_tasks.as<DpTask>() + i;
        L.n_tasks = uint32_t(j - i);
        L.group = G;
        L.win_stride = uint32_t(win_stride_for(max_n));
        L.peq_stride = max_words;
        L.ref_packed = c->refs.packed.as<uint32_t>();
        L.inline_packed = pool.inline_packed.as<uint32_t>();
        L.peq_table = pool.peq.as<uint32_t>();
        L.peq_plane_words = pool.plane_words;
        L.results = c->d_results.as<DpResult>();
        L.trace = c->d_trace.as<uint32_t>();
        size_t const smem = size_t(tpw) * (L.win_stride + size_t(kNumSymbols) * L.peq_stride * 4);
        if (smem > c->smem_limit) return fail(c, FXG_ERR_INVALID_ARGUMENT, "internal: launch needs %zu bytes of shared memory", smem);
        uint32_t const grid = uint32_t((L.n_tasks + tpw - 1) / tpw);
        CUDA_TRY(c, launch_dp(widx, trace, L, grid, smem, c->stream));
        c->ctr.kernel_launches++;
        for (size_t q = i; q < j; ++q) {
            Pass const& p = passes[plan[q].idx];
            Config const& cf = plan[q].cfg;
            uint64_t ws = 0;                      // word-steps actually issued: every block is active for ce - cs + 1 columns
            uint32_t const rows = 32 * W;
            uint32_t const pad = cf.nb * rows - p.m;
            for (uint32_t b = 0; b < cf.nb; ++b) {
                int64_t lo = int64_t(rows) * b + 1 + p.dlo - pad, hi = int64_t(rows) * (b + 1) + p.dhi - pad;
                if (lo < 1) lo = 1;
                if (hi > p.n) hi = p.n;
                if (hi >= lo) ws += uint64_t(hi - lo + 1) * W;
            }
            c->ctr.dp_word_steps += ws;
            c->ctr.dp_cells_full += uint64_t(p.m) * p.n;
        }
        c->ctr.dp_tasks += j - i;
        i = j;
    }
    CUDA_TRY(c, cudaEventRecord(c->ev1, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(results.data(), c->d_results.p, N * sizeof(DpResult), cudaMemcpyDeviceToHost, c->stream));
    c->ctr.d2h_bytes += N * sizeof(DpResult);
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    float ms = 0;
    CUDA_TRY(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    if (trace) c->ctr.trace_kernel_ms += ms; else c->ctr.dp_kernel_ms += ms;
    return FXG_OK;
}

// Trace passes + walks for `reqs`; fills outs[i] and appends the cigars to c->h_cigar_pool (offsets are
// relative to the pool start at entry).  Works in chunks bounded by free device memory.
int run_traces(fxg_ctx* c, Pool const& pool, std::vector<TraceReq> const& reqs, std::vector<TraceOut>& outs) {
    size_t const N = reqs.size();
    outs.assign(N, TraceOut{0, 0, 0});
    if (N == 0) return FXG_OK;
    size_t free_b = 0, total_b = 0;
    CUDA_TRY(c, cudaMemGetInfo(&free_b, &total_b));
    uint64_t budget_words = std::max<uint64_t>((uint64_t(free_b) + c->d_trace.cap) / 2, uint64_t(64) << 20) / 4;
    budget_words = std::min<uint64_t>(budget_words, (uint64_t(48) << 30) / 4);

    std::vector<Config> cfgs(N);
    std::vector<uint64_t> words(N);
    for (size_t i = 0; i < N; ++i) {
        if (!choose_config(reqs[i].pass, c->smem_limit, cfgs[i]))
            return fail(c, FXG_ERR_INVALID_ARGUMENT, "traceback band of query length %u exceeds the supported size", reqs[i].pass.m);
        uint64_t const W = uint64_t(kWidths[cfgs[i].widx]);
        words[i] = ((uint64_t(reqs[i].pass.n) + cfgs[i].nb) * cfgs[i].G * W * 2 + 3) & ~uint64_t(3);
        if (words[i] > budget_words) budget_words = words[i];
    }
    size_t i = 0;
    while (i < N) {
        size_t j = i; uint64_t used = 0, cig_cap_total = 0;
        std::vector<Pass> passes; std::vector<uint64_t> tbase; std::vector<uint64_t> sbase;
        while (j < N && (j == i || used + words[j] <= budget_words)) {
            passes.push_back(reqs[j].pass); tbase.push_back(used); used += words[j];
            sbase.push_back(cig_cap_total); cig_cap_total += uint64_t(2) * reqs[j].s_star + 3;
            ++j;
        }
        size_t const M = j - i;
        CUDA_TRY(c, c->d_trace.ensure(used * 4));
        std::vector<DpResult> res; std::vector<Config> used_cfg;
        int rc = run_passes(c, pool, passes, tbase.data(), &used_cfg, res);
        if (rc != FXG_OK) return rc;
        c->ctr.trace_bytes += used * 4;
        // walks
        std::vector<WalkTask> wt(M);
        for (size_t q = 0; q < M; ++q) {
            TraceReq const& R = reqs[i + q];
            // the sub-window ends at the alignment's end column, whose value must be the score found before
            if (res[q].score != int32_t(R.s_star) || res[q].end_col != R.pass.n)
                return fail(c, FXG_ERR_CUDA, "internal: trace pass disagrees with score pass (%d@%u vs %u@%u)",
                            res[q].score, res[q].end_col, R.s_star, R.pass.n);
            WalkTask& w = wt[q];
            w.trace_base = tbase[q]; w.ref_base = R.pass.ref_base; w.query_base = R.pass.query_base;
            w.n = R.pass.n; w.m = R.pass.m; w.group = used_cfg[q].G; w.words = uint32_t(kWidths[used_cfg[q].widx]);
            w.flags = R.pass.flags; w.cigar_cap = 2 * R.s_star + 3; w.scratch_base = sbase[q]; w.out = uint32_t(q); w.reserved = 0;
        }
        CUDA_TRY(c, c->d_wtasks.ensure(M * sizeof(WalkTask)));
        CUDA_TRY(c, c->d_wresults.ensure(M * sizeof(WalkResult)));
        CUDA_TRY(c, c->d_cig_scratch.ensure(cig_cap_total * 4));
        CUDA_TRY(c, c->d_cig_pool.ensure(cig_cap_total * 4));
        CUDA_TRY(c, c->d_cursor.ensure(8));
        CUDA_TRY(c, cudaMemcpyAsync(c->d_wtasks.p, wt.data(), M * sizeof(WalkTask), cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(c, cudaMemsetAsync(c->d_cursor.p, 0, 8, c->stream));
        c->ctr.h2d_bytes += M * sizeof(WalkTask);
        WalkLaunch WL{};
        WL.tasks = c->d_wtasks.as<WalkTask>(); WL.n_tasks = uint32_t(M); WL.trace = c->d_trace.as<uint32_t>();
        WL.ref_packed = c->refs.packed.as<uint32_t>(); WL.inline_packed = pool.inline_packed.as<uint32_t>();
        WL.query_pool = pool.bytes.as<uint8_t>(); WL.scratch = c->d_cig_scratch.as<uint32_t>();
        WL.cigar_pool = c->d_cig_pool.as<uint32_t>(); WL.cigar_cursor = c->d_cursor.as<unsigned long long>();
        WL.cigar_pool_cap = cig_cap_total; WL.results = c->d_wresults.as<WalkResult>();
        CUDA_TRY(c, cudaEventRecord(c->ev0, c->stream));
        walk_kernel<<<uint32_t((M + 31) / 32), 32, 0, c->stream>>>(WL);
        CUDA_TRY(c, cudaGetLastError());
        CUDA_TRY(c, cudaEventRecord(c->ev1, c->stream));
        c->ctr.kernel_launches++;
        std::vector<WalkResult> wr(M);
        unsigned long long cursor = 0;
        CUDA_TRY(c, cudaMemcpyAsync(wr.data(), c->d_wresults.p, M * sizeof(WalkResult), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaMemcpyAsync(&cursor, c->d_cursor.p, 8, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        float ms = 0;
        CUDA_TRY(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        c->ctr.trace_kernel_ms += ms;
        size_t const pool_at = c->h_cigar_pool.size();
        c->h_cigar_pool.resize(pool_at + cursor);
        if (cursor) CUDA_TRY(c, cudaMemcpy(c->h_cigar_pool.data() + pool_at, c->d_cig_pool.p, cursor * 4, cudaMemcpyDeviceToHost));
        c->ctr.d2h_bytes += M * sizeof(WalkResult) + cursor * 4;
        for (size_t q = 0; q < M; ++q) {
            if (wr[q].cigar_len == 0xffffffffu) return fail(c, FXG_ERR_CUDA, "internal: traceback overflowed its CIGAR scratch");
            outs[i + q] = TraceOut{wr[q].begin_col, pool_at + wr[q].cigar_offset, wr[q].cigar_len};
        }
        i = j;
    }
    return FXG_OK;
}

// ------------------------------------------------------------------------------------------------ pools

int check_ranks(fxg_ctx* c, const uint8_t* p, size_t n, const char* what) {
    uint8_t worst = 0;
    for (size_t i = 0; i < n; ++i) worst = p[i] > worst ? p[i] : worst;
    if (worst > FXG_MAX_RANK) return fail(c, FXG_ERR_INVALID_ARGUMENT, "%s contains rank %u (allowed 0..%d)", what, unsigned(worst), FXG_MAX_RANK);
    return FXG_OK;
}

int upload_packed(fxg_ctx* c, const uint8_t* ranks, uint64_t len, DevBuf& dst, uint64_t word_offset) {
    // pack on the device: upload bytes to a temporary, 8 bases per output word
    if (len == 0) return FXG_OK;
    CUDA_TRY(c, c->d_tmp.ensure(len));
    CUDA_TRY(c, cudaMemcpyAsync(c->d_tmp.p, ranks, len, cudaMemcpyHostToDevice, c->stream));
    c->ctr.h2d_bytes += len;
    uint64_t const n_words = (len + 7) / 8;
    uint32_t const grid = uint32_t(std::min<uint64_t>((n_words + 255) / 256, uint64_t(c->num_sms) * 16));
    pack_nibbles_kernel<<<grid, 256, 0, c->stream>>>(c->d_tmp.as<uint8_t>(), len, dst.as<uint32_t>() + word_offset);
    CUDA_TRY(c, cudaGetLastError());
    c->ctr.kernel_launches++;
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return FXG_OK;
}

// uploads up to two byte ranges back to back (forward / reverse pools) and builds the Peq planes
int stage_pool(fxg_ctx* c, Pool& pool, const uint8_t* a, size_t a_len, const uint8_t* b, size_t b_len) {
    pool.len = a_len + b_len;
    pool.plane_words = (pool.len + 31) / 32 + kPeqFrontPadWords + kPeqBackPadWords;
    CUDA_TRY(c, pool.bytes.ensure(pool.len + 64));
    CUDA_TRY(c, pool.peq.ensure(pool.plane_words * kNumSymbols * 4));
    if (a_len) CUDA_TRY(c, cudaMemcpyAsync(pool.bytes.p, a, a_len, cudaMemcpyHostToDevice, c->stream));
    if (b_len) CUDA_TRY(c, cudaMemcpyAsync(pool.bytes.as<uint8_t>() + a_len, b, b_len, cudaMemcpyHostToDevice, c->stream));
    c->ctr.h2d_bytes += pool.len;
    CUDA_TRY(c, cudaMemsetAsync(pool.peq.p, 0, pool.plane_words * kNumSymbols * 4, c->stream));
    if (pool.len) {
        uint64_t const n_words = (pool.len + 31) / 32;
        uint32_t const grid = uint32_t(std::min<uint64_t>((n_words + 7) / 8, uint64_t(c->num_sms) * 16));
        build_peq_kernel<<<grid, 256, 0, c->stream>>>(pool.bytes.as<uint8_t>(), pool.len, pool.peq.as<uint32_t>(), pool.plane_words);
        CUDA_TRY(c, cudaGetLastError());
        c->ctr.kernel_launches++;
    }
    return FXG_OK;
}

// ------------------------------------------------------------------------------------------------ align semantics

// band of a score pass: every alignment with <= k errors of a query of length m inside a window of length n
// stays on diagonals j - i in [-k, n - m + k]
inline bool score_pass_for(uint64_t ref_base, uint64_t query_base, uint32_t n, uint32_t m, uint32_t k, uint32_t flags, Pass& p) {
    if (m == 0) return false;
    if (int64_t(m) - int64_t(n) > int64_t(k)) return false;          // more insertions needed than errors allowed
    p.ref_base = ref_base; p.query_base = query_base; p.n = n; p.m = m;
    int64_t const lo = -int64_t(k), hi = int64_t(n) - int64_t(m) + int64_t(k);
    p.dlo = int32_t(lo); p.dhi = int32_t(hi); p.flags = flags;
    return true;
}

// sub-window + band of the trace pass for an alignment with score s ending in column end_col (1-based, exclusive end)
inline TraceReq trace_req_for(Pass const& score_pass, uint32_t s, uint32_t end_col) {
    TraceReq r;
    int64_t const c0 = std::max<int64_t>(0, int64_t(end_col) - int64_t(score_pass.m) - int64_t(s) - 1);
    r.col0 = uint32_t(c0);
    r.s_star = s;
    r.pass = score_pass;
    r.pass.ref_base = score_pass.ref_base + uint64_t(c0);
    r.pass.n = end_col - uint32_t(c0);
    int64_t const d_end = int64_t(r.pass.n) - int64_t(score_pass.m);
    r.pass.dlo = int32_t(d_end - int64_t(s) - 1);
    r.pass.dhi = int32_t(d_end + int64_t(s) + 1);
    r.pass.flags = score_pass.flags & ~kFlagReverse;
    return r;
}

}  // namespace

// ================================================================================================ C ABI

extern "C" {

const char* fxg_version(void) { return "floxer_b200 0.1 (sm_100a)"; }

int fxg_create(int device, fxg_ctx** out) {
    if (!out) return FXG_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0 || device < 0 || device >= count) return FXG_ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return FXG_ERR_CUDA;
    fxg_ctx* c = new (std::nothrow) fxg_ctx();
    if (!c) return FXG_ERR_OUT_OF_MEMORY;
    c->device = device;
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete c; return FXG_ERR_CUDA; }
    c->num_sms = prop.multiProcessorCount;
    c->smem_limit = prop.sharedMemPerBlockOptin;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess ||
        set_all_smem_attrs(c->smem_limit) != cudaSuccess) {
        delete c; return FXG_ERR_CUDA;
    }
    *out = c;
    return FXG_OK;
}

void fxg_destroy(fxg_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    c->refs.packed.release();
    for (DevBuf* b : {&c->d_tasks, &c->d_results, &c->d_trace, &c->d_wtasks, &c->d_wresults, &c->d_cig_scratch, &c->d_cig_pool, &c->d_cursor, &c->d_tmp}) b->release();
    cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1);
    cudaStreamDestroy(c->stream);
    delete c;
}

const char* fxg_last_error(const fxg_ctx* c) { return c ? c->err.c_str() : "no context"; }

int fxg_get_counters(const fxg_ctx* c, fxg_counters* out) { if (!c || !out) return FXG_ERR_INVALID_ARGUMENT; *out = c->ctr; return FXG_OK; }
int fxg_reset_counters(fxg_ctx* c) { if (!c) return FXG_ERR_INVALID_ARGUMENT; c->ctr = fxg_counters{}; return FXG_OK; }

int fxg_set_references(fxg_ctx* c, size_t n_refs, const uint8_t* const* ranks, const uint64_t* lens) {
    if (!c || (n_refs && (!ranks || !lens))) return FXG_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lock(c->mu);
    CUDA_TRY(c, cudaSetDevice(c->device));
    RefStore& R = c->refs;
    R.base.assign(n_refs, 0); R.len.assign(lens, lens + n_refs);
    uint64_t total = 0;
    for (size_t i = 0; i < n_refs; ++i) {
        int rc = check_ranks(c, ranks[i], lens[i], "reference");
        if (rc != FXG_OK) return rc;
        R.base[i] = total;
        total += (lens[i] + 31) / 32 * 32;
    }
    R.total = total;
    CUDA_TRY(c, R.packed.ensure(total / 2 + 64));
    CUDA_TRY(c, cudaMemsetAsync(R.packed.p, 0, total / 2 + 64, c->stream));
    for (size_t i = 0; i < n_refs; ++i) {
        // upload in slices so that the temporary stays small
        uint64_t const slice = uint64_t(256) << 20;
        for (uint64_t at = 0; at < lens[i]; at += slice) {
            uint64_t const n = std::min<uint64_t>(slice, lens[i] - at);
            int rc = upload_packed(c, ranks[i] + at, n, R.packed, (R.base[i] + at) / 8);
            if (rc != FXG_OK) return rc;
        }
    }
    c->d_tmp.release();
    c->have_refs = true;
    return FXG_OK;
}

// ------------------------------------------------------------------------------------------------ align batch

int fxg_align_batch_stage(fxg_ctx* c, const fxg_align_task* tasks, size_t n_tasks, const uint8_t* query_pool, size_t query_pool_len,
                          const uint8_t* inline_ref_pool, size_t inline_ref_pool_len, fxg_batch** out) {
    if (!c || !out || (n_tasks && !tasks)) return FXG_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lock(c->mu);
    *out = nullptr;
    CUDA_TRY(c, cudaSetDevice(c->device));
    for (size_t i = 0; i < n_tasks; ++i) {
        fxg_align_task const& t = tasks[i];
        if (t.mode > FXG_MODE_CIGAR) return fail(c, FXG_ERR_INVALID_ARGUMENT, "task %zu: invalid alignment mode %u", i, unsigned(t.mode));
        if (t.query_offset + t.query_len > query_pool_len) return fail(c, FXG_ERR_INVALID_ARGUMENT, "task %zu: query span outside the pool", i);
        if (t.ref_id == FXG_REF_INLINE) {
            if (t.ref_offset + t.ref_len > inline_ref_pool_len) return fail(c, FXG_ERR_INVALID_ARGUMENT, "task %zu: inline reference span outside the pool", i);
        } else {
            if (!c->have_refs || t.ref_id >= c->refs.len.size()) return fail(c, FXG_ERR_STATE, "task %zu: reference %u not resident", i, t.ref_id);
            if (t.ref_offset + t.ref_len > c->refs.len[t.ref_id]) return fail(c, FXG_ERR_INVALID_ARGUMENT, "task %zu: reference span outside reference %u", i, t.ref_id);
        }
    }
    int rc = check_ranks(c, query_pool, query_pool_len, "query pool");
    if (rc != FXG_OK) return rc;
    rc = check_ranks(c, inline_ref_pool, inline_ref_pool_len, "inline reference pool");
    if (rc != FXG_OK) return rc;
    fxg_batch* b = new (std::nothrow) fxg_batch();
    if (!b) return FXG_ERR_OUT_OF_MEMORY;
    b->tasks.assign(tasks, tasks + n_tasks);
    rc = stage_pool(c, b->pool, query_pool, query_pool_len, nullptr, 0);
    if (rc == FXG_OK && inline_ref_pool_len) {
        b->pool.inline_len = inline_ref_pool_len;
        cudaError_t e = b->pool.inline_packed.ensure(inline_ref_pool_len / 2 + 64);
        if (e != cudaSuccess) rc = fail(c, FXG_ERR_CUDA, "inline pool allocation: %s", cudaGetErrorString(e));
        else { cudaMemsetAsync(b->pool.inline_packed.p, 0, inline_ref_pool_len / 2 + 64, c->stream); rc = upload_packed(c, inline_ref_pool, inline_ref_pool_len, b->pool.inline_packed, 0); }
    }
    if (rc == FXG_OK && cudaStreamSynchronize(c->stream) != cudaSuccess) rc = fail(c, FXG_ERR_CUDA, "staging failed");
    if (rc != FXG_OK) { b->pool.release(); delete b; return rc; }
    *out = b;
    return FXG_OK;
}

int fxg_align_batch_run(fxg_ctx* c, fxg_batch* b) {
    if (!c || !b) return FXG_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lock(c->mu);
    CUDA_TRY(c, cudaSetDevice(c->device));
    size_t const N = b->tasks.size();
    b->results.assign(N, fxg_align_result{});
    b->cigars.clear();
    std::vector<Pass> passes; passes.reserve(N);
    std::vector<uint32_t> owner; owner.reserve(N);
    for (size_t i = 0; i < N; ++i) {
        fxg_align_task const& t = b->tasks[i];
        b->results[i].orientation = t.orientation;
        uint32_t flags = (t.mode == FXG_MODE_NO_CIGAR ? kFlagReverse : 0u) | (t.ref_id == FXG_REF_INLINE ? kFlagInlineRef : 0u);
        uint64_t const ref_base = (t.ref_id == FXG_REF_INLINE ? 0 : c->refs.base[t.ref_id]) + t.ref_offset;
        Pass p;
        if (t.query_len == 0) {
            // empty query: aligns with zero errors; the rightmost minimum of an all-zero last row is column n
            b->results[i].exists = 1;
            b->results[i].start_in_reference = t.reference_span_offset + (t.mode == FXG_MODE_CIGAR ? t.ref_len : 0);
            if (t.mode == FXG_MODE_EXISTS) b->results[i].start_in_reference = 0;
            continue;
        }
        if (score_pass_for(ref_base, t.query_offset, t.ref_len, t.query_len, t.max_errors, flags, p)) { passes.push_back(p); owner.push_back(uint32_t(i)); }
    }
    std::vector<DpResult> res;
    int rc = run_passes(c, b->pool, passes, nullptr, nullptr, res);
    if (rc != FXG_OK) return rc;
    std::vector<TraceReq> reqs; std::vector<uint32_t> req_owner;
    for (size_t q = 0; q < passes.size(); ++q) {
        fxg_align_task const& t = b->tasks[owner[q]];
        fxg_align_result& r = b->results[owner[q]];
        if (res[q].score > int32_t(t.max_errors)) continue;
        r.exists = 1;
        if (t.mode == FXG_MODE_EXISTS) continue;
        r.num_errors = uint32_t(res[q].score);
        if (t.mode == FXG_MODE_NO_CIGAR) r.start_in_reference = t.reference_span_offset + (t.ref_len - res[q].end_col);   // alignment.cpp:135-139
        else { reqs.push_back(trace_req_for(passes[q], uint32_t(res[q].score), res[q].end_col)); req_owner.push_back(owner[q]); }
    }
    c->h_cigar_pool.clear();
    std::vector<TraceOut> touts;
    rc = run_traces(c, b->pool, reqs, touts);
    if (rc != FXG_OK) return rc;
    b->cigars.swap(c->h_cigar_pool);
    for (size_t q = 0; q < reqs.size(); ++q) {
        fxg_align_task const& t = b->tasks[req_owner[q]];
        fxg_align_result& r = b->results[req_owner[q]];
        r.start_in_reference = t.reference_span_offset + reqs[q].col0 + touts[q].begin_col;   // alignment.cpp:175
        r.cigar_offset = touts[q].cigar_offset; r.cigar_len = touts[q].cigar_len;
    }
    b->ran = true;
    return FXG_OK;
}

int fxg_align_batch_fetch(fxg_ctx* c, fxg_batch* b, fxg_align_result* results, uint32_t* cigar_pool, size_t cigar_capacity, size_t* cigar_used) {
    if (!c || !b || !b->ran || (b->tasks.size() && !results)) return c ? fail(c, FXG_ERR_STATE, "batch has not been run") : FXG_ERR_INVALID_ARGUMENT;
    if (cigar_used) *cigar_used = b->cigars.size();
    if (b->cigars.size() > cigar_capacity) return fail(c, FXG_ERR_OVERFLOW, "cigar pool needs %zu entries, capacity is %zu", b->cigars.size(), cigar_capacity);
    std::memcpy(results, b->results.data(), b->results.size() * sizeof(fxg_align_result));
    if (!b->cigars.empty()) std::memcpy(cigar_pool, b->cigars.data(), b->cigars.size() * 4);
    return FXG_OK;
}

void fxg_batch_free(fxg_ctx* c, fxg_batch* b) {
    if (!b) return;
    if (c) { std::lock_guard<std::mutex> lock(c->mu); cudaSetDevice(c->device); b->pool.release(); }
    delete b;
}

int fxg_align_batch(fxg_ctx* c, const fxg_align_task* tasks, size_t n_tasks, const uint8_t* query_pool, size_t query_pool_len,
                    const uint8_t* inline_ref_pool, size_t inline_ref_pool_len, fxg_align_result* results,
                    uint32_t* cigar_pool, size_t cigar_capacity, size_t* cigar_used) {
    fxg_batch* b = nullptr;
    int rc = fxg_align_batch_stage(c, tasks, n_tasks, query_pool, query_pool_len, inline_ref_pool, inline_ref_pool_len, &b);
    if (rc != FXG_OK) return rc;
    rc = fxg_align_batch_run(c, b);
    if (rc == FXG_OK) rc = fxg_align_batch_fetch(c, b, results, cigar_pool, cigar_capacity, cigar_used);
    fxg_batch_free(c, b);
    return rc;
}

// ------------------------------------------------------------------------------------------------ verify

int fxg_verify_stage(fxg_ctx* c, const fxg_verify_config* cfg, const fxg_read* reads, size_t n_reads,
                     const uint8_t* fwd, const uint8_t* rc_pool, size_t pool_len,
                     const fxg_pex_node* nodes, size_t n_nodes, const fxg_anchor* anchors, size_t n_anchors, fxg_job** out) {
    if (!c || !cfg || !out) return FXG_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lock(c->mu);
    *out = nullptr;
    if (!c->have_refs) return fail(c, FXG_ERR_STATE, "fxg_set_references must be called first");
    if (cfg->verification_kind != FXG_KIND_DIRECT_FULL && cfg->verification_kind != FXG_KIND_HIERARCHICAL)
        return fail(c, FXG_ERR_INVALID_ARGUMENT, "Internal bug in verification kind (should not happen)");   // verification.cpp:19
    CUDA_TRY(c, cudaSetDevice(c->device));
    for (size_t i = 0; i < n_reads; ++i) {
        fxg_read const& R = reads[i];
        if (R.query_offset + R.query_len > pool_len) return fail(c, FXG_ERR_INVALID_ARGUMENT, "read %zu: query outside the pool", i);
        if (R.query_len > FXG_MAX_QUERY_LENGTH) return fail(c, FXG_ERR_INVALID_ARGUMENT, "read %zu: longer than %d", i, FXG_MAX_QUERY_LENGTH);
        if (R.node_offset + R.num_inner + R.num_leaves > n_nodes || R.num_leaves == 0) return fail(c, FXG_ERR_INVALID_ARGUMENT, "read %zu: bad node range", i);
        if (R.anchor_offset + R.num_anchors_forward + R.num_anchors_reverse > n_anchors) return fail(c, FXG_ERR_INVALID_ARGUMENT, "read %zu: bad anchor range", i);
        const fxg_pex_node* nd = nodes + R.node_offset;
        for (uint32_t q = 0; q < R.num_inner + R.num_leaves; ++q) {
            if (nd[q].query_index_to < nd[q].query_index_from || nd[q].query_index_to >= R.query_len) return fail(c, FXG_ERR_INVALID_ARGUMENT, "read %zu: node %u outside the query", i, q);
            if (nd[q].parent_id != FXG_NULL_ID && nd[q].parent_id >= R.num_inner) return fail(c, FXG_ERR_INVALID_ARGUMENT, "read %zu: node %u has a bad parent", i, q);
        }
        const fxg_anchor* an = anchors + R.anchor_offset;
        for (uint32_t q = 0; q < R.num_anchors_forward + R.num_anchors_reverse; ++q) {
            if (an[q].pex_leaf_index >= R.num_leaves) return fail(c, FXG_ERR_INVALID_ARGUMENT, "read %zu: anchor %u names leaf %llu", i, q, (unsigned long long)an[q].pex_leaf_index);
            if (an[q].reference_id >= c->refs.len.size()) return fail(c, FXG_ERR_INVALID_ARGUMENT, "read %zu: anchor %u names reference %llu", i, q, (unsigned long long)an[q].reference_id);
            if (an[q].reference_position >= c->refs.len[an[q].reference_id]) return fail(c, FXG_ERR_INVALID_ARGUMENT, "read %zu: anchor %u outside its reference", i, q);
        }
    }
    int rc = check_ranks(c, fwd, pool_len, "forward pool");
    if (rc == FXG_OK) rc = check_ranks(c, rc_pool, pool_len, "reverse-complement pool");
    if (rc != FXG_OK) return rc;
    fxg_job* j = new (std::nothrow) fxg_job();
    if (!j) return FXG_ERR_OUT_OF_MEMORY;
    j->cfg = *cfg;
    j->reads.assign(reads, reads + n_reads);
    j->nodes.assign(nodes, nodes + n_nodes);
    j->anchors.assign(anchors, anchors + n_anchors);
    j->pool_len = pool_len;
    rc = stage_pool(c, j->pool, fwd, pool_len, rc_pool, pool_len);
    if (rc == FXG_OK && cudaStreamSynchronize(c->stream) != cudaSuccess) rc = fail(c, FXG_ERR_CUDA, "staging failed");
    if (rc != FXG_OK) { j->pool.release(); delete j; return rc; }
    *out = j;
    return FXG_OK;
}

namespace {

// math::floating_point_error_aware_ceil, include/math.hpp:22-27
inline uint64_t ceil_eps(double v) { return uint64_t(std::ceil(v - 0.000000001) + 0.000000001); }

struct Span { uint64_t offset, length, extra; };

// verification::internal::compute_reference_span_start_and_length, src/lib/verification.cpp:157-184
inline Span compute_span(uint64_t anchor_pos, fxg_pex_node const& node, uint64_t leaf_from, uint64_t ref_len, double ratio) {
    uint64_t const base = (node.query_index_to - node.query_index_from + 1) + 2 * node.num_errors + 1;
    uint64_t const extra = ceil_eps(double(base) * ratio);
    int64_t const s = int64_t(anchor_pos) - int64_t(leaf_from - node.query_index_from) - int64_t(node.num_errors) - int64_t(extra);
    Span r;
    r.offset = s >= 0 ? uint64_t(s) : 0;
    r.length = std::min<uint64_t>(base + 2 * extra, ref_len - r.offset);
    r.extra = extra;
    return r;
}

// half_open_interval::trim_from_both_sides, src/lib/intervals.cpp:48-58
inline void trim(uint64_t& start, uint64_t& end, uint64_t amount) {
    uint64_t const e = amount > end ? 0 : end - amount;
    uint64_t const ne = std::max(start + 1, e);
    uint64_t const ns = std::min(ne - 1, start + amount);
    start = ns; end = ne;
}

enum : uint8_t { W_WAITING = 0, W_WALKING = 1, W_DONE = 2 };

struct Walk {                 // one query_verifier::verify() call
    uint32_t read, anchor;    // anchor = index into the job's anchor array
    uint8_t orient, state;
    bool reached_root;        // the root window was inserted into verified_intervals
    bool hit;
    const fxg_pex_node* node; // node to align next
    uint64_t r_start, r_end;  // root window
    uint64_t t_start, t_end;  // root window trimmed by the extra length (root_was_already_verified)
    Span root_span;
    // result of an accepted root
    uint64_t start_in_reference; uint32_t num_errors; uint64_t cigar_offset; uint32_t cigar_len;
};

}  // namespace

int fxg_verify_run(fxg_ctx* c, fxg_job* J) {
    if (!c || !J) return FXG_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lock(c->mu);
    CUDA_TRY(c, cudaSetDevice(c->device));
    J->alignments.clear(); J->cigars.clear(); J->stats = fxg_stats{};
    bool const ivopt = J->cfg.interval_optimization != 0;
    bool const direct = J->cfg.verification_kind == FXG_KIND_DIRECT_FULL;
    double const ratio = J->cfg.extra_verification_ratio;

    // ---- one Walk per anchor, in the reference's order: read, forward package, reverse package ----
    std::vector<Walk> walks; walks.reserve(J->anchors.size());
    // groups: anchors of one (read, orientation, reference) share a verified_intervals set (parallelization.cpp:224-226)
    struct Group { std::vector<uint32_t> members; size_t first_open = 0; std::vector<uint32_t> inserted; };
    std::vector<Group> groups;
    std::vector<uint32_t> group_of;
    for (size_t ri = 0; ri < J->reads.size(); ++ri) {
        fxg_read const& R = J->reads[ri];
        const fxg_pex_node* inner = J->nodes.data() + R.node_offset;
        const fxg_pex_node* leaves = inner + R.num_inner;
        fxg_pex_node const& root = R.num_inner ? inner[0] : leaves[0];
        for (int orient = 0; orient < 2; ++orient) {
            uint32_t const a0 = uint32_t(R.anchor_offset) + (orient ? R.num_anchors_forward : 0);
            uint32_t const na = orient ? R.num_anchors_reverse : R.num_anchors_forward;
            size_t const group_base = groups.size();
            std::vector<int64_t> ref_to_group;
            for (uint32_t q = 0; q < na; ++q) {
                fxg_anchor const& A = J->anchors[a0 + q];
                fxg_pex_node const& leaf = leaves[A.pex_leaf_index];
                Walk w{};
                w.read = uint32_t(ri); w.anchor = a0 + q; w.orient = uint8_t(orient); w.state = W_WAITING;
                w.root_span = compute_span(A.reference_position, root, leaf.query_index_from, c->refs.len[A.reference_id], ratio);
                w.r_start = w.root_span.offset; w.r_end = w.root_span.offset + w.root_span.length;
                w.t_start = w.r_start; w.t_end = w.r_end;
                trim(w.t_start, w.t_end, w.root_span.extra);
                w.node = (direct || leaf.parent_id == FXG_NULL_ID) ? &root : &inner[leaf.parent_id];
                if (ref_to_group.size() <= A.reference_id) ref_to_group.resize(A.reference_id + 1, -1);
                if (ref_to_group[A.reference_id] < 0) { ref_to_group[A.reference_id] = int64_t(groups.size()); groups.emplace_back(); }
                groups[size_t(ref_to_group[A.reference_id])].members.push_back(uint32_t(walks.size()));
                group_of.push_back(uint32_t(ref_to_group[A.reference_id]));
                walks.push_back(w);
            }
            (void)group_base;
        }
    }

    std::vector<uint32_t> active;           // walks that need a DP pass in this wave
    std::vector<Pass> passes; std::vector<uint32_t> pass_walk; std::vector<uint8_t> pass_root;
    std::vector<TraceReq> reqs; std::vector<uint32_t> req_walk;
    std::vector<DpResult> res;
    size_t n_done = 0;

    auto span_of = [&](Walk const& w) -> Span {
        if (w.node->parent_id == FXG_NULL_ID) return w.root_span;
        fxg_read const& R = J->reads[w.read];
        const fxg_pex_node* leaves = J->nodes.data() + R.node_offset + R.num_inner;
        fxg_anchor const& A = J->anchors[w.anchor];
        return compute_span(A.reference_position, *w.node, leaves[A.pex_leaf_index].query_index_from, c->refs.len[A.reference_id], 0.0);
    };

    while (n_done < walks.size()) {
        // ---- admission: a waiting walk starts once no earlier unresolved walk of its group could still verify its root window ----
        for (Group& g : groups) {
            while (g.first_open < g.members.size() && walks[g.members[g.first_open]].state == W_DONE) ++g.first_open;
            for (size_t q = g.first_open; q < g.members.size(); ++q) {
                Walk& w = walks[g.members[q]];
                if (w.state != W_WAITING) continue;
                bool skip = false, blocked = false;
                if (ivopt) {
                    // verified_intervals::contains over the windows inserted by EARLIER anchors (intervals.cpp:94-127)
                    for (uint32_t ins : g.inserted) {
                        if (ins < g.members[q] && walks[ins].r_start <= w.t_start && walks[ins].r_end >= w.t_end) { skip = true; break; }
                    }
                    if (!skip) {
                        for (size_t e = g.first_open; e < q; ++e) {
                            Walk const& u = walks[g.members[e]];
                            if (u.state != W_DONE && u.r_start <= w.t_start && u.r_end >= w.t_end) { blocked = true; break; }
                        }
                    }
                }
                if (skip) {                                      // root_was_already_verified, verification.cpp:119-136
                    J->stats.n_avoided_root++; J->stats.sum_avoided_root += w.root_span.length;
                    w.state = W_DONE; ++n_done;
                } else if (!blocked) {
                    w.state = W_WALKING; active.push_back(g.members[q]);
                }
            }
        }
        if (active.empty()) {
            if (n_done < walks.size()) return fail(c, FXG_ERR_STATE, "internal: verification scheduler stalled");
            break;
        }
        // ---- one DP pass per active walk ----
        passes.clear(); pass_walk.clear(); pass_root.clear();
        std::vector<uint32_t> next_active;
        std::vector<std::pair<uint32_t, bool>> no_pass;     // walks whose align call needs no DP (impossible by length)
        for (uint32_t wi : active) {
            Walk& w = walks[wi];
            fxg_anchor const& A = J->anchors[w.anchor];
            bool const is_root = w.node->parent_id == FXG_NULL_ID;
            Span const sp = span_of(w);
            uint32_t const m = uint32_t(w.node->query_index_to - w.node->query_index_from + 1);
            uint32_t flags = (is_root && J->cfg.without_cigar) ? kFlagReverse : 0u;
            uint64_t const qbase = (w.orient ? J->pool_len : 0) + J->reads[w.read].query_offset + w.node->query_index_from;
            // statistics, verification.cpp:238-242
            if (is_root) { J->stats.n_aligned_root++; J->stats.sum_aligned_root += sp.length; J->stats.cells_root += uint64_t(m) * sp.length; }
            else { J->stats.n_aligned_inner++; J->stats.sum_aligned_inner += sp.length; J->stats.cells_inner += uint64_t(m) * sp.length; }
            Pass p;
            if (score_pass_for(c->refs.base[A.reference_id] + sp.offset, qbase, uint32_t(sp.length), m, uint32_t(w.node->num_errors), flags, p)) {
                passes.push_back(p); pass_walk.push_back(wi); pass_root.push_back(uint8_t(is_root));
            } else {
                no_pass.emplace_back(wi, is_root);
            }
        }
        int rc = run_passes(c, J->pool, passes, nullptr, nullptr, res);
        if (rc != FXG_OK) return rc;
        c->ctr.waves++;
        auto finish = [&](uint32_t wi, bool is_root, bool exists, DpResult const* r, Pass const* p) {
            Walk& w = walks[wi];
            if (is_root) {
                w.reached_root = true;                          // verified_intervals.insert, verification.cpp:106-109 / :40-41
                if (ivopt) groups[group_of[wi]].inserted.push_back(wi);
                if (exists) {
                    w.hit = true; w.num_errors = uint32_t(r->score);
                    Span const sp = w.root_span;
                    if (J->cfg.without_cigar) w.start_in_reference = sp.offset + (sp.length - r->end_col);
                    else { reqs.push_back(trace_req_for(*p, uint32_t(r->score), r->end_col)); req_walk.push_back(wi); }
                }
                w.state = W_DONE; ++n_done;
            } else if (exists) {
                fxg_read const& R = J->reads[w.read];
                w.node = &J->nodes[R.node_offset + w.node->parent_id];      // pex_tree::get_parent_of_child, pex.cpp:70-76
                next_active.push_back(wi);
            } else {
                w.state = W_DONE; ++n_done;
            }
        };
        for (size_t q = 0; q < passes.size(); ++q) {
            Walk const& w = walks[pass_walk[q]];
            bool const exists = res[q].score <= int32_t(w.node->num_errors);
            finish(pass_walk[q], pass_root[q] != 0, exists, &res[q], &passes[q]);
        }
        for (auto const& np : no_pass) finish(np.first, np.second, false, nullptr, nullptr);
        active.swap(next_active);
    }

    // ---- tracebacks for accepted roots ----
    c->h_cigar_pool.clear();
    std::vector<TraceOut> touts;
    int rc = run_traces(c, J->pool, reqs, touts);
    if (rc != FXG_OK) return rc;
    J->cigars.swap(c->h_cigar_pool);
    for (size_t q = 0; q < reqs.size(); ++q) {
        Walk& w = walks[req_walk[q]];
        w.start_in_reference = w.root_span.offset + reqs[q].col0 + touts[q].begin_col;
        w.cigar_offset = touts[q].cigar_offset; w.cigar_len = touts[q].cigar_len;
    }
    // ---- emit in anchor order (= insertion order of the reference's single-thread run) ----
    for (Walk const& w : walks) {
        if (!w.hit) continue;
        fxg_alignment a{};
        a.start_in_reference = w.start_in_reference; a.cigar_offset = w.cigar_offset; a.cigar_len = w.cigar_len;
        a.num_errors = w.num_errors; a.read_index = w.read; a.reference_id = uint32_t(J->anchors[w.anchor].reference_id);
        a.orientation = w.orient;
        J->alignments.push_back(a);
    }
    J->ran = true;
    return FXG_OK;
}

size_t fxg_job_num_alignments(const fxg_job* j) { return j ? j->alignments.size() : 0; }
const fxg_alignment* fxg_job_alignments(const fxg_job* j) { return j ? j->alignments.data() : nullptr; }
size_t fxg_job_cigar_len(const fxg_job* j) { return j ? j->cigars.size() : 0; }
const uint32_t* fxg_job_cigar_pool(const fxg_job* j) { return j ? j->cigars.data() : nullptr; }
const fxg_stats* fxg_job_stats(const fxg_job* j) { return j ? &j->stats : nullptr; }

void fxg_job_free(fxg_ctx* c, fxg_job* j) {
    if (!j) return;
    if (c) { std::lock_guard<std::mutex> lock(c->mu); cudaSetDevice(c->device); j->pool.release(); }
    delete j;
}

int fxg_verify_reads(fxg_ctx* c, const fxg_verify_config* cfg, const fxg_read* reads, size_t n_reads,
                     const uint8_t* fwd, const uint8_t* rc_pool, size_t pool_len,
                     const fxg_pex_node* nodes, size_t n_nodes, const fxg_anchor* anchors, size_t n_anchors, fxg_job** out) {
    fxg_job* j = nullptr;
    int rc = fxg_verify_stage(c, cfg, reads, n_reads, fwd, rc_pool, pool_len, nodes, n_nodes, anchors, n_anchors, &j);
    if (rc != FXG_OK) return rc;
    rc = fxg_verify_run(c, j);
    if (rc != FXG_OK) { fxg_job_free(c, j); return rc; }
    *out = j;
    return FXG_OK;
}

// ------------------------------------------------------------------------------------------------ int32 peak

int fxg_measure_int32_peak(fxg_ctx* c, double* out) {
    if (!c || !out) return FXG_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lock(c->mu);
    CUDA_TRY(c, cudaSetDevice(c->device));
    CUDA_TRY(c, c->d_tmp.ensure(64));
    uint32_t const iters = 4096, threads = 256, grid = uint32_t(c->num_sms) * 8;
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        CUDA_TRY(c, cudaEventRecord(c->ev0, c->stream));
        int32_peak_kernel<<<grid, threads, 0, c->stream>>>(c->d_tmp.as<uint32_t>(), iters, 12345u + rep);
        CUDA_TRY(c, cudaGetLastError());
        CUDA_TRY(c, cudaEventRecord(c->ev1, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        float ms = 0;
        CUDA_TRY(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        double const instr = double(grid) * threads * double(iters) * 8.0 * 11.0;
        if (rep > 0) best = std::max(best, instr / (double(ms) * 1e-3));
    }
    *out = best;
    return FXG_OK;
}

}  // extern "C"
