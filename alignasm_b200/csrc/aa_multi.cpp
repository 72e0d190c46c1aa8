// aa_multi.cpp — contig sharding across the GPUs of one box (host side only, no CUDA in this file).
// The reference parallelises over contigs with tbb::parallel_for (src/alignasm.cpp:351-359); contigs are
// independent, so here they are partitioned by a cost estimate (longest processing time first), every shard is
// solved by its own aa_ctx on its own device from its own host thread, and the per-contig row lists are merged
// back in input order.  No collective, no NCCL: the only inter-GPU "traffic" is the host scattering the
// structure-of-arrays shards and gathering a few hundred bytes of rows per contig.
#include "../../include/alignasm_b200.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <condition_variable>
#include <mutex>
#include <cstdlib>
#include <cstring>
#include <string>
#include <atomic>
#include <thread>
#include <vector>

namespace {
thread_local std::string g_multi_err;

// Contexts are kept between calls (a context owns its pooled workspace: a warm one makes no cudaMalloc).  This is the one
// piece of process-wide state of the library, and it is guarded: a call CHECKS OUT one context per shard for the whole
// shard solve (a context is one bump pool, one pinned staging buffer and one pair of streams: never shared), a device whose
// cached contexts are all busy gets a fresh one, and aa_multi_release waits for running solves.
struct CtxSlot {
    int32_t device;
    aa_ctx *ctx;
    bool busy;
};
std::mutex g_ctx_mutex;
std::condition_variable g_ctx_cv;
std::vector<CtxSlot> g_ctx_cache;

aa_status checkout_ctx(int32_t device, aa_ctx **out, std::string &err) {
    {
        std::lock_guard<std::mutex> lock(g_ctx_mutex);
        for (auto &s : g_ctx_cache)
            if (s.device == device && !s.busy) {
                s.busy = true;
                *out = s.ctx;
                return AA_OK;
            }
    }
    aa_ctx *ctx = nullptr;  // created outside the lock: cudaMalloc / stream creation take milliseconds
    aa_status st = aa_create(&ctx, device);
    if (st != AA_OK) {
        err = aa_last_error(nullptr);
        return st;
    }
    std::lock_guard<std::mutex> lock(g_ctx_mutex);
    g_ctx_cache.push_back({device, ctx, true});
    *out = ctx;
    return AA_OK;
}
void checkin_ctx(aa_ctx *ctx) {
    {
        std::lock_guard<std::mutex> lock(g_ctx_mutex);
        for (auto &s : g_ctx_cache)
            if (s.ctx == ctx) s.busy = false;
    }
    g_ctx_cv.notify_all();
}

struct Shard {
    std::vector<int64_t> ctgs;     // input contig ids, ascending
    std::vector<int64_t> ctg_off;  // block offsets of those contigs inside the shard
    aa_result res{};
    aa_status st = AA_OK;
    std::string err;
};

template <class T>
T *host_n(int64_t n) {
    return (T *)std::calloc((size_t)(n > 0 ? n : 1), sizeof(T));
}
template <class T>
T *host_raw(int64_t n) {  // every element is written by the merge
    return (T *)std::malloc((size_t)(n > 0 ? n : 1) * sizeof(T));
}
void rows_alloc(aa_rows &r, int64_t n) {
    r.n = n;
    r.ctg_index = host_raw<int32_t>(n);
    r.qry_str = host_raw<int64_t>(n);
    r.qry_end = host_raw<int64_t>(n);
    r.ref_str = host_raw<int64_t>(n);
    r.ref_end = host_raw<int64_t>(n);
    r.is_alt = host_raw<uint8_t>(n);
}
void rows_copy(aa_rows &dst, int64_t at, const aa_rows &src, int64_t from, int64_t n) {
    if (n <= 0) return;
    std::memcpy(dst.ctg_index + at, src.ctg_index + from, (size_t)n * 4);
    std::memcpy(dst.qry_str + at, src.qry_str + from, (size_t)n * 8);
    std::memcpy(dst.qry_end + at, src.qry_end + from, (size_t)n * 8);
    std::memcpy(dst.ref_str + at, src.ref_str + from, (size_t)n * 8);
    std::memcpy(dst.ref_end + at, src.ref_end + from, (size_t)n * 8);
    std::memcpy(dst.is_alt + at, src.is_alt + from, (size_t)n);
}
}  // namespace

extern "C" {

const char *aa_multi_last_error(void) { return g_multi_err.c_str(); }

/* release the contexts aa_solve_multi keeps between calls (waits for solves that are still running on them) */
void aa_multi_release(void) {
    std::unique_lock<std::mutex> lock(g_ctx_mutex);
    g_ctx_cv.wait(lock, [] {
        for (auto &s : g_ctx_cache)
            if (s.busy) return false;
        return true;
    });
    for (auto &s : g_ctx_cache) aa_destroy(s.ctx);
    g_ctx_cache.clear();
}

/* cost model of one contig = the SM time it takes (one warp per contig in the serial phases): the K-walk enumeration is
   ~1.3 us per walk whatever the contig's size, the heap / relax / walk-0 chains ~0.6 us per block (C2 on one B200, round 2:
   13 ms of enumeration for every contig above a few thousand blocks, 22 ms of heap inserts for the 43 099-block contig).
   Largest first also puts the longest serial chains, which bound a shard from below, on different devices. */
void aa_shard_contigs(const aa_batch *b, int32_t max_walks, int32_t n_shards, int32_t *shard_of) {
    const int64_t C = b->n_ctg;
    const double K = max_walks > 0 ? std::min<double>(max_walks, 10000) : 10000.0;
    std::vector<double> cost((size_t)C);
    std::vector<int64_t> order((size_t)C);
    for (int64_t c = 0; c < C; c++) {
        const double n = (double)(b->ctg_off[c + 1] - b->ctg_off[c]);
        cost[(size_t)c] = (n > 1 ? 1.3 * K : 0.0) + 0.6 * n;
        order[(size_t)c] = c;
    }
    std::stable_sort(order.begin(), order.end(), [&](int64_t x, int64_t y) { return cost[(size_t)x] > cost[(size_t)y]; });
    std::vector<double> load((size_t)std::max(1, n_shards), 0.0);
    for (int64_t c : order) {
        const size_t k = (size_t)(std::min_element(load.begin(), load.end()) - load.begin());
        shard_of[c] = (int32_t)k;
        load[k] += cost[(size_t)c];
    }
}

aa_status aa_solve_multi(const int32_t *devices, int32_t n_dev, const aa_batch *b, const aa_opts *opts, aa_result *res) {
    if (!devices || n_dev <= 0 || !b || !res || b->n_ctg <= 0) {
        g_multi_err = "aa_solve_multi: bad arguments";
        return AA_ERR_INVALID;
    }
    const int64_t C = b->n_ctg;
    aa_opts o{};
    if (opts) o = *opts;
    if (o.keep_debug) {
        g_multi_err = "aa_solve_multi: keep_debug is only available on a single device";
        return AA_ERR_INVALID;
    }
    const bool trace = std::getenv("AA_MULTI_TRACE") != nullptr;  // stderr: where the wall time of the call goes
    const auto t_begin = std::chrono::steady_clock::now();
    auto ms_since = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(); };
    std::vector<int32_t> shard_of((size_t)C);
    aa_shard_contigs(b, o.max_walks, n_dev, shard_of.data());
    std::vector<Shard> sh((size_t)n_dev);
    for (int64_t c = 0; c < C; c++) sh[(size_t)shard_of[(size_t)c]].ctgs.push_back(c);
    std::vector<std::thread> pool;
    for (int32_t k = 0; k < n_dev; k++)
        pool.emplace_back([&, k]() {
            Shard &s = sh[(size_t)k];
            if (s.ctgs.empty()) return;
            s.ctg_off.assign(1, 0);
            for (int64_t c : s.ctgs) s.ctg_off.push_back(s.ctg_off.back() + (b->ctg_off[c + 1] - b->ctg_off[c]));
            aa_ctx *ctx = nullptr;
            s.st = checkout_ctx(devices[k], &ctx, s.err);  // exclusively ours until the shard is solved
            if (s.st != AA_OK) return;
            const double t0 = ms_since();
            s.st = aa_solve_subset(ctx, b, s.ctgs.data(), (int64_t)s.ctgs.size(), &o, &s.res);  // staged from the caller's arrays
            if (s.st != AA_OK) s.err = aa_last_error(ctx);
            if (trace)
                std::fprintf(stderr, "[aa_multi] shard %d on device %d: %zu contigs, solve call %.1f .. %.1f ms (device %.1f ms)\n", k, devices[k],
                             s.ctgs.size(), t0, ms_since(), s.res.stats.ms_total);
            checkin_ctx(ctx);
        });
    for (auto &t : pool) t.join();
    aa_status st = AA_OK;
    for (auto &s : sh)
        if (s.st != AA_OK && s.st != AA_ERR_UNSOLVABLE) {
            st = s.st;
            g_multi_err = s.err;
        }
    if (st == AA_OK)
        for (auto &s : sh)
            if (s.st == AA_ERR_UNSOLVABLE) {
                st = s.st;
                g_multi_err = s.err;
            }
    if (st != AA_OK && st != AA_ERR_UNSOLVABLE) {
        for (auto &s : sh) aa_result_free(&s.res);
        return st;
    }
    if (trace) std::fprintf(stderr, "[aa_multi] shards done at %.1f ms\n", ms_since());
    // ---- merge in input contig order ----
    std::memset(res, 0, sizeof *res);
    res->n_ctg = C;
    res->out_off = host_n<int64_t>(C + 1);
    res->alt_off = host_n<int64_t>(C + 1);
    res->all_path_off = host_n<int64_t>(C + 1);
    res->sorted_index = host_raw<int32_t>(b->n_blk);
    std::vector<int64_t> local((size_t)C);  // position of every contig inside its shard
    for (auto &s : sh)
        for (size_t i = 0; i < s.ctgs.size(); i++) local[(size_t)s.ctgs[i]] = (int64_t)i;
    int64_t n_out = 0, n_alt = 0, n_paths = 0, n_all = 0;
    for (int64_t c = 0; c < C; c++) {
        const aa_result &r = sh[(size_t)shard_of[(size_t)c]].res;
        const int64_t l = local[(size_t)c];
        n_out += r.out_off[l + 1] - r.out_off[l];
        n_alt += r.alt_off[l + 1] - r.alt_off[l];
        const int64_t p0 = r.all_path_off[l], p1 = r.all_path_off[l + 1];
        n_paths += p1 - p0;
        n_all += r.all_row_off[p1] - r.all_row_off[p0];
        res->out_off[c + 1] = n_out;
        res->alt_off[c + 1] = n_alt;
        res->all_path_off[c + 1] = n_paths;
    }
    res->all_row_off = host_n<int64_t>(n_paths + 1);
    rows_alloc(res->out, n_out);
    rows_alloc(res->alt, n_alt);
    rows_alloc(res->all, n_all);
    // where the .all paths / rows of every contig start (sequential, cheap), then the copies on several host threads
    std::vector<int64_t> all_at((size_t)C + 1, 0);
    for (int64_t c = 0; c < C; c++) {
        const aa_result &r = sh[(size_t)shard_of[(size_t)c]].res;
        const int64_t l = local[(size_t)c], p0 = r.all_path_off[l], p1 = r.all_path_off[l + 1];
        all_at[(size_t)c + 1] = all_at[(size_t)c] + (r.all_row_off[p1] - r.all_row_off[p0]);
    }
    {
        const int nt = (int)std::max<int64_t>(1, std::min<int64_t>({(int64_t)std::thread::hardware_concurrency(), 16, C}));
        std::atomic<int64_t> next{0};
        auto work = [&]() {
            const int64_t CH = 16;
            for (;;) {
                const int64_t c0 = next.fetch_add(CH);
                if (c0 >= C) break;
                for (int64_t c = c0; c < std::min(C, c0 + CH); c++) {
                    const Shard &s = sh[(size_t)shard_of[(size_t)c]];
                    const aa_result &r = s.res;
                    const int64_t l = local[(size_t)c];
                    rows_copy(res->out, res->out_off[c], r.out, r.out_off[l], r.out_off[l + 1] - r.out_off[l]);
                    rows_copy(res->alt, res->alt_off[c], r.alt, r.alt_off[l], r.alt_off[l + 1] - r.alt_off[l]);
                    int64_t at_all = all_at[(size_t)c], at_path = res->all_path_off[c];
                    for (int64_t p = r.all_path_off[l]; p < r.all_path_off[l + 1]; p++) {
                        const int64_t n = r.all_row_off[p + 1] - r.all_row_off[p];
                        rows_copy(res->all, at_all, r.all, r.all_row_off[p], n);
                        at_all += n;
                        res->all_row_off[++at_path] = at_all;
                    }
                    const int64_t nb = b->ctg_off[c + 1] - b->ctg_off[c];
                    std::memcpy(res->sorted_index + b->ctg_off[c], r.sorted_index + s.ctg_off[(size_t)l], (size_t)nb * 4);
                }
            }
        };
        std::vector<std::thread> mpool;
        for (int t = 1; t < nt; t++) mpool.emplace_back(work);
        work();
        for (auto &t : mpool) t.join();
    }
    // statistics: sizes add up, times are the slowest shard's
    aa_stats &t = res->stats;
    for (auto &s : sh) {
        if (s.ctgs.empty()) continue;
        const aa_stats &x = s.res.stats;
        t.n_ctg += x.n_ctg;
        t.n_blk += x.n_blk;
        t.n_run += x.n_run;
        t.n_pair += x.n_pair;
        t.n_vtx += x.n_vtx;
        t.n_edge += x.n_edge;
        t.n_heap += x.n_heap;
        t.n_walk += x.n_walk;
        t.n_task += x.n_task;
        t.n_launch += x.n_launch;
        t.algo_bytes += x.algo_bytes;
        if (x.ms_total > t.ms_total) {
            t.ms_total = x.ms_total;
            std::memcpy(t.ms_phase, x.ms_phase, sizeof t.ms_phase);
        }
        for (int p = 0; p < 16; p++) t.algo_bytes_phase[p] += x.algo_bytes_phase[p];
    }
    if (trace) std::fprintf(stderr, "[aa_multi] merged at %.1f ms\n", ms_since());
    for (auto &s : sh) aa_result_free(&s.res);
    if (trace) std::fprintf(stderr, "[aa_multi] shard results freed at %.1f ms\n", ms_since());
    return st;
}

}  // extern "C"
