// aa_pipeline.cuh — phase orchestration of the hot path, written once against a Backend policy.
//
// Backend = where the items of a phase run.  The product instantiates it with CudaBackend
// (aa_backend_cuda.cuh: kernels on sm_100a, CUB scans/sorts, a pooled HBM allocator).  tests/emul
// instantiates it with a host-loop backend to check the same phase functions on the CPU against the
// reference; that instantiation is never part of libalignasm_b200.so.
//
// Phases (names from aa_phase_name, one CUDA-event pair each):
//   0 sort      host std::sort of (qry_str, qry_end) per contig — OC1: the reference's unstable
//               std::sort must see the same comparison outcomes (paf_data.cpp:241), then gather on device
//   1 parts     paf_data.cpp:249-261
//   2 pairs     candidate pairs, cut points, compaction into pair vertices (paf_data.cpp:294-378)
//   3 edges     out-degree count, scan, fill in adjacency order (paf_data.cpp:531-696)
//   4 reverse   stable radix sort of edges by destination -> reverse CSR (k_shortest_walks.hpp:180-183)
//   5 relax     reverse Kahn + first-wins relaxation + min-anom DP
//   6 topo      forward Kahn order (paf_data.cpp:742-746)
//   7 heaps     persistent leftist sidetrack heaps (k_shortest_walks.hpp:191-215)
//   8 enum      K-walk enumeration (k_shortest_walks.hpp:217-251)
//   9 plan      which walks the selection recovers, in call order (paf_data.cpp:1585-1649)
//  10 walksA    recover + upgrade every planned walk: coverage, row count, block marks
//  11 select    primary / alt / .all winners
//  12 walksB    rows of the winners
//  13 d2h       result download
//
// Streams (device backend): the phases run on the main stream.  Three things run beside it: the forward order (topo)
// and, after the relax, the main chain of walk 0 (trace, speculate, resolve, rows) on the side stream — joined before
// walksA; the heaps of the small contigs on the aux stream; and, on batches of many contigs, the small contigs' whole
// heaps -> enumeration pipeline on the aux stream while the large contigs' heap chains are still being built (two-group
// pipelining).  The CUDA-event pair of a phase is recorded on the main stream: with pipelining the
// small group's enumeration is inside the `heaps` phase time and `enum` is the large group's.
#pragma once
#include "../../include/alignasm_b200.h"
#include "aa_core.cuh"

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace aa {

enum Phase { PH_SORT = 0, PH_PARTS, PH_PAIRS, PH_EDGES, PH_REVERSE, PH_RELAX, PH_TOPO, PH_HEAPS, PH_ENUM, PH_PLAN,
             PH_WALKS_A, PH_SELECT, PH_WALKS_B, PH_D2H, PH_COUNT };
inline const char *phase_name(int p) {
    static const char *names[PH_COUNT] = {"sort", "parts", "pairs", "edges", "reverse", "relax", "topo", "heaps",
                                          "enum", "plan", "walksA", "select", "walksB", "d2h"};
    return (p >= 0 && p < PH_COUNT) ? names[p] : nullptr;
}

// ---- phase functors: one call per item ----------------------------------------------------------------
#define AA_FUNCTOR(Name, body)                                  \
    struct Name {                                               \
        Ws w;                                                   \
        AA_HD void operator()(int64_t i, void *scratch) const {  \
            (void)scratch;                                      \
            body;                                               \
        }        \
    };
AA_FUNCTOR(FnGather, f_gather(w, i))
AA_FUNCTOR(FnCandCount, f_cand_count(w, i))
AA_FUNCTOR(FnCuts, f_cuts(w, i))
AA_FUNCTOR(FnVtxOff, f_vtx_off(w, i))
AA_FUNCTOR(FnDegree, f_degree(w, i))
AA_FUNCTOR(FnFill, f_fill(w, i))
AA_FUNCTOR(FnHeapPrep, f_heap_prep(w, i))
AA_FUNCTOR(FnTourBuild, f_tour_build(w, i))
AA_FUNCTOR(FnKlKeys, f_kl_keys(w, i))
AA_FUNCTOR(FnKlPull, f_kl_pull(w, i))
AA_FUNCTOR(FnKlFinish, f_kl_finish(w, i))
template <bool REV>
struct FnKlInit {
    Ws w;
    AA_HD void operator()(int64_t i, void *) const { f_kl_init<REV>(w, i); }
};
template <bool REV>
struct FnKlAssign {
    Ws w;
    int64_t n;
    AA_HD void operator()(int64_t i, void *) const { f_kl_assign<REV>(w, i, n); }
};
template <bool REV>
struct FnKlExpand {
    Ws w;
    AA_HD void operator()(int64_t i, void *) const { f_kl_expand<REV>(w, i); }
};
struct FnKlCount {
    Ws w;
    int64_t n;
    AA_HD void operator()(int64_t i, void *) const { f_kl_count(w, i, n); }
};
struct FnCtgEdges {
    Ws w;
    int64_t *out;
    AA_HD void operator()(int64_t i, void *) const { f_ctg_edges(w, i, out); }
};
AA_FUNCTOR(FnOpsClass, f_ops_class(w, i))
AA_FUNCTOR(FnOwnerJump, (f_owner_jump(w, i), f_owner_jump(w, i)))
AA_FUNCTOR(FnOpsFill, f_ops_fill(w, i))
AA_FUNCTOR(FnChainSlot, f_chain_slot(w, i))
AA_FUNCTOR(FnRootFill, f_root_fill(w, i))
AA_FUNCTOR(FnLeafFlag, f_leaf_flag(w, i))
AA_FUNCTOR(FnLeafList, f_leaf_list(w, i))
AA_FUNCTOR(FnNodeRank, f_node_rank(w, i))
struct FnLvlOff {
    Ws w;
    const int32_t *ctgs;
    int32_t per;
    int64_t *out;
    AA_HD void operator()(int64_t i, void *) const { f_lvl_off(w, i, ctgs, per, out); }
};
#if defined(__CUDACC__)
struct FnHeapsLevel {  // one warp per tree vertex of one depth (device only)
    Ws w;
    int64_t first;
    __device__ void operator()(int64_t i, void *) const { f_heaps_level(w, first + i); }
};
#endif
struct FnHeapsLeaf {  // one warp per tree leaf with inserts
    Ws w;
    AA_HD void operator()(int64_t i, void *) const { f_heaps_leaf_any(w, (int64_t)w.leaf_list[i]); }
};
AA_FUNCTOR(FnBfsPos, f_bfs_pos(w, i))
AA_FUNCTOR(FnVInfo, f_vinfo(w, i))
AA_FUNCTOR(FnInsFill, f_ins_fill(w, i))
struct FnTourJump {
    const Tour *src;
    Tour *dst;
    AA_HD void operator()(int64_t i, void *) const { tour_jump(src, dst, i); }
};
struct FnBfsKey {
    Ws w;
    const Tour *tour;
    AA_HD void operator()(int64_t i, void *) const { f_bfs_key(w, i, tour); }
};
AA_FUNCTOR(FnRevPack, f_rev_pack(w, i))
AA_FUNCTOR(FnENext, f_enext(w, i))
AA_FUNCTOR(FnXRec, f_xrec(w, i))
AA_FUNCTOR(FnRelaxInit, f_relax_init(w, i))
AA_FUNCTOR(FnRelaxUnpack, f_relax_unpack(w, i))
AA_FUNCTOR(FnMainSpec, f_main_spec(w, i))
AA_FUNCTOR(FnMainRows, f_main_rows(w, i))
AA_FUNCTOR(FnTasksA1, f_tasks_a1(w, i))
AA_FUNCTOR(FnTasksA1Solo, f_tasks_a1_solo(w, i))
struct FnMainTotals {
    Ws w;
    AA_HD void operator()(int64_t i, void *) const { f_main_totals(w, i); }
};
struct FnTasksA0 {
    Ws w;
    const int32_t *ord;
    AA_HD void operator()(int64_t i, void *) const { f_tasks_a0(w, i, ord); }
};
struct FnCompact {
    Ws w;
    int64_t ncand;
    AA_HD void operator()(int64_t i, void *) const { f_compact(w, i, ncand); }
};
struct FnRevOff {
    Ws w;
    int64_t E;
    AA_HD void operator()(int64_t i, void *) const { f_rev_off(w, i, E); }
};
struct FnTasksB {
    Ws w;
    int64_t n_items, n_paths;
    AA_HD void operator()(int64_t i, void *) const { f_tasks_b(w, i, n_items, n_paths); }
};
// per-contig phases go through an order array (largest contigs first)
#define AA_CTG_FUNCTOR(Name, body)                                  \
    struct Name {                                                   \
        Ws w;                                                       \
        const int32_t *ord;                                         \
        AA_HD void operator()(int64_t i, void *scratch) const {     \
            (void)scratch;                                          \
            int64_t c = ord[i];                                     \
            body;                                                   \
        }                                                           \
    };
AA_CTG_FUNCTOR(FnParts, f_parts_any(w, c))
struct FnPartsBnd {
    Ws w;
    const int64_t *pm;
    int32_t *bl, *br;
    AA_HD void operator()(int64_t i, void *) const { f_parts_bnd(w, i, pm, bl, br); }
};
struct FnPartsFin {
    Ws w;
    const int32_t *br;
    AA_HD void operator()(int64_t i, void *) const { f_parts_fin(w, i, br); }
};
AA_CTG_FUNCTOR(FnRelax, f_relax_any(w, c, scratch))
AA_CTG_FUNCTOR(FnRelaxRedo, f_relax_redo_any(w, c, scratch))
AA_CTG_FUNCTOR(FnRelaxSweep, f_relax_sweep_any(w, c, scratch))
struct FnSegBounds {  // one bucket of SEG_BLOCKS sorted blocks
    Ws w;
    AA_HD void operator()(int64_t i, void *) const { f_seg_bounds_any(w, i); }
};
struct FnTopoSeg {
    Ws w;
    AA_HD void operator()(int64_t i, void *scratch) const { f_topo_seg_any(w, i, scratch); }
};
AA_CTG_FUNCTOR(FnTopoRedo, f_topo_redo_any(w, c, scratch))
struct FnRelaxSeg {
    Ws w;
    AA_HD void operator()(int64_t i, void *scratch) const { f_relax_seg_any(w, i, scratch); }
};
AA_CTG_FUNCTOR(FnTopo, f_topo_any(w, c, scratch))
AA_CTG_FUNCTOR(FnHeaps, f_heaps_any(w, c, scratch))
AA_CTG_FUNCTOR(FnEnum, f_enum_any(w, c, scratch))
AA_CTG_FUNCTOR(FnPlan, f_plan_any(w, c))
AA_CTG_FUNCTOR(FnTaskCompact, f_task_compact(w, c))
AA_CTG_FUNCTOR(FnAllList, f_all_list(w, c))
AA_CTG_FUNCTOR(FnMainTrace, f_main_trace(w, c))
struct FnSelect {
    Ws w;
    const int32_t *ord;
    int32_t want_all;
    AA_HD void operator()(int64_t i, void *) const { f_select(w, ord[i], want_all); }
};

// ---- a batch staged on the device ------------------------------------------------------------------------
struct DevBatch {
    int64_t C = 0, B = 0, R = 0;
    // device copies of the aa_batch arrays
    int64_t *ctg_off = nullptr, *qs = nullptr, *qe = nullptr, *rs = nullptr, *re = nullptr, *qtot = nullptr;
    int32_t *chr = nullptr;
    uint8_t *fwd = nullptr, *mapq = nullptr;
    int64_t *run_off = nullptr, *run_ql = nullptr, *run_qr = nullptr, *run_rl = nullptr;
    // host mirror of what the host-side sort needs (OC1)
    std::vector<int64_t> own_ctg_off, own_qs, own_qe;  // copies, for a batch that stays resident (or a staged subset)
    std::vector<int64_t> own_run_off;                  // rebased run offsets of a staged subset
    const int64_t *h_ctg_off = nullptr, *h_qs = nullptr, *h_qe = nullptr;  // the caller's arrays for a one-shot solve
    std::vector<void *> owned;
    bool pooled = false;  // arrays live in the solve workspace (one-shot aa_solve): nothing to free
    // a resident batch is sorted once, when it is uploaded (the order depends on the batch alone); a one-shot solve sorts inside
    bool sorted = false;
    std::vector<int32_t> perm, sorted_index, ctg_order;
    int32_t *d_perm = nullptr, *d_ord = nullptr;
};

inline void rows_alloc_host(aa_rows &r, int64_t n) {
    r.n = n;
    size_t m = (size_t)(n > 0 ? n : 1);
    r.ctg_index = (int32_t *)std::malloc(m * 4);
    r.qry_str = (int64_t *)std::malloc(m * 8);
    r.qry_end = (int64_t *)std::malloc(m * 8);
    r.ref_str = (int64_t *)std::malloc(m * 8);
    r.ref_end = (int64_t *)std::malloc(m * 8);
    r.is_alt = (uint8_t *)std::malloc(m);
}
inline void rows_free_host(aa_rows &r) {
    std::free(r.ctg_index);
    std::free(r.qry_str);
    std::free(r.qry_end);
    std::free(r.ref_str);
    std::free(r.ref_end);
    std::free(r.is_alt);
    std::memset(&r, 0, sizeof r);
}
template <class T>
inline T *host_n(int64_t n) {
    return (T *)std::calloc((size_t)(n > 0 ? n : 1), sizeof(T));
}
// A device backend may hand out the arrays of a result as ONE slab of pinned host memory (the rows are then copied from the
// device straight into the caller's arrays, no staging copy and no page faults); out_off is the slab's base.  The hook gives
// the slab back and says whether `base` was one.
inline bool (*g_result_slab_release)(void *base) = nullptr;
inline void result_free_host(aa_result *res) {
    if (!res) return;
    if (res->out_off && g_result_slab_release && g_result_slab_release(res->out_off)) {
        // every array below lived in the slab
    } else {
        std::free(res->out_off);
        std::free(res->alt_off);
        std::free(res->all_path_off);
        std::free(res->all_row_off);
        std::free(res->sorted_index);
        rows_free_host(res->out);
        rows_free_host(res->alt);
        rows_free_host(res->all);
    }
    if (aa_debug *g = res->dbg) {
        void *ptrs[] = {g->vtx_off, g->edge_off, g->walk_off, g->e_src, g->e_dst, g->e_qry, g->e_ref, g->e_anom, g->e_qnz,
                        g->e_qtot, g->d_reach, g->d_sum, g->d_anom, g->d_qnz, g->d_qtot, g->best, g->order, g->w_sum,
                        g->w_anom, g->w_qnz, g->w_qtot, g->anom_dis};
        for (void *p : ptrs) std::free(p);
        std::free(g);
    }
    std::memset(res, 0, sizeof *res);
}

// Validate offsets; returns an error text or empty
inline std::string validate_batch(const aa_batch *b) {
    if (!b) return "null batch";
    if (b->n_ctg <= 0 || b->n_blk <= 0) return "empty batch";
    if (!b->ctg_off || !b->qry_str || !b->qry_end || !b->ref_str || !b->ref_end || !b->qry_total || !b->ref_chr ||
        !b->aln_fwd || !b->map_qul || !b->run_off)
        return "null array in batch";
    if (b->ctg_off[0] != 0 || b->ctg_off[b->n_ctg] != b->n_blk) return "ctg_off does not span the blocks";
    for (int64_t c = 0; c < b->n_ctg; c++)
        if (b->ctg_off[c + 1] <= b->ctg_off[c]) return "empty contig (the reference asserts at least one block)";
    if (b->run_off[0] != 0 || b->run_off[b->n_blk] != b->n_run) return "run_off does not span the runs";
    for (int64_t i = 0; i < b->n_blk; i++)
        if (b->run_off[i + 1] < b->run_off[i]) return "run_off not monotone";
    if (b->n_run > 0 && (!b->run_ql || !b->run_qr || !b->run_rl)) return "null run array";
    return "";
}

// Host std::sort per contig with the reference comparator (paf_data.hpp:69-73).  perm[b0+i] = original
// in-contig position of the block that lands at sorted position i.
inline void host_sort_perm(const DevBatch &d, std::vector<int32_t> &perm, std::vector<int32_t> &sorted_index,
                           std::vector<int32_t> &ctg_order, int threads) {
    const int64_t C = d.C;
    perm.resize((size_t)d.B);
    sorted_index.resize((size_t)d.B);
    struct Key {
        int64_t qs, qe;
        int32_t idx;
        bool operator<(const Key &r) const { return qs != r.qs ? qs < r.qs : qe < r.qe; }
    };
    std::atomic<int64_t> next{0};
    auto worker = [&]() {
        std::vector<Key> keys;
        for (;;) {
            int64_t c0 = next.fetch_add(16);
            if (c0 >= C) break;
            int64_t c1 = std::min<int64_t>(C, c0 + 16);
            for (int64_t c = c0; c < c1; c++) {
                int64_t b0 = d.h_ctg_off[(size_t)c], n = d.h_ctg_off[(size_t)c + 1] - b0;
                keys.resize((size_t)n);
                for (int64_t i = 0; i < n; i++) keys[(size_t)i] = {d.h_qs[(size_t)(b0 + i)], d.h_qe[(size_t)(b0 + i)], (int32_t)i};
                if (n > 1) std::sort(keys.begin(), keys.end());
                for (int64_t i = 0; i < n; i++) {
                    perm[(size_t)(b0 + i)] = keys[(size_t)i].idx;
                    sorted_index[(size_t)(b0 + keys[(size_t)i].idx)] = (int32_t)i;
                }
            }
        }
    };
    if (threads <= 1 || C < 64) {
        worker();
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; t++) pool.emplace_back(worker);
        for (auto &t : pool) t.join();
    }
    // contig processing order for the per-contig phases: largest first
    ctg_order.resize((size_t)C);
    for (int64_t c = 0; c < C; c++) ctg_order[(size_t)c] = (int32_t)c;
    std::stable_sort(ctg_order.begin(), ctg_order.end(), [&](int32_t a, int32_t b) {
        return d.h_ctg_off[(size_t)a + 1] - d.h_ctg_off[(size_t)a] > d.h_ctg_off[(size_t)b + 1] - d.h_ctg_off[(size_t)b];
    });
}

template <class BK>
struct Pipeline {
    BK &bk;
    std::string err;
    explicit Pipeline(BK &b) : bk(b) {}

    template <class T>
    T *A(int64_t n) {
        return (T *)bk.alloc_bytes((size_t)(n > 0 ? n : 1) * sizeof(T));
    }

    // pooled: stage the batch in the workspace pool (which the caller has just reset) instead of cudaMalloc'ing it;
    // the solve that follows must then keep the pool (solve(..., keep_pool = true))
    aa_status upload(const aa_batch *b, DevBatch *&out, bool pooled = false) {
        std::string v = validate_batch(b);
        if (!v.empty()) {
            err = v;
            return AA_ERR_INVALID;
        }
        DevBatch *d = new DevBatch();
        d->C = b->n_ctg;
        d->B = b->n_blk;
        d->R = b->n_run;
        d->pooled = pooled;
        auto up = [&](auto *&dst, const auto *src, int64_t n) {
            using T = std::remove_const_t<std::remove_pointer_t<decltype(src)>>;
            const size_t bytes = (size_t)(n > 0 ? n : 1) * sizeof(T);
            dst = (T *)(pooled ? bk.alloc_bytes(bytes) : bk.alloc_persistent(bytes));
            if (!dst) return false;
            if (!pooled) d->owned.push_back(dst);
            if (n > 0) {
                if (pooled) bk.stage(dst, src, (size_t)n * sizeof(T));  // one-shot solve: all arrays go up together
                else bk.h2d(dst, src, (size_t)n * sizeof(T));
            }
            return true;
        };
        bool ok = up(d->ctg_off, b->ctg_off, d->C + 1) && up(d->qs, b->qry_str, d->B) && up(d->qe, b->qry_end, d->B) &&
                  up(d->rs, b->ref_str, d->B) && up(d->re, b->ref_end, d->B) && up(d->qtot, b->qry_total, d->B) &&
                  up(d->chr, b->ref_chr, d->B) && up(d->fwd, b->aln_fwd, d->B) && up(d->mapq, b->map_qul, d->B) &&
                  up(d->run_off, b->run_off, d->B + 1) && up(d->run_ql, b->run_ql, d->R) &&
                  up(d->run_qr, b->run_qr, d->R) && up(d->run_rl, b->run_rl, d->R);
        if (!ok) {
            free_batch(d);
            err = "device allocation failed while staging the batch";
            return AA_ERR_NOMEM;
        }
        if (pooled) bk.flush_staged();
        if (pooled) {
            d->h_ctg_off = b->ctg_off;
            d->h_qs = b->qry_str;
            d->h_qe = b->qry_end;
        } else {
            d->own_ctg_off.assign(b->ctg_off, b->ctg_off + d->C + 1);
            d->own_qs.assign(b->qry_str, b->qry_str + d->B);
            d->own_qe.assign(b->qry_end, b->qry_end + d->B);
            d->h_ctg_off = d->own_ctg_off.data();
            d->h_qs = d->own_qs.data();
            d->h_qe = d->own_qe.data();
        }
        if (!pooled) {  // resident: the host sort (OC1) and the contig order are part of the upload
            host_sort_perm(*d, d->perm, d->sorted_index, d->ctg_order, bk.host_threads());
            d->d_perm = (int32_t *)bk.alloc_persistent((size_t)std::max<int64_t>(d->B, 1) * 4);
            d->d_ord = (int32_t *)bk.alloc_persistent((size_t)std::max<int64_t>(d->C, 1) * 4);
            if (!d->d_perm || !d->d_ord) {
                if (d->d_perm) d->owned.push_back(d->d_perm);
                if (d->d_ord) d->owned.push_back(d->d_ord);
                free_batch(d);
                err = "device allocation failed while staging the batch";
                return AA_ERR_NOMEM;
            }
            d->owned.push_back(d->d_perm);
            d->owned.push_back(d->d_ord);
            bk.h2d(d->d_perm, d->perm.data(), (size_t)d->B * 4);
            bk.h2d(d->d_ord, d->ctg_order.data(), (size_t)d->C * 4);
            d->sorted = true;
        }
        if (!pooled) bk.sync();  // (the staged copies of a one-shot solve are ordered before its kernels on the stream)
        out = d;
        return AA_OK;
    }
    // One-shot staging of a SUBSET of the batch's contigs (ascending ids), straight from the caller's arrays: what a
    // shard of aa_solve_multi solves.  The pieces of every contig go into the pinned buffer one after the other, so the
    // shard is never materialised on the host; only its offsets and sort keys (needed by the host sort) are built here.
    aa_status upload_subset(const aa_batch *b, const int64_t *ctgs, int64_t n_ctgs, DevBatch *&out) {
        std::string v = validate_batch(b);
        if (v.empty() && (!ctgs || n_ctgs <= 0)) v = "empty contig subset";
        for (int64_t k = 0; v.empty() && k < n_ctgs; k++)
            if (ctgs[k] < 0 || ctgs[k] >= b->n_ctg || (k > 0 && ctgs[k] <= ctgs[k - 1])) v = "contig subset not ascending / out of range";
        if (!v.empty()) {
            err = v;
            return AA_ERR_INVALID;
        }
        DevBatch *d = new DevBatch();
        d->pooled = true;
        d->C = n_ctgs;
        d->own_ctg_off.resize((size_t)n_ctgs + 1);
        int64_t B = 0, R = 0;
        for (int64_t k = 0; k < n_ctgs; k++) {
            const int64_t b0 = b->ctg_off[ctgs[k]], b1 = b->ctg_off[ctgs[k] + 1];
            d->own_ctg_off[(size_t)k] = B;
            B += b1 - b0;
            R += b->run_off[b1] - b->run_off[b0];
        }
        d->own_ctg_off[(size_t)n_ctgs] = B;
        d->B = B;
        d->R = R;
        d->own_qs.resize((size_t)B);
        d->own_qe.resize((size_t)B);
        std::vector<int64_t> &roff = d->own_run_off;
        roff.resize((size_t)B + 1);
        auto dev = [&](auto *&dst, int64_t n) {
            using T = std::remove_pointer_t<std::remove_reference_t<decltype(dst)>>;
            dst = (T *)bk.alloc_bytes((size_t)(n > 0 ? n : 1) * sizeof(T));
            return dst != nullptr;
        };
        bool ok = dev(d->ctg_off, d->C + 1) && dev(d->qs, B) && dev(d->qe, B) && dev(d->rs, B) && dev(d->re, B) && dev(d->qtot, B) &&
                  dev(d->chr, B) && dev(d->fwd, B) && dev(d->mapq, B) && dev(d->run_off, B + 1) && dev(d->run_ql, R) &&
                  dev(d->run_qr, R) && dev(d->run_rl, R);
        if (!ok) {
            free_batch(d);
            err = "device allocation failed while staging the batch";
            return AA_ERR_NOMEM;
        }
        // host side: the shard's sort keys and rebased run offsets; device side: one array after the other, so that the
        // pieces of an array are contiguous in the staging buffer and go up in one copy (bk.stage coalesces them)
        std::vector<int64_t> at_of((size_t)n_ctgs + 1, 0), rat_of((size_t)n_ctgs + 1, 0);
        for (int64_t k = 0; k < n_ctgs; k++) {
            const int64_t b0 = b->ctg_off[ctgs[k]], b1 = b->ctg_off[ctgs[k] + 1];
            at_of[(size_t)k + 1] = at_of[(size_t)k] + (b1 - b0);
            rat_of[(size_t)k + 1] = rat_of[(size_t)k] + (b->run_off[b1] - b->run_off[b0]);
        }
        for (int64_t k = 0; k < n_ctgs; k++) {
            const int64_t b0 = b->ctg_off[ctgs[k]], n = at_of[(size_t)k + 1] - at_of[(size_t)k], at = at_of[(size_t)k];
            const int64_t r0 = b->run_off[b0], rat = rat_of[(size_t)k];
            std::memcpy(d->own_qs.data() + at, b->qry_str + b0, (size_t)n * 8);
            std::memcpy(d->own_qe.data() + at, b->qry_end + b0, (size_t)n * 8);
            for (int64_t i = 0; i < n; i++) roff[(size_t)(at + i)] = b->run_off[b0 + i] - r0 + rat;
        }
        auto stage_blocks = [&](auto *dst, const auto *src) {
            for (int64_t k = 0; k < n_ctgs; k++)
                bk.stage(dst + at_of[(size_t)k], src + b->ctg_off[ctgs[k]], (size_t)(at_of[(size_t)k + 1] - at_of[(size_t)k]) * sizeof(*src));
        };
        auto stage_runs = [&](auto *dst, const auto *src) {
            for (int64_t k = 0; k < n_ctgs; k++)
                bk.stage(dst + rat_of[(size_t)k], src + b->run_off[b->ctg_off[ctgs[k]]], (size_t)(rat_of[(size_t)k + 1] - rat_of[(size_t)k]) * sizeof(*src));
        };
        stage_blocks(d->rs, b->ref_str);
        stage_blocks(d->re, b->ref_end);
        stage_blocks(d->qtot, b->qry_total);
        stage_blocks(d->chr, b->ref_chr);
        stage_blocks(d->fwd, b->aln_fwd);
        stage_blocks(d->mapq, b->map_qul);
        if (R > 0) {
            stage_runs(d->run_ql, b->run_ql);
            stage_runs(d->run_qr, b->run_qr);
            stage_runs(d->run_rl, b->run_rl);
        }
        roff[(size_t)B] = R;
        bk.stage(d->ctg_off, d->own_ctg_off.data(), (size_t)(d->C + 1) * 8);
        bk.stage(d->qs, d->own_qs.data(), (size_t)B * 8);
        bk.stage(d->qe, d->own_qe.data(), (size_t)B * 8);
        bk.stage(d->run_off, roff.data(), (size_t)(B + 1) * 8);
        bk.flush_staged();
        d->h_ctg_off = d->own_ctg_off.data();
        d->h_qs = d->own_qs.data();
        d->h_qe = d->own_qe.data();
        out = d;
        return AA_OK;
    }
    void free_batch(DevBatch *d) {
        if (!d) return;
        for (void *p : d->owned) bk.free_persistent(p);
        delete d;
    }

#define AA_BK_CHECK()                                             \
    do {                                                          \
        if (!bk.ok()) {                                           \
            err = bk.error();                                     \
            return bk_out_of_memory() ? AA_ERR_NOMEM : AA_ERR_CUDA; \
        }                                                         \
    } while (0)
    // (a failed workspace allocation is reported as AA_ERR_NOMEM, as the header documents; backends without the notion say no)
    template <class B = BK>
    auto bk_oom_impl(int) -> decltype(std::declval<B &>().oom, bool()) { return bk.oom; }
    template <class B = BK>
    bool bk_oom_impl(long) { return false; }
    bool bk_out_of_memory() { return bk_oom_impl<BK>(0); }
    // ---- the whole hot path over a staged batch ------------------------------------------------------
    aa_status solve(DevBatch &d, const aa_opts &opt, aa_result *res, bool keep_pool = false) {
        bk.begin_solve(keep_pool);
        aa_stats st;
        std::memset(&st, 0, sizeof st);
        const int64_t C = d.C, B = d.B;
        const int32_t K = opt.max_walks > 0 ? opt.max_walks : 10000;
        if (B >= (int64_t)1 << 31 || C >= (int64_t)1 << 30) {
            err = "batch too large for 32-bit block ids";
            return AA_ERR_NOMEM;
        }
        Ws w;
        std::memset(&w, 0, sizeof w);
        w.C = C;
        w.B = B;
        w.R = d.R;
        w.nsl = opt.non_skip_linkable ? 1 : 0;
        w.K = K;
        w.ctg_off = d.ctg_off;
        w.in_qs = d.qs;
        w.in_qe = d.qe;
        w.in_rs = d.rs;
        w.in_re = d.re;
        w.in_qtot = d.qtot;
        w.in_chr = d.chr;
        w.in_fwd = d.fwd;
        w.in_mapq = d.mapq;
        w.run_off = d.run_off;
        w.run_ql = d.run_ql;
        w.run_qr = d.run_qr;
        w.run_rl = d.run_rl;

        // ---- phase 0: host sort (OC1) + device gather ----
        bk.phase_begin(PH_SORT);
        std::vector<int32_t> perm_l, sorted_index_l, ctg_order_l;
        int32_t *d_perm = d.d_perm, *d_ord = d.d_ord;
        if (!d.sorted) {
            host_sort_perm(d, perm_l, sorted_index_l, ctg_order_l, bk.host_threads());
            d_perm = A<int32_t>(B);
            d_ord = A<int32_t>(C);
            bk.h2d(d_perm, perm_l.data(), (size_t)B * 4);
            bk.h2d(d_ord, ctg_order_l.data(), (size_t)C * 4);
        }
        const std::vector<int32_t> &sorted_index = d.sorted ? d.sorted_index : sorted_index_l;
        const std::vector<int32_t> &ctg_order = d.sorted ? d.ctg_order : ctg_order_l;
        w.perm = d_perm;
        w.blk_ctg = A<int32_t>(B);
        w.qs = A<int64_t>(B);
        w.qe = A<int64_t>(B);
        w.rs = A<int64_t>(B);
        w.re = A<int64_t>(B);
        w.qtot = A<int64_t>(B);
        w.chr = A<int32_t>(B);
        w.orig = A<int32_t>(B);
        w.fwd = A<uint8_t>(B);
        w.mapq = A<uint8_t>(B);
        w.run_beg = A<int64_t>(B);
        w.run_cnt = A<int32_t>(B);
        w.first_call = A<int32_t>(B);
        w.part_l = A<int32_t>(B);
        w.part_r = A<int32_t>(B);
        w.status = A<int32_t>(C);
        bk.for_each("gather", B, FnGather{w});
        bk.phase_end(PH_SORT);
        AA_BK_CHECK();

        // ---- phase 1: parts ----
        bk.phase_begin(PH_PARTS);
#if defined(__CUDACC__)
        if (bk.device_kahn()) {  // three segmented scans over the blocks (a 43 k-block contig took one warp 0.9 ms)
            const size_t mark = bk.alloc_mark();
            int64_t *pm = A<int64_t>(B);
            int32_t *bl = A<int32_t>(B), *br = A<int32_t>(B), *brs = A<int32_t>(B);
            if (!pm || !bl || !br || !brs) {
                err = "device allocation failed (parts)";
                return AA_ERR_NOMEM;
            }
            bk.seg_excl_max_i64(w.blk_ctg, w.qe, pm, B);
            bk.for_each("parts_bnd", B, FnPartsBnd{w, pm, bl, br});
            bk.seg_incl_max_i32(w.blk_ctg, bl, w.part_l, B);
            bk.seg_rexcl_min_i32(w.blk_ctg, br, brs, B);
            bk.for_each("parts_fin", B, FnPartsFin{w, brs});
            bk.release_to(mark);  // (stream-ordered: the next allocation's first use comes after these kernels)
        } else
#endif
            bk.for_each_contig("parts", C, FnParts{w, d_ord});
        bk.phase_end(PH_PARTS);
        AA_BK_CHECK();

        // ---- phase 2: pair vertices ----
        bk.phase_begin(PH_PAIRS);
        w.cand_cnt = A<int32_t>(B + 1);
        w.cand_off = A<int64_t>(B + 2);
        bk.zero(w.cand_cnt + B, 4);
        bk.for_each("cand_count", B, FnCandCount{w});
        bk.scan_i32(w.cand_cnt, w.cand_off, B + 1);  // cand_off[B] = total; [B+1] unused
        const int64_t ncand = bk.read_i64(w.cand_off + B);
        w.cand = A<CandRec>(ncand);
        w.cand_ok = A<int32_t>(ncand + 1);
        w.cand_rank = A<int64_t>(ncand + 2);
        bk.zero(w.cand_ok + ncand, 4);
        bk.for_each("cuts", B, FnCuts{w});
        bk.scan_i32(w.cand_ok, w.cand_rank, ncand + 1);
        const int64_t P = bk.read_i64(w.cand_rank + ncand);
        w.pair = A<CandRec>(P);
        w.pair_beg = A<int64_t>(B + 1);
        bk.for_each("compact", std::max<int64_t>(ncand, B + 1), FnCompact{w, ncand});
        w.vtx_off = A<int64_t>(C + 1);
        w.walk_off = A<int64_t>(C + 1);
        bk.for_each("vtx_off", C + 1, FnVtxOff{w});
        const int64_t Vtot = B + P + 2 * C;
        std::vector<int64_t> h_voff((size_t)C + 1);  // the one download of the vertex offsets (sizes the walk scratch, stats)
        bk.d2h(h_voff.data(), w.vtx_off, (size_t)(C + 1) * 8);
        for (int64_t c = 0; c < C; c++)
            if (h_voff[(size_t)c + 1] - h_voff[(size_t)c] >= ((int64_t)1 << 27)) {  // Edge.dst_fl keeps the head in 27 bits
                err = "contig " + std::to_string(c) + " has 2^27 or more graph vertices (not supported: the edge record packs the head in 27 bits)";
                return AA_ERR_NOMEM;
            }
        bk.phase_end(PH_PAIRS);
        AA_BK_CHECK();
        if (Vtot >= ((int64_t)1 << 32) - 1) {
            err = "too many vertices for one device batch";
            return AA_ERR_NOMEM;
        }

        // ---- phase 3: edges ----
        bk.phase_begin(PH_EDGES);
        w.deg = A<int32_t>(Vtot + 1);
        w.eoff = A<int64_t>(Vtot + 2);
        bk.zero(w.deg + Vtot, 4);
        bk.for_each("degree", Vtot, FnDegree{w});
        bk.scan_i32(w.deg, w.eoff, Vtot + 1);
        const int64_t E = bk.read_i64(w.eoff + Vtot);
        if (E >= ((int64_t)1 << 32) - 1) {
            err = "too many edges for one device batch (dense contig: see DESIGN.md, lazy edge mode is future work)";
            return AA_ERR_NOMEM;
        }
        w.edge = A<Edge>(E);
        w.e_src = A<int32_t>(E);
        w.rkey_in = A<uint32_t>(E);
        w.rval_in = A<uint32_t>(E);
        w.rkey = A<uint32_t>(E);
        w.rev_eid = A<uint32_t>(E);
        if (!w.edge || !w.e_src || !w.rkey_in || !w.rval_in || !w.rkey || !w.rev_eid) {
            err = "device allocation failed (edges)";
            return AA_ERR_NOMEM;
        }
        bk.for_each("fill", Vtot, FnFill{w});
        bk.phase_end(PH_EDGES);
        AA_BK_CHECK();

        // ---- phase 4: reverse CSR (stable sort by destination keeps ascending source order: OC4) ----
        bk.phase_begin(PH_REVERSE);
        int vbits = 1;
        while (((int64_t)1 << vbits) < Vtot) vbits++;
        bk.sort_pairs_u32(w.rkey_in, w.rkey, w.rval_in, w.rev_eid, E, vbits);
        w.rev_off = A<int64_t>(Vtot + 1);
        bk.for_each("rev_off", Vtot + 1, FnRevOff{w, E});
        bk.phase_end(PH_REVERSE);
        AA_BK_CHECK();

        // ---- which contigs take the level-synchronous Kahn passes (wide, shallow DAGs: dense contigs) ----
        w.rmode = A<int32_t>(C);
        bk.fill_ff(w.rmode, (size_t)C * 4);
        int32_t n_lm = 0;
        if (bk.device_kahn()) {
            int64_t *d_ce = A<int64_t>(C + 1);
            std::vector<int64_t> h_ce((size_t)C + 1);
            bk.for_each("ctg_edges", C + 1, FnCtgEdges{w, d_ce});
            bk.d2h(h_ce.data(), d_ce, (size_t)(C + 1) * 8);
            h_ce[(size_t)C] = E;
            std::vector<int32_t> h_rmode((size_t)C, -1);
            for (int64_t k = 0; k < C && n_lm < 64; k++) {  // largest contigs first
                const int64_t c = ctg_order[(size_t)k];
                const int64_t Vc = h_voff[(size_t)c + 1] - h_voff[(size_t)c], Ec = h_ce[(size_t)c + 1] - h_ce[(size_t)c];
                if (Vc >= 8192 && Vc < ((int64_t)1 << KL_POSB) && Ec >= 16 * Vc) h_rmode[(size_t)c] = n_lm++;
            }
            if (n_lm > 0) bk.h2d(w.rmode, h_rmode.data(), (size_t)C * 4);
        }

        // ---- phase 5/6: relax + forward order ----
        w.d = A<D4>(Vtot);
        w.best = A<int32_t>(Vtot);
        w.cnt = A<int32_t>(Vtot);
        w.queue = A<int32_t>(Vtot);
        w.order = A<int32_t>(Vtot);
        w.topo = A<int32_t>(Vtot);
        w.amin = A<int32_t>(Vtot);
        w.hroot = A<int32_t>(Vtot);
        w.anom_dis = A<int64_t>(C);
        w.heap_used = A<int64_t>(C);
        if (!w.d || !w.best || !w.cnt || !w.queue || !w.order || !w.topo || !w.amin || !w.hroot) {
            err = "device allocation failed (vertex state)";
            return AA_ERR_NOMEM;
        }
        // the forward Kahn order is only consumed by the walk phases: it runs on the side stream, concurrently
        // with relax / heaps / enum
        if (n_lm > 0) {
            w.kl_cnt = A<int32_t>(Vtot);
            w.kl_last = (unsigned long long *)A<uint64_t>(Vtot);
            w.kl_pos = A<int32_t>(Vtot);
            w.kl_front = A<uint32_t>(Vtot);
            w.kl_next = A<uint32_t>(Vtot);
            w.kl_nnext = A<int32_t>(1);
            w.kl_key_in = (unsigned long long *)A<uint64_t>(Vtot);
            w.kl_key = (unsigned long long *)A<uint64_t>(Vtot);
            w.kl_val = A<uint32_t>(Vtot);
            w.kl_done = A<int32_t>(64);
            if (!w.kl_cnt || !w.kl_last || !w.kl_pos || !w.kl_front || !w.kl_next || !w.kl_key_in || !w.kl_key || !w.kl_val) {
                err = "device allocation failed (level-synchronous passes)";
                return AA_ERR_NOMEM;
            }
        }
        w.seg_mode = nullptr;
        w.TB = 0;
        if (const char *sg = std::getenv("AA_SEG_GUESS")) w.seg_guess = std::atoi(sg);
        std::vector<int64_t> h_boff;
        if (bk.device_kahn()) {
            w.rrec = A<RevRec>(E);
            w.vs = A<VState>(Vtot);
            w.cnt2 = A<int32_t>(Vtot);
            // buckets of SEG_BLOCKS sorted blocks: the units of the segment-parallel relax
            h_boff.assign((size_t)C + 1, 0);
            for (int64_t c = 0; c < C; c++) {
                const int64_t n = d.h_ctg_off[(size_t)c + 1] - d.h_ctg_off[(size_t)c];
                h_boff[(size_t)c + 1] = h_boff[(size_t)c] + (n + SEG_BLOCKS - 1) / SEG_BLOCKS;
            }
            w.TB = h_boff[(size_t)C];
            w.seg_boff = A<int64_t>(C + 1);
            w.seg_bnd = A<int32_t>(w.TB);
            w.seg_ncon = A<int32_t>(w.TB);
            w.seg_con = A<SegCon>(w.TB * SEG_MAXCON);
            w.seg_seed = A<int32_t>(w.TB * 2);
            w.seg_shift = A<SegShift>(w.TB);
            w.seg_mode = A<int32_t>(C);
            w.topo_redo = A<int32_t>(C);
            if (!w.rrec || !w.vs || !w.cnt2 || !w.seg_boff || !w.seg_bnd || !w.seg_ncon || !w.seg_con || !w.seg_seed || !w.seg_shift || !w.seg_mode || !w.topo_redo) {
                err = "device allocation failed (relax records)";
                return AA_ERR_NOMEM;
            }
            bk.h2d(w.seg_boff, h_boff.data(), (size_t)(C + 1) * 8);
            bk.zero(w.seg_mode, (size_t)C * 4);
            bk.zero(w.topo_redo, (size_t)C * 4);
        }
        bk.phase_begin(PH_RELAX);
        if (bk.device_kahn()) {
            bk.for_each("rev_pack", E, FnRevPack{w});
            bk.for_each("relax_init", Vtot, FnRelaxInit{w});
        }
        // chain-like contigs are cut at articulation blocks; every segment is ordered / relaxed by its own warp
        if (bk.device_kahn()) bk.for_each_contig("seg_bounds", w.TB, FnSegBounds{w});
        bk.side_begin();
        bk.phase_begin(PH_TOPO);
        if (bk.device_kahn()) {
            bk.for_each_contig("topo_seg", w.TB, FnTopoSeg{w}, KAHN_SMEM_BYTES);
            bk.for_each_contig("topo_redo", C, FnTopoRedo{w, d_ord}, KAHN_SMEM_BYTES);
        } else {
            bk.for_each_contig("topo", C, FnTopo{w, d_ord}, KAHN_SMEM_BYTES);
        }
        bk.phase_end(PH_TOPO);
        bk.side_end();
        if (bk.device_kahn()) {
            bk.for_each_contig("relax_seg", w.TB, FnRelaxSeg{w}, RELAX_SMEM_C_BYTES);
            bk.for_each_contig("relax_sweep", C, FnRelaxSweep{w, d_ord}, RELAX_SMEM_C_BYTES);
            bk.for_each_contig("relax_redo", C, FnRelaxRedo{w, d_ord}, RELAX_SMEM_BYTES);
            if (std::getenv("AA_SEG_DEBUG")) {
                std::vector<int32_t> hb((size_t)w.TB), hf((size_t)w.TB);
                bk.d2h(hb.data(), w.seg_bnd, (size_t)w.TB * 4);
                bk.d2h(hf.data(), w.seg_ncon, (size_t)w.TB * 4);
                int64_t nseg = 0, nflag = 0;
                for (int64_t i = 0; i < w.TB; i++) {
                    nflag += hf[(size_t)i] != 0;
                }
                for (int64_t c = 0; c < C; c++)
                    for (int64_t m = 0; m < h_boff[(size_t)c + 1] - h_boff[(size_t)c]; m++)
                        nseg += (m == 0 || hb[(size_t)(h_boff[(size_t)c] + m)] >= 0);
                const int64_t cb = ctg_order[0];
                int64_t bs = 0, bf = 0;
                for (int64_t m = 0; m < h_boff[(size_t)cb + 1] - h_boff[(size_t)cb]; m++) {
                    bs += (m == 0 || hb[(size_t)(h_boff[(size_t)cb] + m)] >= 0);
                    bf += hf[(size_t)(h_boff[(size_t)cb] + m)] != 0;
                }
                std::fprintf(stderr, "[aa_seg] buckets %lld segments %lld flagged %lld | largest contig: segments %lld flagged %lld\n",
                             (long long)w.TB, (long long)nseg, (long long)nflag, (long long)bs, (long long)bf);
            }
        } else {
            bk.for_each_contig("relax", C, FnRelax{w, d_ord}, RELAX_SMEM_BYTES);
        }
        if (n_lm > 0) {
            if (!kahn_levels<true>(w, Vtot, n_lm)) return AA_ERR_CUDA;   // d, best, min anom of the dense contigs
            bk.for_each("kl_finish", C, FnKlFinish{w});
            if (!kahn_levels<false>(w, Vtot, n_lm)) return AA_ERR_CUDA;  // their forward order
        }
        if (bk.device_kahn()) bk.for_each("relax_unpack", Vtot, FnRelaxUnpack{w});
        bk.phase_end(PH_RELAX);
        AA_BK_CHECK();

        // walk 0 of every contig is traced on the side stream (after topo), concurrently with heaps / enum
        w.main_pos = A<int32_t>(Vtot);
        w.main_walk = A<int32_t>(Vtot);
        w.m_cs = A<int32_t>(Vtot);
        w.m_cov = A<int64_t>(Vtot);
        w.m_rows = A<int32_t>(Vtot);
        w.m_done = A<uint8_t>(Vtot);
        w.sp_cs = A<int32_t>(Vtot);
        w.sp_rows = A<int32_t>(Vtot);
        w.sp_cov = A<int64_t>(Vtot);
        w.sp_used = A<uint8_t>(Vtot);
        w.mr_blk = A<int32_t>(Vtot);
        w.mr_qs = A<int64_t>(Vtot);
        w.mr_qe = A<int64_t>(Vtot);
        w.mr_rs = A<int64_t>(Vtot);
        w.mr_re = A<int64_t>(Vtot);
        w.m_tot_cov = A<int64_t>(C);
        w.m_tot_rows = A<int32_t>(C);
        w.m_len = A<int32_t>(C);
        w.task_off = A<int64_t>(C + 2);
        if (!w.main_pos || !w.main_walk || !w.m_cs || !w.m_cov || !w.m_rows || !w.m_done || !w.sp_cs || !w.sp_rows ||
            !w.sp_cov || !w.sp_used || !w.mr_blk || !w.mr_qs || !w.mr_qe || !w.mr_rs || !w.mr_re) {
            err = "device allocation failed (main chain)";
            return AA_ERR_NOMEM;
        }
        // ... and so is the rest of the main chain (speculate every step in parallel, resolve sequentially, emit rows in
        // parallel): it needs walk 0 only, and the resolve of the largest contig is a serial chain of several ms.  It
        // has a DP scratch of its own (a few slots: one per resolving warp) and its own work counter.
        int64_t maxV = 3;
        for (int64_t c = 0; c < C; c++) maxV = std::max(maxV, h_voff[(size_t)c + 1] - h_voff[(size_t)c]);
        w.slot_stride = maxV + 1;
        Ws wm = w;
        int64_t S0 = 0;
        {
            const int64_t sb = w.slot_stride * (int64_t)(4 + sizeof(D5) + 4 + 1);
            S0 = std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(C, 2 * bk.max_workers() / 32), bk.scratch_budget() / 4 / std::max<int64_t>(1, sb)));
            wm.sc_walk = nullptr;
            wm.sc_side = nullptr;
            wm.sc_up = A<int32_t>(S0 * w.slot_stride);
            wm.sc_dp = A<D5>(S0 * w.slot_stride);
            wm.sc_pre = A<int32_t>(S0 * w.slot_stride);
            wm.sc_seen = A<uint8_t>(S0 * w.slot_stride);
            wm.task_next = (unsigned long long *)A<int64_t>(1);
            if (!wm.sc_up || !wm.sc_dp || !wm.sc_pre || !wm.sc_seen || !wm.task_next) {
                err = "device allocation failed (main chain scratch)";
                return AA_ERR_NOMEM;
            }
        }
        bk.side_begin();
        bk.for_each_contig("main_trace", C, FnMainTrace{w, d_ord});
        bk.for_each("main_spec", Vtot, FnMainSpec{w});
        bk.zero(wm.task_next, 8);
        bk.workers("main_resolve", S0, FnTasksA0{wm, d_ord});
        bk.for_each("main_rows", Vtot, FnMainRows{w});
        bk.side_end();

        // ---- phase 7: sidetrack heaps (arena doubles on overflow) ----
        bk.phase_begin(PH_HEAPS);
        w.heap_top = (unsigned long long *)A<int64_t>(1);
        w.skey = A<SKey>(E);
        w.child = A<int32_t>(E);
        w.nchild = A<int32_t>(Vtot);
        w.nins = A<int32_t>(Vtot);
        w.cslot = A<uint32_t>(Vtot);
        w.tour_a = A<Tour>(2 * Vtot);
        w.tour_b = A<Tour>(2 * Vtot);
        w.bkey_in = A<uint64_t>(Vtot);
        w.bkey = A<uint64_t>(Vtot);
        w.bval_in = A<uint32_t>(Vtot);
        w.bfs_vtx = A<uint32_t>(Vtot);
        w.bfspos = A<int32_t>(Vtot);
        w.ntree = A<int32_t>(C);
        w.vinfo = A<VInfo>(Vtot);
        w.ins_cnt = A<int32_t>(Vtot + 1);
        w.ins_off = A<int64_t>(Vtot + 2);
        w.root_at = A<int32_t>(Vtot);
        if (!w.skey || !w.child || !w.nchild || !w.nins || !w.cslot || !w.tour_a || !w.tour_b || !w.bkey_in || !w.bkey ||
            !w.bval_in || !w.bfs_vtx || !w.bfspos || !w.ntree || !w.vinfo || !w.ins_cnt || !w.ins_off || !w.root_at) {
            err = "device allocation failed (sidetrack keys / tree order)";
            return AA_ERR_NOMEM;
        }
        bk.for_each("heap_prep", Vtot, FnHeapPrep{w});
        int64_t n_ins = 0;
        // BFS order of every contig's shortest-path tree, in parallel: Euler tour -> list ranking by pointer
        // jumping -> sort by (contig, depth, preorder); then the flat stream of inserts in that order
        {
            int64_t maxV = 3;
            for (int64_t c = 0; c < C; c++) maxV = std::max(maxV, h_voff[(size_t)c + 1] - h_voff[(size_t)c]);
            int kb = 1, cb = 1, rounds = 1;
            while (((int64_t)1 << kb) <= maxV + 1) kb++;
            while (((int64_t)1 << cb) < C) cb++;
            while (((int64_t)1 << rounds) < 2 * maxV) rounds++;
            if (2 * kb + cb > 64) {
                err = "batch shape not supported by the tree-order sort key (contig count x largest contig)";
                return AA_ERR_NOMEM;
            }
            w.key_bits = kb;
            w.cdepth = A<int32_t>(C);
            w.hmode = A<int32_t>(C);
            bk.zero(w.cdepth, (size_t)C * 4);
            bk.zero(w.hmode, (size_t)C * 4);
            bk.for_each("tour_build", Vtot, FnTourBuild{w});
            Tour *src = w.tour_a, *dst = w.tour_b;
            for (int r = 0; r < rounds; r++) {
                bk.for_each("tour_jump", 2 * Vtot, FnTourJump{src, dst});
                std::swap(src, dst);
            }
            bk.for_each("bfs_key", Vtot, FnBfsKey{w, src});
            bk.sort_pairs_u64(w.bkey_in, w.bkey, w.bval_in, w.bfs_vtx, Vtot, 2 * kb + cb);
            bk.for_each("bfs_pos", Vtot, FnBfsPos{w});
            bk.zero(w.ins_cnt + Vtot, 4);
            bk.for_each("vinfo", Vtot, FnVInfo{w});
            bk.scan_i32(w.ins_cnt, w.ins_off, Vtot + 1);
            n_ins = bk.read_i64(w.ins_off + Vtot);
            w.ins = A<InsKey>(n_ins);
            if (!w.ins) {
                err = "device allocation failed (insert stream)";
                return AA_ERR_NOMEM;
            }
            bk.for_each("ins_fill", Vtot, FnInsFill{w});
            bk.fill_ff(w.hroot, (size_t)Vtot * 4);  // vertices outside the tree have no heap
#ifdef AA_HEAP_DUMP  // tools/heap_lab only: the insert stream of the largest contig, for the stand-alone kernel harness
            if (const char *dp = std::getenv("AA_HEAP_DUMP_PATH")) {
                const int64_t c = ctg_order[0];
                const int64_t v0 = h_voff[(size_t)c], Vc = h_voff[(size_t)c + 1] - v0;
                int32_t nt = 0;
                bk.d2h(&nt, w.ntree + c, 4);
                std::vector<VInfo> hv((size_t)nt);
                bk.d2h(hv.data(), w.vinfo + v0, (size_t)nt * sizeof(VInfo));
                int64_t io[2];
                bk.d2h(&io[0], w.ins_off + v0, 8);
                bk.d2h(&io[1], w.ins_off + v0 + nt, 8);
                std::vector<InsKey> hk((size_t)(io[1] - io[0]));
                bk.d2h(hk.data(), w.ins + io[0], hk.size() * sizeof(InsKey));
                for (auto &v : hv) v.ins_beg -= (uint32_t)io[0];
                if (FILE *f = std::fopen(dp, "wb")) {
                    const int64_t hdr[4] = {nt, (int64_t)hk.size(), Vc, c};
                    std::fwrite(hdr, 8, 4, f);
                    std::fwrite(hv.data(), sizeof(VInfo), hv.size(), f);
                    std::fwrite(hk.data(), sizeof(InsKey), hk.size(), f);
                    std::fclose(f);
                }
            }
#endif
        }
        // shallow, wide trees (dense contigs) are built level by level with one warp per vertex; everything else by the
        // streaming builder (one warp per contig)
        std::vector<int32_t> m1;             // contigs in level mode
        std::vector<int64_t> h_lvl;          // [m1][depth + 1] first BFS slot of each depth (last: end of the tree)
        int32_t lvl_per = 0;
        if (bk.device_kahn()) {
            std::vector<int32_t> h_depth((size_t)C);
            bk.d2h(h_depth.data(), w.cdepth, (size_t)C * 4);
            std::vector<int32_t> h_mode((size_t)C, 0);
            for (int64_t k = 0; k < C && m1.size() < 64; k++) {  // largest contigs first
                const int64_t c = ctg_order[(size_t)k];
                const int64_t Vc = h_voff[(size_t)c + 1] - h_voff[(size_t)c];
                if (Vc >= 8192 && h_depth[(size_t)c] >= 1 && h_depth[(size_t)c] <= 64) {
                    h_mode[(size_t)c] = 1;
                    m1.push_back((int32_t)c);
                    lvl_per = std::max(lvl_per, h_depth[(size_t)c] + 1);
                }
            }
            if (!m1.empty()) {
                bk.h2d(w.hmode, h_mode.data(), (size_t)C * 4);
                int32_t *d_m1 = A<int32_t>((int64_t)m1.size());
                int64_t *d_lvl = A<int64_t>((int64_t)m1.size() * lvl_per);
                bk.h2d(d_m1, m1.data(), m1.size() * 4);
                bk.for_each("lvl_off", (int64_t)m1.size() * lvl_per, FnLvlOff{w, d_m1, lvl_per, d_lvl});
                h_lvl.resize(m1.size() * (size_t)lvl_per);
                bk.d2h(h_lvl.data(), d_lvl, h_lvl.size() * 8);
            }
        }
        const bool any_m1 = !m1.empty();
        w.lvl_overflow = A<int32_t>(1);
        // tree leaves with inserts leave the serial builder (they feed no other heap): one warp each, after it
        int64_t n_leaf = 0;
        w.leaf_flag = A<int32_t>(Vtot + 1);
        w.leaf_off = A<int64_t>(Vtot + 2);
        w.leaf_base = A<int32_t>(Vtot);
        if (bk.device_kahn()) {
            w.vcnt = A<int32_t>(Vtot + 1);
            w.vbase = A<int64_t>(Vtot + 2);
        }
        bk.zero(w.leaf_flag + Vtot, 4);
        bk.for_each("leaf_flag", Vtot, FnLeafFlag{w});
        bk.scan_i32(w.leaf_flag, w.leaf_off, Vtot + 1);
        n_leaf = bk.read_i64(w.leaf_off + Vtot);
        w.leaf_list = A<uint32_t>(n_leaf);
        bk.for_each("leaf_list", Vtot, FnLeafList{w});
        // the operation stream of the serial builder (f_heaps_chain): inserts of the chain vertices + one id reservation per
        // leaf, in BFS order; every vertex is pointed at the nearest chain vertex among itself and its tree ancestors
        w.op_cnt = A<int32_t>(Vtot + 1);
        w.op_off = A<int64_t>(Vtot + 2);
        w.chain_flag = A<int32_t>(Vtot + 1);
        w.chain_ord = A<int64_t>(Vtot + 2);
        w.owner = A<int32_t>(Vtot);
        w.chain_root = A<int32_t>(Vtot);
        w.ops = A<HOp>(n_ins);
        w.leaf_need = A<int32_t>(Vtot + 1);
        w.leaf_lp = A<int64_t>(Vtot + 2);
        w.chain_slot = A<int32_t>(Vtot);
        w.resv_base = A<int32_t>(Vtot + C + 1);
        if (!w.leaf_flag || !w.leaf_off || !w.leaf_base || !w.leaf_list || !w.op_cnt || !w.op_off || !w.chain_flag || !w.chain_ord ||
            !w.owner || !w.chain_root || !w.ops || !w.leaf_need || !w.leaf_lp || !w.chain_slot || !w.resv_base) {
            err = "device allocation failed (heap operation stream)";
            return AA_ERR_NOMEM;
        }
        bk.zero(w.op_cnt + Vtot, 4);
        bk.zero(w.chain_flag + Vtot, 4);
        bk.zero(w.leaf_need + Vtot, 4);
        bk.for_each("ops_class", Vtot, FnOpsClass{w});
        bk.scan_i32(w.op_cnt, w.op_off, Vtot + 1);
        bk.scan_i32(w.chain_flag, w.chain_ord, Vtot + 1);
        bk.scan_i32(w.leaf_need, w.leaf_lp, Vtot + 1);
        bk.for_each("chain_slot", Vtot, FnChainSlot{w});
        {
            int64_t maxV = 3;
            for (int64_t c = 0; c < C; c++) maxV = std::max(maxV, h_voff[(size_t)c + 1] - h_voff[(size_t)c]);
            int jumps = 1;
            while (((int64_t)1 << jumps) < maxV) jumps++;
            for (int r = 0; r < (jumps + 1) / 2; r++) bk.for_each("owner_jump", Vtot, FnOwnerJump{w});  // two hops per launch
        }
        bk.for_each("ops_fill", Vtot, FnOpsFill{w});
        // node cache of the serial builder: as large as still lets every contig of the batch be resident at once
        w.heap_cache_bits = C <= 2 * 148 ? 11 : (C <= 4 * 148 ? 10 : (C <= 6 * 148 ? 9 : 8));
        // (a leaf of a streaming-mode contig reserves 32 ids per insert, at most a chunk; level mode wastes < 64 per vertex)
        int64_t hcap = 6 * E + C * (int64_t)HEAP_CHUNK + 256 * n_leaf + 64 * (any_m1 ? Vtot : 0) + ((int64_t)1 << 20);
        std::vector<int32_t> h_status((size_t)C);
        auto h_blk_of_order = [&](int64_t k) { return d.h_ctg_off[(size_t)ctg_order[(size_t)k] + 1] - d.h_ctg_off[(size_t)ctg_order[(size_t)k]]; };
        int64_t heap_top_h = 0;  // node ids handed out (device path)
        if (const char *tn = std::getenv("AA_TUNE")) w.heaps_variant = std::atoi(tn);  // 1: no serial steps, 2: no expansion records, 4: one backlog region, 8: wide keys for every contig, 16: no two-group pipelining, 32: pipelining for every batch (tests)
        // large group: the contigs with long serial chains (>= 4096 blocks, at most one per SM)
        int64_t n_big = 0;
        while (n_big < C && n_big < 148 && h_blk_of_order(n_big) >= 4096) n_big++;
        if ((w.heaps_variant & 32) && C >= 2 && n_big == 0) n_big = (C + 3) / 4;
        // pipelining needs the streaming builder for every contig (no level mode) and pays when the large group is a small part of
        // the batch: at least 8 contigs per SM-filling wave are left for the small group
        const bool overlap = bk.device_kahn() && !any_m1 && !(w.heaps_variant & 16) && n_big > 0 && n_big < C &&
                             (C - n_big >= 4 * n_big || (w.heaps_variant & 32));
        std::vector<int8_t> h_grp;
        if (overlap) {
            h_grp.assign((size_t)C, 0);
            for (int64_t k = 0; k < n_big; k++) h_grp[(size_t)ctg_order[(size_t)k]] = 1;
            int8_t *d_grp = A<int8_t>(C);
            if (!d_grp) {
                err = "device allocation failed (contig groups)";
                return AA_ERR_NOMEM;
            }
            bk.h2d(d_grp, h_grp.data(), (size_t)C);
            w.grp = d_grp;
        }
        w.grp_sel = -1;
        const int64_t WK = C * (int64_t)K;
        // the arrays of the enumeration (inside an attempt when the two phases are pipelined: a retry releases them with the arena)
        auto alloc_enum = [&]() -> bool {
            w.n_walk = A<int32_t>(C + 1);
            w.wdist = A<D4>(WK);
            w.wlast = A<int32_t>(WK);
            w.ent_node = A<int32_t>(3 * WK);
            w.ent_prev = A<int32_t>(3 * WK);
            w.pq = A<PQEnt>(3 * WK);
            if (!w.wdist || !w.wlast || !w.ent_node || !w.ent_prev || !w.pq) {
                err = "device allocation failed (walk enumeration)";
                return false;
            }
            w.pq_far = nullptr;
            if (bk.device_kahn()) {
                w.enext = A<ENext>(E);
                if (!w.enext) {
                    err = "device allocation failed (enumeration)";
                    return false;
                }
                // the second backlog region is an optimisation: without the memory for it the queue keeps one region
                if (!(w.heaps_variant & 4) && 3 * WK * (int64_t)sizeof(PQEnt) <= bk.scratch_budget()) w.pq_far = A<PQEnt>(3 * WK);
            }
            return true;
        };
        bool enum_done = false;
        for (int attempt = 0;; attempt++) {
            if (hcap > 0x7ffffff0LL) hcap = 0x7ffffff0LL;
            w.Hcap = hcap;
            const size_t arena_mark = bk.alloc_mark();
            w.hn = A<HNode>(hcap);
            w.hn_eid = A<int32_t>(hcap);
            w.chunk_ctg = bk.device_kahn() ? A<int32_t>(hcap / 64 + 66) : nullptr;
            if (!w.hn || !w.hn_eid || (bk.device_kahn() && !w.chunk_ctg)) {
                err = "device allocation failed (sidetrack heap arena)";
                return AA_ERR_NOMEM;
            }
            if (bk.device_kahn()) {
                w.hn_key = (unsigned long long *)A<uint64_t>(hcap);
                if (!w.hn_key) {
                    err = "device allocation failed (heap node order keys)";
                    return AA_ERR_NOMEM;
                }
                bk.fill_ff(w.hn_key, (size_t)hcap * 8);
                bk.fill_ff(w.hn_eid, (size_t)hcap * 4);  // ids that are never allocated stay -1 (f_xrec skips them)
                bk.fill_ff(w.chunk_ctg, (size_t)(hcap / 64 + 66) * 4);  // (-1: no owner yet)
                bk.zero(w.vcnt, (size_t)(Vtot + 1) * 4);
            }
            bk.zero(w.heap_top, 8);
            bk.zero(w.heap_used, (size_t)C * 8);
            bk.zero(w.lvl_overflow, 4);
            // The contigs with long serial chains get a kernel of their own with the largest node cache (180 KB of shared memory per
            // CTA leaves room for one small contig beside them, so their warps are not slowed by seven neighbours); the rest run
            // concurrently on the aux stream.
            if (overlap) {
                // Two-group pipelining: the passes between the heaps and the enumeration, and the enumeration itself, are launched per
                // group, so the small contigs enumerate on the aux stream WHILE the large contigs' heap chains are still being built
                // (on the 2 080-contig input the largest chain takes 31 ms and everything else of that phase 12).
                Ws wA = w, wB = w;
                wA.grp_sel = 1;
                wB.grp_sel = 0;
                wA.heap_cache_bits = 12;
                bk.aux_begin();
                bk.for_each_contig("heaps_small", C - n_big, FnHeaps{wB, d_ord + n_big}, heaps_chain_smem_bytes(wB.heap_cache_bits));
                bk.for_each("root_fill", Vtot, FnRootFill{wB});
                if (n_leaf > 0) bk.for_each_contig("heaps_leaf", n_leaf, FnHeapsLeaf{wB});
                bk.aux_end();
                bk.for_each_contig("heaps", n_big, FnHeaps{wA, d_ord}, heaps_chain_smem_bytes(wA.heap_cache_bits));
                bk.for_each("root_fill", Vtot, FnRootFill{wA});
                if (n_leaf > 0) bk.for_each_contig("heaps_leaf", n_leaf, FnHeapsLeaf{wA});
                // the small group is done long before the large one: read its outcome on the aux stream (the main stream runs on)
                bk.aux_enter();
                int32_t h_lo = 0;
                bk.d2h(&h_lo, w.lvl_overflow, 4);
                bk.d2h(h_status.data(), w.status, (size_t)C * 4);
                AA_BK_CHECK();
                bool overflow = h_lo != 0;
                for (int32_t st3 : h_status) overflow = overflow || st3 == 3;
                int64_t top_b = 0;
                if (!overflow) {
                    top_b = bk.read_i64((const int64_t *)w.heap_top);  // ids handed out so far: all of the small group's, some of the large one's
                    if (!alloc_enum()) return AA_ERR_NOMEM;
                    w.xrec = nullptr;
                    if (!(w.heaps_variant & 2) && top_b > 0 && top_b * (int64_t)sizeof(XRec) <= bk.scratch_budget()) w.xrec = A<XRec>(top_b);
                    wB = w;
                    wB.grp_sel = 0;
                    bk.for_each("enext", Vtot, FnENext{wB});
                    if (w.xrec) bk.for_each("xrec", top_b, FnXRec{wB});
                    bk.for_each_contig("enum_small", C - n_big, FnEnum{wB, d_ord + n_big}, ENUM_SMEM_BYTES);
                }
                bk.aux_end();
                // ... and then the large group's, on the main stream
                bk.d2h(&h_lo, w.lvl_overflow, 4);
                bk.d2h(h_status.data(), w.status, (size_t)C * 4);
                AA_BK_CHECK();
                overflow = overflow || h_lo != 0;
                for (int32_t st3 : h_status) overflow = overflow || st3 == 3;
                if (!overflow) {
                    const int64_t top = bk.read_i64((const int64_t *)w.heap_top);
                    heap_top_h = top;
                    bk.phase_end(PH_HEAPS);
                    bk.phase_begin(PH_ENUM);
                    wA = w;
                    wA.grp_sel = 1;
                    if (w.xrec) {
                        // the records of the ids handed out since then go directly behind the first array when the allocator is in one
                        // block (heap_top moves in whole chunks, so the first array ends on the allocator's alignment); else this group
                        // gets an array of full length of its own
                        if (top * (int64_t)sizeof(XRec) > bk.scratch_budget()) {
                            wA.xrec = nullptr;
                        } else if (top > top_b) {
                            XRec *more = A<XRec>(top - top_b);
                            if (more != w.xrec + top_b) wA.xrec = A<XRec>(top);  // (nullptr: this group expands without records)
                        }
                    }
                    bk.for_each("enext", Vtot, FnENext{wA});
                    if (wA.xrec) bk.for_each("xrec", top, FnXRec{wA});
                    bk.for_each_contig("enum", n_big, FnEnum{wA, d_ord}, ENUM_SMEM_BYTES);
                    bk.aux_join();
                    bk.phase_end(PH_ENUM);
                    AA_BK_CHECK();
                    enum_done = true;
                    break;
                }
                bk.aux_join();
                bk.sync();  // (an enumeration of the small group may still be reading the arena)
                if (hcap >= 0x7ffffff0LL || attempt > 8) {
                    err = "sidetrack heap arena exhausted (contig too dense for one device)";
                    return AA_ERR_NOMEM;
                }
                bk.release_to(arena_mark);
                hcap *= 4;
                continue;
            }
            if (n_big > 0 && n_big < C) {
                Ws wb = w;
                wb.heap_cache_bits = 12;
                bk.aux_begin();
                bk.for_each_contig("heaps_small", C - n_big, FnHeaps{w, d_ord + n_big}, heaps_chain_smem_bytes(w.heap_cache_bits));
                bk.aux_end();
                bk.for_each_contig("heaps", n_big, FnHeaps{wb, d_ord}, heaps_chain_smem_bytes(wb.heap_cache_bits));
                bk.aux_join();
            } else {
                bk.for_each_contig("heaps", C, FnHeaps{w, d_ord}, heaps_chain_smem_bytes(w.heap_cache_bits));
            }
            bk.for_each("root_fill", Vtot, FnRootFill{w});
            bool overflow = false;
            if (n_leaf > 0) bk.for_each_contig("heaps_leaf", n_leaf, FnHeapsLeaf{w});
#if defined(__CUDACC__)
            if (any_m1)
                for (int32_t d = 1; d < lvl_per; d++)
                    for (size_t k = 0; k < m1.size(); k++) {
                        const int64_t lo = h_lvl[k * (size_t)lvl_per + (size_t)d - 1], hi = h_lvl[k * (size_t)lvl_per + (size_t)d];
                        if (hi > lo) bk.for_each_contig("heaps_level", hi - lo, FnHeapsLevel{w, lo});
                    }
            if (n_leaf > 0 || any_m1) {
                int32_t h_lo = 0;
                bk.d2h(&h_lo, w.lvl_overflow, 4);
                overflow = h_lo != 0;
            }
#endif
            bk.d2h(h_status.data(), w.status, (size_t)C * 4);
            AA_BK_CHECK();
            for (int32_t s : h_status) overflow = overflow || s == 3;
            if (!overflow && bk.device_kahn()) {
                // (owner slot, number) of every node -> its rank in the sequential allocation order
                const int64_t top = bk.read_i64((const int64_t *)w.heap_top);
                heap_top_h = top;
                bk.scan_i32(w.vcnt, w.vbase, Vtot + 1);
                bk.for_each("node_rank", top, FnNodeRank{w});
            }
            if (!overflow) break;
            if (hcap >= 0x7ffffff0LL || attempt > 8) {
                err = "sidetrack heap arena exhausted (contig too dense for one device)";
                return AA_ERR_NOMEM;
            }
            bk.release_to(arena_mark);  // give the arena arrays (and the scan scratch) back before growing
            hcap *= 4;
        }
        if (!enum_done) {
            bk.phase_end(PH_HEAPS);
            AA_BK_CHECK();

            // ---- phase 8: enumeration ----
            bk.phase_begin(PH_ENUM);
            if (!alloc_enum()) return AA_ERR_NOMEM;
            if (bk.device_kahn()) {
                bk.for_each("enext", Vtot, FnENext{w});
                // expansion records: one load per pop instead of a chain of three (skipped when the arena is too large for them)
                w.xrec = nullptr;
                if (!(w.heaps_variant & 2) && heap_top_h > 0 && heap_top_h * (int64_t)sizeof(XRec) <= bk.scratch_budget()) {
                    w.xrec = A<XRec>(heap_top_h);
                    if (!w.xrec) {
                        err = "device allocation failed (expansion records)";
                        return AA_ERR_NOMEM;
                    }
                    bk.for_each("xrec", heap_top_h, FnXRec{w});
                }
            }
            bk.for_each_contig("enum", C, FnEnum{w, d_ord}, ENUM_SMEM_BYTES);
            bk.phase_end(PH_ENUM);
            AA_BK_CHECK();
        }

        // ---- phase 9: plan ----
        bk.phase_begin(PH_PLAN);
        w.task = A<Task>(2 * WK);
        w.n_task = A<int32_t>(C + 1);
        w.n_tie = A<int32_t>(C);
        w.last_group = A<int32_t>(C);
        if (!w.task) {
            err = "device allocation failed (task plan)";
            return AA_ERR_NOMEM;
        }
        bk.zero(w.n_task + C, 4);
        bk.for_each_contig("plan", C, FnPlan{w, d_ord});
        bk.scan_i32(w.n_task, w.task_off, C + 1);
        const int64_t NT = bk.read_i64(w.task_off + C);
        w.n_tasks_total = NT;
        w.tasks = A<Task>(NT);
        w.task_cov = A<int64_t>(NT);
        w.task_rows = A<int32_t>(NT);
        bk.for_each_contig("task_compact", C, FnTaskCompact{w, d_ord});
        bk.phase_end(PH_PLAN);
        AA_BK_CHECK();

        // ---- phase 10: walks, pass A ----
        bk.side_join();
        bk.phase_begin(PH_WALKS_A);
        const int64_t slot_bytes = w.slot_stride * (int64_t)(4 + 4 + 4 + sizeof(D5) + 4 + 1);
        int64_t S = std::min<int64_t>(std::max<int64_t>(NT, 2 * C), bk.max_workers());
        S = std::max<int64_t>(1, std::min<int64_t>(S, bk.scratch_budget() / std::max<int64_t>(1, slot_bytes)));
        w.sc_walk = A<int32_t>(S * w.slot_stride);
        w.sc_up = A<int32_t>(S * w.slot_stride);
        w.sc_side = A<int32_t>(S * w.slot_stride);
        w.sc_dp = A<D5>(S * w.slot_stride);
        w.sc_pre = A<int32_t>(S * w.slot_stride);
        w.sc_seen = A<uint8_t>(S * w.slot_stride);
        w.task_next = (unsigned long long *)A<int64_t>(1);
        if (!w.sc_walk || !w.sc_up || !w.sc_side || !w.sc_dp || !w.sc_pre || !w.sc_seen) {
            err = "device allocation failed (walk scratch)";
            return AA_ERR_NOMEM;
        }
        // walk 0 of every contig was resolved on the side stream: its totals become task 0
        bk.for_each("main_totals", C, FnMainTotals{w});
        // every other planned walk
        if (NT > C) {
            w.fb_list = A<int32_t>(NT);
            w.fb_n = A<int32_t>(1);
            bk.zero(w.fb_n, 4);
            bk.for_each("walksA1_solo", NT, FnTasksA1Solo{w});  // one thread per task
            int32_t n_fb = 0;
            bk.d2h(&n_fb, w.fb_n, 4);
            w.fb_count = n_fb;
            if (n_fb > 0) {  // the tasks that did not fit the private scratch: one warp each, big scratch
                bk.zero(w.task_next, 8);
                bk.workers("walksA1", std::min<int64_t>(S, n_fb), FnTasksA1{w});
            }
        }
        bk.phase_end(PH_WALKS_A);
        AA_BK_CHECK();

        // ---- phase 11: select ----
        bk.phase_begin(PH_SELECT);
        w.win_out = A<int32_t>(C);
        w.win_alt = A<int32_t>(C);
        w.out_cnt = A<int32_t>(C + 1);
        w.alt_cnt = A<int32_t>(C + 1);
        w.all_cnt = A<int32_t>(C + 1);
        w.out_off = A<int64_t>(C + 2);
        w.alt_off = A<int64_t>(C + 2);
        w.all_path_off = A<int64_t>(C + 2);
        bk.zero(w.out_cnt + C, 4);
        bk.zero(w.alt_cnt + C, 4);
        bk.zero(w.all_cnt + C, 4);
        bk.for_each_contig("select", C, FnSelect{w, d_ord, opt.want_all ? 1 : 0});
        bk.scan_i32(w.out_cnt, w.out_off, C + 1);
        bk.scan_i32(w.alt_cnt, w.alt_off, C + 1);
        bk.scan_i32(w.all_cnt, w.all_path_off, C + 1);
        const int64_t n_out = bk.read_i64(w.out_off + C), n_alt = bk.read_i64(w.alt_off + C);
        const int64_t n_paths = bk.read_i64(w.all_path_off + C);
        w.all_task = A<int32_t>(n_paths);
        w.all_rows = A<int32_t>(n_paths + 1);
        w.all_row_off = A<int64_t>(n_paths + 2);
        bk.zero(w.all_rows + n_paths, 4);
        if (n_paths > 0) bk.for_each_contig("all_list", C, FnAllList{w, d_ord});
        bk.scan_i32(w.all_rows, w.all_row_off, n_paths + 1);
        const int64_t n_all = bk.read_i64(w.all_row_off + n_paths);
        bk.phase_end(PH_SELECT);
        AA_BK_CHECK();

        // ---- phase 12: walks, pass B (rows of the winners) ----
        bk.phase_begin(PH_WALKS_B);
        const int64_t nrow[3] = {n_out, n_alt, n_all};
        for (int k = 0; k < 3; k++) {
            w.r_idx[k] = A<int32_t>(nrow[k]);
            w.r_qs[k] = A<int64_t>(nrow[k]);
            w.r_qe[k] = A<int64_t>(nrow[k]);
            w.r_rs[k] = A<int64_t>(nrow[k]);
            w.r_re[k] = A<int64_t>(nrow[k]);
            w.r_alt[k] = A<uint8_t>(nrow[k]);
            if (!w.r_idx[k] || !w.r_qs[k] || !w.r_qe[k] || !w.r_rs[k] || !w.r_re[k] || !w.r_alt[k]) {
                err = "device allocation failed (result rows; .aln.all.paf too large?)";
                return AA_ERR_NOMEM;
            }
        }
        bk.zero(w.task_next, 8);
        const int64_t n_items = 2 * C + n_paths;
        bk.workers("walksB", std::min<int64_t>(S, n_items), FnTasksB{w, n_items, n_paths});
        bk.phase_end(PH_WALKS_B);
        AA_BK_CHECK();

        // ---- phase 13: download ----
        bk.phase_begin(PH_D2H);
        bk.d2h(h_status.data(), w.status, (size_t)C * 4);
        std::vector<int32_t> h_nwalk((size_t)C);
        std::vector<int64_t> h_heap((size_t)C);
        bk.d2h(h_nwalk.data(), w.n_walk, (size_t)C * 4);
        bk.d2h(h_heap.data(), w.heap_used, (size_t)C * 8);
        bool unsolvable = false;
        st.n_ctg = C;
        st.n_blk = B;
        st.n_run = d.R;
        st.n_pair = P;
        st.n_edge = E;
        st.n_task = NT;
        for (int64_t c = 0; c < C; c++) {
            if (h_status[(size_t)c] == 2) unsolvable = true;
            if (h_status[(size_t)c] == 0) {
                st.n_walk += h_nwalk[(size_t)c];
                st.n_heap += h_heap[(size_t)c];
            }
        }
        for (int64_t c = 0; c < C; c++)
            if (h_status[(size_t)c] != 1) st.n_vtx += h_voff[(size_t)c + 1] - h_voff[(size_t)c];
        if (res) {
            std::memset(res, 0, sizeof *res);
            res->n_ctg = C;
            // one pinned slab for every array of the result when the backend has one to give
            auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
            size_t slab_need = 3 * al((size_t)(C + 1) * 8) + al((size_t)(n_paths + 1) * 8) + al((size_t)B * 4);
            for (int k = 0; k < 3; k++) {
                const size_t m = (size_t)std::max<int64_t>(nrow[k], 1);
                slab_need += al(m * 4) + 4 * al(m * 8) + al(m);
            }
            char *slab = (char *)bk.result_slab(slab_need);
            size_t slab_off = 0;
            auto carve = [&](size_t b) {
                void *q = slab + slab_off;
                slab_off += al(b);
                return q;
            };
            if (slab) {
                res->out_off = (int64_t *)carve((size_t)(C + 1) * 8);
                res->alt_off = (int64_t *)carve((size_t)(C + 1) * 8);
                res->all_path_off = (int64_t *)carve((size_t)(C + 1) * 8);
                res->all_row_off = (int64_t *)carve((size_t)(n_paths + 1) * 8);
                res->sorted_index = (int32_t *)carve((size_t)B * 4);
            } else {
                res->out_off = host_n<int64_t>(C + 1);
                res->alt_off = host_n<int64_t>(C + 1);
                res->all_path_off = host_n<int64_t>(C + 1);
                res->all_row_off = host_n<int64_t>(n_paths + 1);
                res->sorted_index = host_n<int32_t>(B);
            }
            std::memcpy(res->sorted_index, sorted_index.data(), (size_t)B * 4);
            bk.d2h(res->out_off, w.out_off, (size_t)(C + 1) * 8);
            bk.d2h(res->alt_off, w.alt_off, (size_t)(C + 1) * 8);
            bk.d2h(res->all_path_off, w.all_path_off, (size_t)(C + 1) * 8);
            bk.d2h(res->all_row_off, w.all_row_off, (size_t)(n_paths + 1) * 8);
            aa_rows *dst[3] = {&res->out, &res->alt, &res->all};
            for (int k = 0; k < 3; k++) {
                if (slab) {
                    const size_t m = (size_t)std::max<int64_t>(nrow[k], 1);
                    dst[k]->n = nrow[k];
                    dst[k]->ctg_index = (int32_t *)carve(m * 4);
                    dst[k]->qry_str = (int64_t *)carve(m * 8);
                    dst[k]->qry_end = (int64_t *)carve(m * 8);
                    dst[k]->ref_str = (int64_t *)carve(m * 8);
                    dst[k]->ref_end = (int64_t *)carve(m * 8);
                    dst[k]->is_alt = (uint8_t *)carve(m);
                } else {
                    rows_alloc_host(*dst[k], nrow[k]);
                }
                if (nrow[k] > 0) {
                    bk.stage_d2h(dst[k]->ctg_index, w.r_idx[k], (size_t)nrow[k] * 4);
                    bk.stage_d2h(dst[k]->qry_str, w.r_qs[k], (size_t)nrow[k] * 8);
                    bk.stage_d2h(dst[k]->qry_end, w.r_qe[k], (size_t)nrow[k] * 8);
                    bk.stage_d2h(dst[k]->ref_str, w.r_rs[k], (size_t)nrow[k] * 8);
                    bk.stage_d2h(dst[k]->ref_end, w.r_re[k], (size_t)nrow[k] * 8);
                    bk.stage_d2h(dst[k]->is_alt, w.r_alt[k], (size_t)nrow[k]);
                }
            }
            bk.flush_d2h(slab != nullptr);
            if (opt.keep_debug) download_debug(w, res, h_status, h_voff, h_nwalk, E, Vtot);
        }
        bk.phase_end(PH_D2H);
        AA_BK_CHECK();
        bk.end_solve(st);
        // algorithmic bytes (DESIGN.md §"roofline"; SURVEY.md §8(d) with this repo's record sizes)
        {
            double *a = st.algo_bytes_phase;
            const double Bd = (double)B, Rd = (double)d.R, Pd = (double)P, Vd = (double)Vtot, Ed = (double)E;
            const double Hd = (double)st.n_heap, Kd = (double)st.n_walk;
            a[PH_SORT] = 2.0 * 50.0 * Bd;                 // gather: read + write one 50-B block record
            a[PH_PARTS] = 16.0 * Bd + 8.0 * Bd;           // qs,qe read; part_l,part_r write
            a[PH_PAIRS] = 24.0 * Rd + 16.0 * Bd + 2.0 * 48.0 * Pd;  // runs once, keys, pair record write+compact
            a[PH_EDGES] = 50.0 * Bd + 48.0 * Pd + (16.0 + 4.0 + 8.0) * Ed;  // vertex data read, edge+src+key write
            a[PH_REVERSE] = 2.0 * 8.0 * Ed * 4.0 + 8.0 * Vd;   // 4 radix passes over (key,val), offsets
            a[PH_RELAX] = (16.0 + 4.0 + 4.0) * Ed + (24.0 + 4.0) * Vd;      // edge, rev id, src read; d + best write
            a[PH_TOPO] = 16.0 * Ed + 8.0 * Vd;
            a[PH_HEAPS] = 16.0 * Ed + 24.0 * Vd + 44.0 * Hd;  // node 32 + edge id 4 + order key 8
            a[PH_ENUM] = Kd * (32.0 + 36.0 + 3.0 * (32.0 + 32.0 + 8.0) + 28.0);  // pop, node, <=3 (node read, push, entry), out
            a[PH_PLAN] = 24.0 * Kd;
            a[PH_WALKS_A] = 0;  // data dependent; reported as time only
            st.algo_bytes = 0;
            for (int p = 0; p < PH_COUNT; p++) st.algo_bytes += a[p];
        }
        last_stats = st;
        if (res) res->stats = st;
        if (unsolvable) {
            err = "a contig has no src->dest walk (reference: assert, paf_data.cpp:732)";
            return AA_ERR_UNSOLVABLE;
        }
        return AA_OK;
    }

    // level-synchronous Kahn pass over the contigs with rmode >= 0 (see f_kl_* in aa_core.cuh)
    template <bool REV>
    bool kahn_levels(const Ws &w, int64_t Vtot, int32_t n_lm) {
        bk.zero(w.kl_nnext, 4);
        bk.zero(w.kl_done, 64 * 4);
        bk.for_each("kl_init", Vtot, FnKlInit<REV>{w});
        for (;;) {
            int32_t n = 0;
            bk.d2h(&n, w.kl_nnext, 4);
            if (!bk.ok()) {
                err = bk.error();
                return false;
            }
            if (n == 0) break;
            const size_t mark = bk.alloc_mark();  // the sort scratch of one level is reused by the next (stream order)
            bk.for_each("kl_keys", n, FnKlKeys{w});
            bk.sort_pairs_u64((const uint64_t *)w.kl_key_in, (uint64_t *)w.kl_key, w.kl_next, w.kl_val, n, 2 * KL_POSB + 6);
            bk.for_each("kl_assign", n, FnKlAssign<REV>{w, n});
            bk.for_each("kl_count", n_lm, FnKlCount{w, n});
            bk.zero(w.kl_nnext, 4);
            if (REV) bk.for_each_contig("kl_pull", n, FnKlPull{w});
            bk.for_each_contig("kl_expand", n, FnKlExpand<REV>{w});
            bk.release_to(mark);
        }
        return true;
    }

    aa_stats last_stats{};

    void download_debug(const Ws &w, aa_result *res, const std::vector<int32_t> &h_status, const std::vector<int64_t> &h_voff,
                        const std::vector<int32_t> &h_nwalk, int64_t E, int64_t Vtot) {
        const int64_t C = w.C;
        aa_debug *g = (aa_debug *)std::calloc(1, sizeof(aa_debug));
        res->dbg = g;
        std::vector<int64_t> h_eoff((size_t)Vtot + 1);
        bk.d2h(h_eoff.data(), w.eoff, (size_t)(Vtot + 1) * 8);
        std::vector<Edge> h_edge((size_t)E);
        std::vector<int32_t> h_src((size_t)E);
        if (E > 0) {
            bk.d2h(h_edge.data(), w.edge, (size_t)E * sizeof(Edge));
            bk.d2h(h_src.data(), w.e_src, (size_t)E * 4);
        }
        std::vector<D4> h_d((size_t)Vtot);
        std::vector<int32_t> h_best((size_t)Vtot), h_order((size_t)Vtot);
        bk.d2h(h_d.data(), w.d, (size_t)Vtot * sizeof(D4));
        bk.d2h(h_best.data(), w.best, (size_t)Vtot * 4);
        bk.d2h(h_order.data(), w.order, (size_t)Vtot * 4);
        std::vector<int64_t> h_anom((size_t)C);
        bk.d2h(h_anom.data(), w.anom_dis, (size_t)C * 8);
        g->vtx_off = host_n<int64_t>(C + 1);
        g->edge_off = host_n<int64_t>(C + 1);
        g->walk_off = host_n<int64_t>(C + 1);
        g->anom_dis = host_n<int64_t>(C);
        int64_t tv = 0, te = 0, tw = 0;
        for (int64_t c = 0; c < C; c++) {
            if (h_status[(size_t)c] == 0) {
                tv += h_voff[(size_t)c + 1] - h_voff[(size_t)c];
                te += h_eoff[(size_t)h_voff[(size_t)c + 1]] - h_eoff[(size_t)h_voff[(size_t)c]];
                tw += h_nwalk[(size_t)c];
            }
            g->vtx_off[c + 1] = tv;
            g->edge_off[c + 1] = te;
            g->walk_off[c + 1] = tw;
        }
        g->e_src = host_n<int32_t>(te);
        g->e_dst = host_n<int32_t>(te);
        g->e_qry = host_n<int64_t>(te);
        g->e_ref = host_n<int64_t>(te);
        g->e_anom = host_n<int32_t>(te);
        g->e_qnz = host_n<int32_t>(te);
        g->e_qtot = host_n<int32_t>(te);
        g->d_reach = host_n<uint8_t>(tv);
        g->d_sum = host_n<int64_t>(tv);
        g->d_anom = host_n<int32_t>(tv);
        g->d_qnz = host_n<int32_t>(tv);
        g->d_qtot = host_n<int32_t>(tv);
        g->best = host_n<int32_t>(tv);
        g->order = host_n<int32_t>(tv);
        g->w_sum = host_n<int64_t>(tw);
        g->w_anom = host_n<int32_t>(tw);
        g->w_qnz = host_n<int32_t>(tw);
        g->w_qtot = host_n<int32_t>(tw);
        std::vector<D4> h_w;
        for (int64_t c = 0; c < C; c++) {
            if (h_status[(size_t)c] != 0) {
                g->anom_dis[c] = -1;
                continue;
            }
            g->anom_dis[c] = h_anom[(size_t)c];
            int64_t v0 = h_voff[(size_t)c], nv = h_voff[(size_t)c + 1] - v0;
            int64_t e0 = h_eoff[(size_t)v0], ne = h_eoff[(size_t)(v0 + nv)] - e0;
            int64_t go = g->edge_off[c], gvv = g->vtx_off[c], gw = g->walk_off[c];
            for (int64_t k = 0; k < ne; k++) {
                const Edge &e = h_edge[(size_t)(e0 + k)];
                g->e_src[go + k] = h_src[(size_t)(e0 + k)];
                g->e_dst[go + k] = e_dst(e);
                g->e_qry[go + k] = e.qry;
                g->e_ref[go + k] = e.ref;
                g->e_anom[go + k] = e_anom(e);
                g->e_qnz[go + k] = e_nz(e);
                g->e_qtot[go + k] = e_tot(e);
            }
            for (int64_t v = 0; v < nv; v++) {
                const D4 &x = h_d[(size_t)(v0 + v)];
                g->d_reach[gvv + v] = x.aux ? 1 : 0;
                g->d_sum[gvv + v] = x.aux ? x.sum : 0;
                g->d_anom[gvv + v] = x.aux ? x.anom : 0;
                g->d_qnz[gvv + v] = x.aux ? x.nz : 0;
                g->d_qtot[gvv + v] = x.aux ? x.tot : 0;
                g->best[gvv + v] = h_best[(size_t)(v0 + v)];
                g->order[gvv + v] = h_order[(size_t)(v0 + v)];
            }
            int64_t nw = h_nwalk[(size_t)c];
            h_w.resize((size_t)nw);
            if (nw > 0) bk.d2h(h_w.data(), w.wdist + c * (int64_t)w.K, (size_t)nw * sizeof(D4));
            for (int64_t k = 0; k < nw; k++) {
                g->w_sum[gw + k] = h_w[(size_t)k].sum;
                g->w_anom[gw + k] = h_w[(size_t)k].anom;
                g->w_qnz[gw + k] = h_w[(size_t)k].nz;
                g->w_qtot[gw + k] = h_w[(size_t)k].tot;
            }
        }
    }
};

}  // namespace aa
