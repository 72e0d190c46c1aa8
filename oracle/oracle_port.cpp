// oracle/oracle_port.cpp — TEST INFRASTRUCTURE ONLY.  A CPU restatement of the alignasm hot path,
// written from scratch (no reference code), used as the checker for the CUDA product.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it; the product never does.
//
// PARITY PINNED: checked against the reference's own solve_ctg_read compiled unmodified
// (oracle/_ref, see oracle/Makefile) — outputs byte-identical and graph / d / best / walk lists equal
// to the hook dump (tests/test_oracle_vs_ref.py, tests/golden/).  The reference ships no tests or
// golden vectors of its own.
//
// Each function cites the reference lines it follows.  Tie order is the canonical one of
// leftist-heap nodes compare by allocation order where the reference compares raw
// pointers (k_shortest_walks.hpp:231-247).
#include "../include/alignasm_b200.h"

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <queue>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

// ---- PafDistance (paf_data.hpp:121-189) -------------------------------------------------------
struct Dist {
    int64_t qry, ref, anom, qnz, qtot;
};
const Dist DMAX = {-1, -1, -1, -1, 0};  // paf_data.hpp:135-137
const Dist DZERO = {0, 0, 0, 0, 0};

inline int64_t den(int64_t t) { return t ? t : 1; }
inline bool d_eq(const Dist &a, const Dist &b) {  // paf_data.hpp:163-168
    return a.qry == b.qry && a.ref == b.ref && a.anom == b.anom && a.qnz * den(b.qtot) == b.qnz * den(a.qtot);
}
inline bool d_less(const Dist &a, const Dist &b, bool qry_mode) {  // paf_data.hpp:142-159
    if (d_eq(a, DMAX)) return false;
    if (d_eq(b, DMAX)) return true;
    if (!qry_mode) {
        if (a.qry + a.ref != b.qry + b.ref) return a.qry + a.ref < b.qry + b.ref;
    } else {
        if (a.qry != b.qry) return a.qry < b.qry;
        if (a.ref != b.ref) return a.ref < b.ref;
    }
    if (a.anom != b.anom) return a.anom < b.anom;
    return a.qnz * den(b.qtot) > b.qnz * den(a.qtot);
}
inline Dist d_add(const Dist &a, const Dist &b) {
    return {a.qry + b.qry, a.ref + b.ref, a.anom + b.anom, a.qnz + b.qnz, a.qtot + b.qtot};
}
inline Dist d_sub(const Dist &a, const Dist &b) {
    return {a.qry - b.qry, a.ref - b.ref, a.anom - b.anom, a.qnz - b.qnz, a.qtot - b.qtot};
}

struct Run {
    int64_t ql, qr, rl;
};
struct Blk {
    int64_t qs, qe, rs, re, qtot;
    int32_t chr, orig;
    uint8_t fwd, mapq;
    const Run *runs;
    int64_t nrun;
};
struct Row {
    int32_t ctg_index;
    int64_t qs, qe, rs, re;
    uint8_t is_alt;
};
struct Edge {
    int32_t to;
    Dist w;
};
struct PairRec {
    int32_t i, j;
    int64_t pe_q, pe_r;  // edited_loc_pre_end[i][j]
    int64_t st_q, st_r;  // edited_loc_str[i][j]
};

constexpr int64_t SV_BASELINE = 1000000, SV_TRANS = 2000, SV_INV = 500, FRONT_END = 2, REF_NEG = 2;  // paf_data.hpp:21-29

// ---- persistent leftist heap (leftist_heap.hpp:17-41), index-addressed -------------------------
struct HNode {
    int32_t rank;
    Dist key;
    int32_t u, v;  // value = edge (u, v)
    int32_t left, right;
};

struct Solver {
    bool nsl;
    int64_t n = 0;
    std::vector<Blk> S;  // sorted blocks
    std::vector<int32_t> part_idx, part_l, part_r;
    std::vector<PairRec> pairs;
    std::vector<int32_t> pair_begin;  // pairs of i are [pair_begin[i], pair_begin[i+1])
    int32_t V = 0, src = 0, dest = 0;
    std::vector<std::vector<Edge>> g;
    int64_t anom_dis_dest = -1;
    // k-walk state
    std::vector<Dist> d;
    std::vector<int32_t> best;
    std::vector<HNode> heap;
    std::vector<int32_t> h;
    std::vector<Dist> distances;
    std::vector<int32_t> nodes, prev_node, path_last;
    // select state
    std::vector<int32_t> sorted_vertices, order;
    std::vector<uint8_t> seen_block;  // not_alt_vertex_map, keyed by ctg_index

    int32_t pair_id(int32_t i, int32_t j) const {  // index_of_vtx[i][j] for i<j (paf_data.cpp:282,371)
        for (int32_t p = pair_begin[i]; p < pair_begin[i + 1]; p++)
            if (pairs[p].j == j) return (int32_t)n + p;
        return -1;
    }
    bool contains(int32_t a, int32_t b) const { return S[a].qs <= S[b].qs && S[b].qe <= S[a].qe; }  // paf_data.hpp:74-77
    bool partial(int32_t a, int32_t b) const {                                                       // paf_data.hpp:78-86
        if (S[a].qs < S[b].qs) return S[b].qs <= S[a].qe && S[a].qe < S[b].qe;
        if (S[b].qs < S[a].qs) return S[a].qs <= S[b].qe && S[b].qe < S[a].qe;
        return false;
    }

    // ---- parts (paf_data.cpp:249-261) ----
    void make_parts() {
        part_idx.assign(n, 0);
        part_l.assign(n, 0);
        part_r.assign(n, 0);
        int64_t part_end = -1;
        int32_t cur = -1, start = 0;
        for (int32_t i = 0; i < n; i++) {
            if (part_end < S[i].qs) {
                for (int32_t k = start; k < i; k++) part_r[k] = i;
                start = i;
                cur++;
            }
            part_idx[i] = cur;
            part_l[i] = start;
            part_end = std::max(part_end, S[i].qe);
        }
        for (int32_t k = start; k < n; k++) part_r[k] = (int32_t)n;
    }

    // ---- cut point of a partially-overlapping pair (paf_data.cpp:302-376) ----
    bool find_cut(int32_t i, int32_t j, PairRec &pr) const {
        const Blk &a = S[i], &b = S[j];
        int64_t step_b = b.fwd ? 1 : -1, step_a = a.fwd ? 1 : -1;
        int64_t min_gap = -1, gi = -1, gj = -1;
        int64_t pi = 0, pj = 0;
        while (pi < a.nrun && pj < b.nrun) {
            int64_t li = a.runs[pi].ql, ri = a.runs[pi].qr;
            int64_t lj = b.runs[pj].ql, rj = b.runs[pj].qr;
            if (li == lj) {
                if (lj == rj) {
                    pj++;
                    continue;
                }
                pr.pe_q = li;
                pr.pe_r = a.runs[pi].rl;
                pr.st_q = lj + 1;
                pr.st_r = b.runs[pj].rl + step_b;
                return true;
            }
            if (li < lj) {
                if (lj <= ri + 1) {
                    pr.pe_q = lj - 1;
                    pr.pe_r = a.runs[pi].rl + ((lj - 1) - li) * step_a;
                    pr.st_q = lj;
                    pr.st_r = b.runs[pj].rl;
                    return true;
                }
                int64_t gap = lj - (ri + 1);
                if (min_gap == -1 || gap < min_gap) {
                    min_gap = gap;
                    gi = pi;
                    gj = pj;
                }
                pi++;
            } else {
                if (li <= rj - 1) {
                    pr.pe_q = li;
                    pr.pe_r = a.runs[pi].rl;
                    pr.st_q = li + 1;
                    pr.st_r = b.runs[pj].rl + (li + 1 - lj) * step_b;
                    return true;
                }
                pj++;
            }
        }
        if (min_gap == -1) return false;  // release build drops the pair silently (paf_data.cpp:373-375)
        pr.pe_q = a.runs[gi].qr;
        pr.pe_r = a.runs[gi].rl + (a.runs[gi].qr - a.runs[gi].ql) * step_a;
        pr.st_q = b.runs[gj].ql;
        pr.st_r = b.runs[gj].rl;
        return true;
    }
    void make_pairs() {  // paf_data.cpp:294-378
        pair_begin.assign(n + 1, 0);
        for (int32_t i = 0; i < n; i++) {
            pair_begin[i] = (int32_t)pairs.size();
            for (int32_t j = i + 1; j < n; j++) {
                if (S[i].qe < S[j].qs) break;
                if (!partial(i, j)) continue;
                PairRec pr{i, j, 0, 0, 0, 0};
                if (find_cut(i, j, pr)) pairs.push_back(pr);
            }
        }
        pair_begin[n] = (int32_t)pairs.size();
    }

    // ---- vertex view (Internal_Vertex, paf_data.cpp:392-411) ----
    struct VV {
        int32_t pre, cur;
        int64_t qs, qe, rs, re;
    };
    VV single(int32_t i) const { return {i, i, S[i].qs, S[i].qe, S[i].rs, S[i].re}; }
    VV pairv(int32_t p) const {
        const PairRec &r = pairs[p];
        return {r.i, r.j, r.st_q, S[r.j].qe, r.st_r, S[r.j].re};
    }
    // get_score (paf_data.cpp:449-521); rht_pair >= 0 when the right vertex is a pair vertex
    Dist score(VV l, const VV &r, int32_t rht_pair) const {
        if (rht_pair >= 0) {
            l.qe = pairs[rht_pair].pe_q;
            l.re = pairs[rht_pair].pe_r;
        }
        auto ref_abs = [](int64_t x) { return x < 0 ? -x * REF_NEG : x; };
        Dist w = DZERO;
        int64_t qry_diff = r.qs - l.qe - 1, ref_diff = 0;
        const Blk &a = S[l.cur], &b = S[r.cur];
        if (a.chr == b.chr && a.fwd == b.fwd) {
            int64_t gap = a.fwd ? r.rs - (l.re + 1) : l.re - (r.rs + 1);
            ref_diff += ref_abs(gap);
            if (ref_diff > SV_BASELINE) {
                w.anom += 1;
                ref_diff = SV_BASELINE;
            }
        } else if (a.chr == b.chr) {
            w.anom += 1;
            ref_diff += SV_INV;
            if (a.fwd) ref_diff += ref_abs(r.re - (l.re + 1));
            else ref_diff += ref_abs(r.rs - (l.rs + 1));
            if (ref_diff > SV_BASELINE) {
                w.anom += 1;
                ref_diff = SV_BASELINE;
            }
        } else {
            w.anom += 1;
            ref_diff = SV_TRANS;
        }
        w.qry = qry_diff;
        w.ref = ref_diff;
        if (b.mapq) w.qnz += 1;
        w.qtot += 1;
        return w;
    }

    // ---- make_Graph, restated per source vertex in the reference's per-vertex order (OC3) ----
    void add(int32_t u, int32_t v, const Dist &w) { g[u].push_back({v, w}); }
    void make_graph() {  // paf_data.cpp:531-696
        int32_t P = (int32_t)pairs.size();
        V = (int32_t)n + P + 2;
        src = (int32_t)n + P;
        dest = src + 1;
        g.assign(V, {});
        const int64_t I64MAX = INT64_MAX;
        // src -> first part (paf_data.cpp:540-563)
        {
            int64_t min_qe = I64MAX;
            for (int32_t i = 0; i < part_r[0]; i++) {
                if (nsl) {
                    if (min_qe < S[i].qs) break;
                    min_qe = std::min(min_qe, S[i].qe);
                }
                Dist w = DZERO;
                w.qry = S[i].qs * FRONT_END;
                if (S[i].mapq) w.qnz = 1;
                w.qtot = 1;
                add(src, i, w);
            }
        }
        int32_t last_l = part_l[n - 1];
        int64_t max_qs = S[n - 1].qs;
        auto dest_weight = [&](int32_t i) {
            Dist w = DZERO;
            w.qry = (S[i].qtot - S[i].qe - 1) * FRONT_END;
            return w;
        };
        for (int32_t i = 0; i < n; i++) {
            int32_t r = part_r[i];
            bool last = part_l[i] == last_l;
            VV vi = single(i);
            // (i,i) -> dest (paf_data.cpp:565-585)
            if (last && !(nsl && S[i].qe < max_qs)) add(i, dest, dest_weight(i));
            // inside the part (paf_data.cpp:598-651), (i,i) rows only
            int64_t min_after = I64MAX;
            for (int32_t j = i + 1; j < r; j++) {
                if (contains(i, j)) continue;
                if (nsl) {
                    if (min_after < S[j].qs) break;
                    if (S[i].qe < S[j].qs) min_after = std::min(min_after, S[j].qe);
                }
                if (S[i].qe < S[j].qs) {
                    add(i, j, score(vi, single(j), -1));
                } else {
                    int32_t pid = pair_id(i, j);
                    if (pid >= 0 && vi.qs < pairs[pid - n].st_q) add(i, pid, score(vi, pairv(pid - (int32_t)n), pid - (int32_t)n));
                }
            }
            // to the next part (paf_data.cpp:653-673)
            if (r < n) {
                int32_t r2 = part_r[r];
                int64_t mn = I64MAX;
                for (int32_t k = r; k < r2; k++) {
                    if (nsl) {
                        if (mn < S[k].qs) break;
                        if (S[i].qe < S[k].qs) mn = std::min(mn, S[k].qe);
                    }
                    add(i, k, score(vi, single(k), -1));
                }
            }
        }
        for (int32_t p = 0; p < P; p++) {
            int32_t i = pairs[p].i, j = pairs[p].j, u = (int32_t)n + p;
            (void)i;
            int32_t r = part_r[j];
            bool last = part_l[j] == last_l;
            VV vp = pairv(p);
            // (i,j) -> dest (paf_data.cpp:586-593)
            if (last && !(nsl && S[j].qe < max_qs)) add(u, dest, dest_weight(j));
            // (i,j) -> (k,k) / (j,k) inside the part (paf_data.cpp:627-646)
            int64_t min_after = I64MAX;
            for (int32_t k = j + 1; k < r; k++) {
                if (nsl) {
                    if (min_after < S[k].qs) break;
                    if (S[j].qe < S[k].qs) min_after = std::min(min_after, S[k].qe);
                }
                if (S[j].qe < S[k].qs) add(u, k, score(vp, single(k), -1));  // linkable: paf_data.cpp:440-442
                int32_t pid = pair_id(j, k);
                if (pid >= 0 && vp.qs < pairs[pid - n].st_q)  // linkable: paf_data.cpp:433-436
                    add(u, pid, score(vp, pairv(pid - (int32_t)n), pid - (int32_t)n));
            }
            // (i,j) -> next part (paf_data.cpp:674-692)
            if (r < n) {
                int32_t r2 = part_r[r];
                int64_t mn = I64MAX;
                for (int32_t k = r; k < r2; k++) {
                    if (nsl) {
                        if (mn < S[k].qs) break;
                        if (S[j].qe < S[k].qs) mn = std::min(mn, S[k].qe);
                    }
                    add(u, k, score(vp, single(k), -1));
                }
            }
        }
    }

    // ---- anom_dis[dest] (paf_data.cpp:705-713; k_weighted_bfs.hpp:15-37) ----
    // Dial's BFS with weights in {0,1,2} computes the plain shortest anom distance src -> dest.
    void anom_bfs() {
        std::vector<int64_t> dist(V, -1);
        std::vector<std::vector<int32_t>> buckets(3);
        dist[src] = 0;
        buckets[0].push_back(src);
        for (int64_t dd = 0, maxd = 0; dd <= maxd; dd++) {
            auto &q = buckets[dd % 3];
            while (!q.empty()) {
                int32_t cur = q.back();
                q.pop_back();
                if (dist[cur] != dd) continue;
                for (auto &e : g[cur]) {
                    int64_t nd = dd + e.w.anom;
                    if (dist[e.to] != -1 && dist[e.to] <= nd) continue;
                    dist[e.to] = nd;
                    buckets[nd % 3].push_back(e.to);
                    maxd = std::max(maxd, nd);
                }
            }
        }
        anom_dis_dest = dist[dest];
    }

    // ---- Kahn FIFO order (k_shortest_walks.hpp:132-156) ----
    static std::vector<int32_t> kahn(const std::vector<std::vector<Edge>> &gr) {
        int32_t nv = (int32_t)gr.size();
        std::vector<int32_t> indeg(nv, 0), q;
        for (auto &row : gr)
            for (auto &e : row) indeg[e.to]++;
        q.reserve(nv);
        for (int32_t u = 0; u < nv; u++)
            if (!indeg[u]) q.push_back(u);
        for (size_t head = 0; head < q.size(); head++)
            for (auto &e : gr[q[head]])
                if (--indeg[e.to] == 0) q.push_back(e.to);
        return q;
    }

    int32_t heap_insert(int32_t a, const Dist &k, int32_t u, int32_t v) {  // leftist_heap.hpp:29-40
        // walk down the right spine while a->key < k, then rebuild bottom-up (path copying)
        std::vector<int32_t> spine;
        while (a >= 0 && d_less(heap[a].key, k, false)) {
            spine.push_back(a);
            a = heap[a].right;
        }
        heap.push_back({1, k, u, v, a, -1});
        int32_t r = (int32_t)heap.size() - 1;
        for (size_t s = spine.size(); s-- > 0;) {
            const HNode &o = heap[spine[s]];
            int32_t l = o.left, rr = r;
            if (l < 0 || heap[l].rank < heap[rr].rank) std::swap(l, rr);
            HNode nn{rr >= 0 ? heap[rr].rank + 1 : 0, o.key, o.u, o.v, l, rr};
            heap.push_back(nn);
            r = (int32_t)heap.size() - 1;
        }
        return r;
    }

    bool k_walks(int64_t k) {  // k_shortest_walks.hpp:179-251
        std::vector<std::vector<Edge>> grev(V);
        for (int32_t u = 0; u < V; u++)
            for (auto &e : g[u]) grev[e.to].push_back({u, e.w});
        // shortest_path_dag on the reverse graph (k_shortest_walks.hpp:160-175)
        d.assign(V, DMAX);
        best.assign(V, -1);
        d[dest] = DZERO;
        for (int32_t v : kahn(grev)) {
            if (d_eq(d[v], DMAX)) continue;
            for (auto &e : grev[v]) {
                Dist cand = d_add(d[v], e.w);
                if (d_less(cand, d[e.to], false)) {
                    d[e.to] = cand;
                    best[e.to] = v;
                }
            }
        }
        if (d_eq(d[src], DMAX)) return false;
        std::vector<std::vector<int32_t>> tree(V);
        for (int32_t u = 0; u < V; u++)
            if (best[u] != -1) tree[best[u]].push_back(u);
        h.assign(V, -1);
        {
            std::vector<int32_t> q{dest};
            for (size_t head = 0; head < q.size(); head++) {
                int32_t u = q[head];
                bool seen_p = false;
                for (auto &e : g[u]) {
                    if (d_eq(d[e.to], DMAX)) continue;
                    Dist c = d_sub(d_add(e.w, d[e.to]), d[u]);
                    if (!seen_p && e.to == best[u] && d_eq(c, DZERO)) {
                        seen_p = true;
                        continue;
                    }
                    h[u] = heap_insert(h[u], c, u, e.to);
                }
                for (int32_t p : tree[u]) {
                    h[p] = h[u];
                    q.push_back(p);
                }
            }
        }
        distances.assign(1, d[src]);
        path_last.assign(1, -1);
        nodes.clear();
        prev_node.clear();
        if (h[src] < 0) return true;
        struct Ent {
            Dist dist;
            int32_t node, idx;
        };
        auto after = [](const Ent &a, const Ent &b) {  // std::greater on tuple<Distance, heap_t*, int64_t>
            if (d_less(b.dist, a.dist, false)) return true;
            if (d_less(a.dist, b.dist, false)) return false;
            if (a.node != b.node) return a.node > b.node;  // allocation order stands in for the pointer
            return a.idx > b.idx;
        };
        std::priority_queue<Ent, std::vector<Ent>, decltype(after)> pq(after);
        auto emplace = [&](const Dist &dd, int32_t hn, int32_t pre) {
            pq.push({dd, hn, (int32_t)nodes.size()});
            nodes.push_back(hn);
            prev_node.push_back(pre);
        };
        emplace(d_add(d[src], heap[h[src]].key), h[src], -1);
        while (!pq.empty() && (int64_t)distances.size() < k) {
            Ent t = pq.top();
            pq.pop();
            distances.push_back(t.dist);
            path_last.push_back(t.idx);
            const HNode ch = heap[t.node];
            if (h[ch.v] >= 0) emplace(d_add(t.dist, heap[h[ch.v]].key), h[ch.v], t.idx);
            if (ch.left >= 0) emplace(d_sub(d_add(t.dist, heap[ch.left].key), ch.key), ch.left, prev_node[t.idx]);
            if (ch.right >= 0) emplace(d_sub(d_add(t.dist, heap[ch.right].key), ch.key), ch.right, prev_node[t.idx]);
        }
        return true;
    }

    // walk as a vertex sequence src ... dest (k_shortest_walks.hpp:254-290)
    std::vector<int32_t> recover(int64_t k) const {
        std::vector<std::pair<int32_t, int32_t>> side;
        for (int32_t cur = path_last[k]; cur != -1; cur = prev_node[cur]) side.push_back({heap[nodes[cur]].u, heap[nodes[cur]].v});
        std::reverse(side.begin(), side.end());
        std::vector<int32_t> walk{src};
        size_t idx = 0;
        int32_t cur = src;
        while (cur != dest || idx < side.size()) {
            if (idx < side.size() && cur == side[idx].first) cur = side[idx++].second;
            else cur = best[cur];
            walk.push_back(cur);
        }
        return walk;
    }

    void vtx(int32_t v, int32_t &x, int32_t &y) const {  // index_to_vtx
        if (v < n) x = y = v;
        else {
            x = pairs[v - n].i;
            y = pairs[v - n].j;
        }
    }

    // internal_shortest_path_recover (paf_data.cpp:750-792): QRY_SCORE-mode DP over the forward
    // topological range [order[s], order[t]); returns the vertex sequence s ... t (empty when s == t)
    std::vector<int32_t> sub_path(int32_t s, int32_t t, bool wl_flag, int32_t wl) const {
        std::vector<int32_t> res;
        if (s == t) return res;
        std::unordered_map<int32_t, Dist> dist;
        std::unordered_map<int32_t, int32_t> pre;
        dist[s] = DZERO;
        pre[s] = -1;
        for (int32_t i = order[s]; i < order[t]; i++) {
            int32_t u = sorted_vertices[i];
            auto it = dist.find(u);
            if (it == dist.end()) continue;
            Dist cur = it->second;
            for (auto &e : g[u]) {
                if (wl_flag && e.to == t) {
                    if (u == src || u == dest) continue;
                    int32_t x, y;
                    vtx(u, x, y);
                    if (y != wl) continue;
                }
                Dist nx = d_add(cur, e.w);
                auto jt = dist.find(e.to);
                if (jt == dist.end() || d_less(nx, jt->second, true)) {
                    dist[e.to] = nx;
                    pre[e.to] = u;
                }
            }
        }
        for (int32_t last = t; last != -1; last = pre.at(last)) res.push_back(last);
        std::reverse(res.begin(), res.end());
        return res;
    }

    // upgrade_edge_path_with_alt_path (paf_data.cpp:795-921) on vertex sequences.
    // `walk` = src, v1, ..., vm, dest.  The result is again a vertex sequence src ... dest.
    std::vector<int32_t> upgrade(const std::vector<int32_t> &walk) const {
        std::vector<int32_t> up{src};  // `edge_path` as its vertex sequence; back() == get<1>(edge_path.back())
        auto splice = [&](const std::vector<int32_t> &sp, bool drop_last) {
            // sp starts at up.back(); append the rest (optionally without its final vertex)
            size_t end = sp.size() - (drop_last ? 1 : 0);
            for (size_t k = 1; k < end; k++) up.push_back(sp[k]);
        };
        size_t m = walk.size();
        for (size_t e = 0; e + 1 < m; e++) {  // edge (walk[e], walk[e+1])
            int32_t u = walk[e], v = walk[e + 1];
            if (v == dest) {  // paf_data.cpp:845-858 (u != src here: a walk never is src -> dest)
                int32_t cs = up.back();
                auto sp = sub_path(cs, v, false, -1);
                if (!sp.empty()) splice(sp, false);
                continue;
            }
            int32_t x, y;
            vtx(v, x, y);
            if (u != src && x != y) {  // paf_data.cpp:866-873
                up.push_back(v);
                continue;
            }
            // v is a single vertex (y,y); look at the following edge (v, nv)
            int32_t cs = up.back();
            int32_t nv = walk[e + 2];
            int32_t nx = -1, ny = -1;
            if (nv != dest) vtx(nv, nx, ny);
            if (nv == dest || nx == ny) {  // paf_data.cpp:812-833, 879-899
                auto sp = sub_path(cs, nv, true, y);
                if (sp.empty()) up.push_back(v);
                else splice(sp, true);
            } else {  // nv is the pair (y, ny): paf_data.cpp:834-843, 900-909
                auto sp = sub_path(cs, nv, false, -1);
                if (sp.empty()) {
                    up.push_back(v);
                    up.push_back(nv);
                } else {
                    splice(sp, false);
                }
                e++;  // the edge (v, nv) is consumed
            }
        }
        return up;
    }

    // edge_path_to_paf_path (paf_data.cpp:1489-1568)
    std::vector<Row> to_rows(const std::vector<int32_t> &walk) {
        for (size_t k = 1; k < walk.size(); k++) {
            if (walk[k] == dest) continue;
            int32_t x, y;
            vtx(walk[k], x, y);
            seen_block[S[x].orig] = 1;
            seen_block[S[y].orig] = 1;
        }
        std::vector<int32_t> up = upgrade(walk);
        std::vector<Row> rows;
        auto full = [&](int32_t b) { return Row{S[b].orig, S[b].qs, S[b].qe, S[b].rs, S[b].re, 0}; };
        for (size_t k = 1; k + 1 < up.size(); k++) {
            int32_t v = up[k], x, y;
            vtx(v, x, y);
            rows.push_back(full(y));
            if (x != y) {  // arriving at a pair vertex trims both sides (paf_data.cpp:1523-1531, 1546-1553)
                const PairRec &pr = pairs[v - n];
                Row &px = rows[rows.size() - 2];
                px.qe = pr.pe_q;
                px.re = pr.pe_r;
                Row &py = rows[rows.size() - 1];
                py.qs = pr.st_q;
                py.rs = pr.st_r;
            }
        }
        for (auto &r : rows) r.is_alt = seen_block[r.ctg_index] ? 0 : 1;  // paf_data.cpp:1560-1566
        return rows;
    }
    static int64_t coverage(const std::vector<Row> &rows) {  // paf_data.cpp:1571-1579
        int64_t t = 0;
        for (auto &r : rows) t += (r.qe - r.qs) + std::llabs(r.re - r.rs);
        return t;
    }
};

struct CtgOut {
    std::vector<Row> out, alt;
    std::vector<std::vector<Row>> all;
    int64_t n_pair = 0, n_vtx = 0, n_edge = 0, n_heap = 0, n_walk = 0, n_task = 0;
    bool ok = true;
    // debug
    Solver *keep = nullptr;
};

void solve_contig(const aa_batch *b, int64_t c, const aa_opts *opt, int32_t *sorted_index, CtgOut &o, bool keep_dbg) {
    int64_t lo = b->ctg_off[c], hi = b->ctg_off[c + 1];
    int64_t n = hi - lo;
    std::vector<Run> runs;
    auto mk = [&](int64_t g, int32_t orig) {
        Blk k{b->qry_str[g], b->qry_end[g], b->ref_str[g], b->ref_end[g], b->qry_total[g], b->ref_chr[g], orig, b->aln_fwd[g], b->map_qul[g], nullptr, 0};
        return k;
    };
    if (n == 1) {  // paf_data.cpp:235-239
        sorted_index[lo] = 0;
        Blk k = mk(lo, 0);
        o.out.push_back({0, k.qs, k.qe, k.rs, k.re, 0});
        return;
    }
    Solver *sp = new Solver();
    Solver &s = *sp;
    s.nsl = opt && opt->non_skip_linkable;
    s.n = n;
    // std::sort with the reference comparator (paf_data.cpp:241; paf_data.hpp:69-73) — unstable, so the
    // same algorithm must see the same comparison outcomes
    std::vector<int32_t> perm(n);
    for (int64_t i = 0; i < n; i++) perm[i] = (int32_t)i;
    struct Key {
        int64_t qs, qe;
        int32_t idx;
        bool operator<(const Key &r) const { return qs != r.qs ? qs < r.qs : qe < r.qe; }
    };
    std::vector<Key> keys(n);
    for (int64_t i = 0; i < n; i++) keys[i] = {b->qry_str[lo + i], b->qry_end[lo + i], (int32_t)i};
    std::sort(keys.begin(), keys.end());
    // run storage
    int64_t nr = b->run_off[hi] - b->run_off[lo];
    runs.resize(nr);
    for (int64_t r = 0; r < nr; r++) {
        int64_t g = b->run_off[lo] + r;
        runs[r] = {b->run_ql[g], b->run_qr[g], b->run_rl[g]};
    }
    s.S.resize(n);
    for (int64_t i = 0; i < n; i++) {
        int32_t orig = keys[i].idx;
        sorted_index[lo + orig] = (int32_t)i;
        s.S[i] = mk(lo + orig, orig);
        s.S[i].runs = runs.data() + (b->run_off[lo + orig] - b->run_off[lo]);
        s.S[i].nrun = b->run_off[lo + orig + 1] - b->run_off[lo + orig];
    }
    s.make_parts();
    s.make_pairs();
    s.make_graph();
    s.anom_bfs();
    int64_t K = (opt && opt->max_walks > 0) ? opt->max_walks : 10000;
    if (!s.k_walks(K)) {
        o.ok = false;
        delete sp;
        return;
    }
    s.sorted_vertices = Solver::kahn(s.g);  // paf_data.cpp:742-746
    s.order.assign(s.V, 0);
    for (int32_t i = 0; i < s.V; i++) s.order[s.sorted_vertices[i]] = i;
    s.seen_block.assign(n, 0);

    o.n_pair = (int64_t)s.pairs.size();
    o.n_vtx = s.V;
    for (auto &row : s.g) o.n_edge += (int64_t)row.size();
    o.n_heap = (int64_t)s.heap.size();
    o.n_walk = (int64_t)s.distances.size();

    // ---- selection (paf_data.cpp:1585-1649) ----
    auto same = [](const Dist &a, const Dist &b) { return a.qry + a.ref == b.qry + b.ref && a.anom == b.anom; };
    const Dist mind = s.distances[0];
    auto rows0 = s.to_rows(s.recover(0));
    o.n_task++;
    int64_t max_cov = Solver::coverage(rows0);
    o.out = rows0;
    for (size_t idx = 1; idx < s.distances.size() && same(mind, s.distances[idx]); idx++) {
        auto rows = s.to_rows(s.recover((int64_t)idx));
        o.n_task++;
        int64_t cov = Solver::coverage(rows);
        if (cov > max_cov) {
            max_cov = cov;
            o.out = rows;
            o.all.clear();
        } else if (cov == max_cov) {
            o.all.push_back(rows);
        }
    }
    max_cov = -1;
    if (s.distances.size() >= 2 && mind.anom != s.anom_dis_dest) {
        int64_t ans_up = 0, ans_down = 0, ans_idx = -1;
        for (size_t i = 1; i < s.distances.size(); i++) {
            const Dist &dd = s.distances[i];
            if (dd.anom >= mind.anom) continue;
            int64_t up = (dd.qry + dd.ref) - (mind.qry + mind.ref);
            int64_t down = mind.anom - dd.anom;
            if (ans_idx == -1 || up * ans_down < down * ans_up) {
                ans_up = up;
                ans_down = down;
                ans_idx = (int64_t)i;
                auto rows = s.to_rows(s.recover((int64_t)i));
                o.n_task++;
                max_cov = Solver::coverage(rows);
                o.alt = rows;
            } else if (same(dd, s.distances[ans_idx])) {
                auto rows = s.to_rows(s.recover((int64_t)i));
                o.n_task++;
                int64_t cov = Solver::coverage(rows);
                if (cov > max_cov) {
                    max_cov = cov;
                    o.alt = rows;
                }
            }
        }
    }
    if (keep_dbg) o.keep = sp;
    else delete sp;
}

void rows_alloc(aa_rows &r, int64_t n) {
    r.n = n;
    size_t m = (size_t)(n ? n : 1);
    r.ctg_index = (int32_t *)std::malloc(m * 4);
    r.qry_str = (int64_t *)std::malloc(m * 8);
    r.qry_end = (int64_t *)std::malloc(m * 8);
    r.ref_str = (int64_t *)std::malloc(m * 8);
    r.ref_end = (int64_t *)std::malloc(m * 8);
    r.is_alt = (uint8_t *)std::malloc(m);
}
void rows_put(aa_rows &r, int64_t at, const Row &x) {
    r.ctg_index[at] = x.ctg_index;
    r.qry_str[at] = x.qs;
    r.qry_end[at] = x.qe;
    r.ref_str[at] = x.rs;
    r.ref_end[at] = x.re;
    r.is_alt[at] = x.is_alt;
}
void rows_free(aa_rows &r) {
    std::free(r.ctg_index);
    std::free(r.qry_str);
    std::free(r.qry_end);
    std::free(r.ref_str);
    std::free(r.ref_end);
    std::free(r.is_alt);
    std::memset(&r, 0, sizeof r);
}
template <class T>
T *alloc_n(int64_t n) {
    return (T *)std::calloc((size_t)(n ? n : 1), sizeof(T));
}

}  // namespace

extern "C" {

// Same contract as aa_solve (include/alignasm_b200.h) but on host cores.  threads <= 1: serial.
int oracle_solve(const aa_batch *b, const aa_opts *opt, aa_result *res, int threads) {
    if (!b || !res || b->n_ctg < 0) return AA_ERR_INVALID;
    std::memset(res, 0, sizeof *res);
    int64_t C = b->n_ctg;
    for (int64_t c = 0; c < C; c++)
        if (b->ctg_off[c + 1] <= b->ctg_off[c]) return AA_ERR_INVALID;
    std::vector<CtgOut> outs((size_t)C);
    res->sorted_index = alloc_n<int32_t>(b->n_blk);
    bool keep_dbg = opt && opt->keep_debug;
    bool want_all = opt && opt->want_all;
    std::atomic<int64_t> next{0};
    auto worker = [&]() {
        for (;;) {
            int64_t c = next.fetch_add(1);
            if (c >= C) break;
            solve_contig(b, c, opt, res->sorted_index, outs[(size_t)c], keep_dbg);
            if (!want_all) {
                outs[(size_t)c].all.clear();
                outs[(size_t)c].all.shrink_to_fit();
            }
        }
    };
    if (threads <= 1) worker();
    else {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; t++) pool.emplace_back(worker);
        for (auto &t : pool) t.join();
    }
    res->n_ctg = C;
    res->out_off = alloc_n<int64_t>(C + 1);
    res->alt_off = alloc_n<int64_t>(C + 1);
    res->all_path_off = alloc_n<int64_t>(C + 1);
    int64_t no = 0, na = 0, np = 0, nr = 0;
    bool ok = true;
    for (int64_t c = 0; c < C; c++) {
        CtgOut &o = outs[(size_t)c];
        ok = ok && o.ok;
        no += (int64_t)o.out.size();
        na += (int64_t)o.alt.size();
        np += (int64_t)o.all.size();
        for (auto &p : o.all) nr += (int64_t)p.size();
        res->out_off[c + 1] = no;
        res->alt_off[c + 1] = na;
        res->all_path_off[c + 1] = np;
        res->stats.n_pair += o.n_pair;
        res->stats.n_vtx += o.n_vtx;
        res->stats.n_edge += o.n_edge;
        res->stats.n_heap += o.n_heap;
        res->stats.n_walk += o.n_walk;
        res->stats.n_task += o.n_task;
    }
    res->stats.n_ctg = C;
    res->stats.n_blk = b->n_blk;
    res->stats.n_run = b->n_run;
    rows_alloc(res->out, no);
    rows_alloc(res->alt, na);
    rows_alloc(res->all, nr);
    res->all_row_off = alloc_n<int64_t>(np + 1);
    int64_t io = 0, ia = 0, ip = 0, ir = 0;
    for (int64_t c = 0; c < C; c++) {
        CtgOut &o = outs[(size_t)c];
        for (auto &r : o.out) rows_put(res->out, io++, r);
        for (auto &r : o.alt) rows_put(res->alt, ia++, r);
        for (auto &p : o.all) {
            for (auto &r : p) rows_put(res->all, ir++, r);
            res->all_row_off[++ip] = ir;
        }
    }
    if (keep_dbg) {
        aa_debug *g = (aa_debug *)std::calloc(1, sizeof(aa_debug));
        res->dbg = g;
        g->vtx_off = alloc_n<int64_t>(C + 1);
        g->edge_off = alloc_n<int64_t>(C + 1);
        g->walk_off = alloc_n<int64_t>(C + 1);
        g->anom_dis = alloc_n<int64_t>(C);
        int64_t tv = 0, te = 0, tw = 0;
        for (int64_t c = 0; c < C; c++) {
            Solver *s = outs[(size_t)c].keep;
            if (s) {
                tv += s->V;
                for (auto &row : s->g) te += (int64_t)row.size();
                tw += (int64_t)s->distances.size();
            }
            g->vtx_off[c + 1] = tv;
            g->edge_off[c + 1] = te;
            g->walk_off[c + 1] = tw;
        }
        g->e_src = alloc_n<int32_t>(te);
        g->e_dst = alloc_n<int32_t>(te);
        g->e_qry = alloc_n<int64_t>(te);
        g->e_ref = alloc_n<int64_t>(te);
        g->e_anom = alloc_n<int32_t>(te);
        g->e_qnz = alloc_n<int32_t>(te);
        g->e_qtot = alloc_n<int32_t>(te);
        g->d_reach = alloc_n<uint8_t>(tv);
        g->d_sum = alloc_n<int64_t>(tv);
        g->d_anom = alloc_n<int32_t>(tv);
        g->d_qnz = alloc_n<int32_t>(tv);
        g->d_qtot = alloc_n<int32_t>(tv);
        g->best = alloc_n<int32_t>(tv);
        g->order = alloc_n<int32_t>(tv);
        g->w_sum = alloc_n<int64_t>(tw);
        g->w_anom = alloc_n<int32_t>(tw);
        g->w_qnz = alloc_n<int32_t>(tw);
        g->w_qtot = alloc_n<int32_t>(tw);
        for (int64_t c = 0; c < C; c++) {
            Solver *s = outs[(size_t)c].keep;
            if (!s) {
                g->anom_dis[c] = -1;
                continue;
            }
            int64_t ev = g->edge_off[c], vv = g->vtx_off[c], wv = g->walk_off[c];
            for (int32_t u = 0; u < s->V; u++) {
                for (auto &e : s->g[u]) {
                    g->e_src[ev] = u;
                    g->e_dst[ev] = e.to;
                    g->e_qry[ev] = e.w.qry;
                    g->e_ref[ev] = e.w.ref;
                    g->e_anom[ev] = (int32_t)e.w.anom;
                    g->e_qnz[ev] = (int32_t)e.w.qnz;
                    g->e_qtot[ev] = (int32_t)e.w.qtot;
                    ev++;
                }
                bool reach = !d_eq(s->d[u], DMAX);
                g->d_reach[vv + u] = reach;
                g->d_sum[vv + u] = reach ? s->d[u].qry + s->d[u].ref : 0;
                g->d_anom[vv + u] = reach ? (int32_t)s->d[u].anom : 0;
                g->d_qnz[vv + u] = reach ? (int32_t)s->d[u].qnz : 0;
                g->d_qtot[vv + u] = reach ? (int32_t)s->d[u].qtot : 0;
                g->best[vv + u] = s->best[u];
                g->order[vv + u] = s->order[u];
            }
            for (size_t i = 0; i < s->distances.size(); i++) {
                const Dist &x = s->distances[i];
                g->w_sum[wv + (int64_t)i] = x.qry + x.ref;
                g->w_anom[wv + (int64_t)i] = (int32_t)x.anom;
                g->w_qnz[wv + (int64_t)i] = (int32_t)x.qnz;
                g->w_qtot[wv + (int64_t)i] = (int32_t)x.qtot;
            }
            g->anom_dis[c] = s->anom_dis_dest;
            delete s;
        }
    }
    return ok ? AA_OK : AA_ERR_UNSOLVABLE;
}

void oracle_result_free(aa_result *res) {
    if (!res) return;
    std::free(res->out_off);
    std::free(res->alt_off);
    std::free(res->all_path_off);
    std::free(res->all_row_off);
    std::free(res->sorted_index);
    rows_free(res->out);
    rows_free(res->alt);
    rows_free(res->all);
    if (aa_debug *g = res->dbg) {
        void *ptrs[] = {g->vtx_off, g->edge_off, g->walk_off, g->e_src, g->e_dst, g->e_qry, g->e_ref, g->e_anom, g->e_qnz,
                        g->e_qtot, g->d_reach, g->d_sum, g->d_anom, g->d_qnz, g->d_qtot, g->best, g->order, g->w_sum,
                        g->w_anom, g->w_qnz, g->w_qtot, g->anom_dis};
        for (void *p : ptrs) std::free(p);
        std::free(g);
    }
    std::memset(res, 0, sizeof *res);
}

}  // extern "C"
