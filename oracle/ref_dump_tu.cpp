// oracle/ref_dump_tu.cpp — TEST INFRASTRUCTURE ONLY.
//
// Observation hooks around the UNMODIFIED reference solver.  `graph`, `d`, `best`, the walk list and
// the recovered walks are locals/privates of solve_ctg_read (paf_data.cpp:701-730, 1589-1640), so
// the only way to pin "graph edges, path weights and walk order" without editing reference sources
// is to stand in front of the two generic templates it instantiates:
//   * kShortestWalksSolver (k_shortest_walks.hpp:31-291)  -> wrapped, forwards every call
//   * k_weighted_bfs       (k_weighted_bfs.hpp:15-37)     -> wrapped, forwards
// The real headers are included inside namespace ref_impl (their include guards then keep
// paf_data.cpp from re-including them), and paf_data.cpp is compiled from where it lies.
//
// Dump grammar (one record per line, per contig, in call order):
//   C <contig> <n>                                  (written by ref_driver.cpp)
//   G <V> <E>                                       graph handed to the solver
//   E <u> <v> <qry> <ref> <anom> <qnz> <qtot>       out-edges in adjacency order (OC3)
//   A <anom_dis[dest]>
//   D <v> <reach> <qry> <ref> <anom> <qnz> <qtot> <best>
//   K <count>  then  k <i> <qry> <ref> <anom> <qnz> <qtot>
//   O <order...>                                    forward Kahn order
//   W <k> <len> <u0> <v0> <u1> <v1> ...             recovered walk (edge list)
#include <algorithm>
#include <cassert>
#include <cinttypes>
#include <cstdio>
#include <deque>
#include <queue>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "paf_data.hpp"
#include "graph_operations.hpp"
#include "leftist_heap.hpp"

extern FILE *g_ref_dump_file;

namespace ref_impl {
#define private public
#include "k_shortest_walks.hpp"
#undef private
#include "k_weighted_bfs.hpp"
}  // namespace ref_impl

template <typename GraphT>
void k_weighted_bfs(GraphT &graph, int64_t src, int64_t lim, std::vector<int64_t> &dist, std::vector<int64_t> &pre) {
    ref_impl::k_weighted_bfs(graph, src, lim, dist, pre);
    if (g_ref_dump_file) std::fprintf(g_ref_dump_file, "A %" PRId64 "\n", dist[graph.size() - 1]);
}

template <typename Distance, typename WeightedGraph>
struct kShortestWalksSolver {
    ref_impl::kShortestWalksSolver<Distance, WeightedGraph> impl;
    const WeightedGraph &g;
    Distance mx;
    explicit kShortestWalksSolver(const WeightedGraph &g_, Distance mx_, Distance id_, bool dag_ = false,
                                  bool neg_ = false)
        : impl(g_, mx_, id_, dag_, neg_), g(g_), mx(mx_) {
        FILE *f = g_ref_dump_file;
        if (!f) return;
        int64_t e = 0;
        for (auto &row : g) e += (int64_t)row.size();
        std::fprintf(f, "G %zu %" PRId64 "\n", g.size(), e);
        for (size_t u = 0; u < g.size(); u++)
            for (auto &[v, w] : g[u])
                std::fprintf(f, "E %zu %" PRId64 " %" PRId64 " %" PRId64 " %" PRId64 " %" PRId64 " %" PRId64 "\n", u, v,
                             w.qry_score, w.ref_score, w.anom, w.qul_nonzero, w.qul_total);
    }
    std::vector<int64_t> topology_sort(const WeightedGraph &g_) {
        auto r = impl.topology_sort(g_);
        if (FILE *f = g_ref_dump_file) {
            std::fputs("O", f);
            for (auto v : r) std::fprintf(f, " %" PRId64, v);
            std::fputc('\n', f);
        }
        return r;
    }
    std::vector<Distance> k_shortest_walks(int64_t s, int64_t t, int64_t k) {
        auto r = impl.k_shortest_walks(s, t, k);
        if (FILE *f = g_ref_dump_file) {
            for (size_t v = 0; v < impl.d.size(); v++) {
                const auto &x = impl.d[v];
                bool reach = !(x == mx);
                std::fprintf(f, "D %zu %d %" PRId64 " %" PRId64 " %" PRId64 " %" PRId64 " %" PRId64 " %" PRId64 "\n", v,
                             reach ? 1 : 0, x.qry_score, x.ref_score, x.anom, x.qul_nonzero, x.qul_total,
                             impl.best[v]);
            }
            std::fprintf(f, "K %zu\n", r.size());
            for (size_t i = 0; i < r.size(); i++)
                std::fprintf(f, "k %zu %" PRId64 " %" PRId64 " %" PRId64 " %" PRId64 " %" PRId64 "\n", i, r[i].qry_score,
                             r[i].ref_score, r[i].anom, r[i].qul_nonzero, r[i].qul_total);
        }
        return r;
    }
    std::vector<std::tuple<int64_t, int64_t, Distance>> kth_shortest_walk_recover(int64_t s, int64_t t, int64_t k,
                                                                                    bool call_k_paths = false) {
        auto r = impl.kth_shortest_walk_recover(s, t, k, call_k_paths);
        if (FILE *f = g_ref_dump_file) {
            std::fprintf(f, "W %" PRId64 " %zu", k, r.size());
            for (auto &[u, v, w] : r) std::fprintf(f, " %" PRId64 " %" PRId64, u, v);
            std::fputc('\n', f);
        }
        return r;
    }
};

#include "paf_data.cpp"
