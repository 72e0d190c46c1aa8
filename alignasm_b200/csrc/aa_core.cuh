// aa_core.cuh — data layout in HBM and the per-item device functions of the alignasm hot path.
//
// Everything here is a __host__ __device__ function over plain pointers (struct Ws), so that the
// kernels in aa_solve.cu are thin grids over items (blocks, candidate pairs, vertices, contigs, walk
// tasks), and so that tests/emul can run the very same functions on the CPU to check the logic
// against the reference without a GPU.  The product never runs them on the host.
//
// Reference map (all under reference src/):
//   parts                      paf_data.cpp:249-261
//   pair vertices + cut points paf_data.cpp:294-378
//   linkable / get_score       paf_data.cpp:422-521
//   make_Graph                 paf_data.cpp:531-696     (restated per source vertex, order OC3)
//   anom distance              paf_data.cpp:705-713, k_weighted_bfs.hpp:15-37
//   reverse relax (d, best)    k_shortest_walks.hpp:132-175, 180-184
//   sidetrack heaps            k_shortest_walks.hpp:191-215, leftist_heap.hpp:29-40
//   walk enumeration           k_shortest_walks.hpp:217-251
//   walk recovery              k_shortest_walks.hpp:254-290
//   gap filling DP / upgrade   paf_data.cpp:750-921
//   rows, flags, selection     paf_data.cpp:1489-1649
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#if defined(__CUDACC__)
#define AA_HD __host__ __device__ __forceinline__
#define AA_HDN __host__ __device__
#else
#define AA_HD inline
#define AA_HDN inline
#endif

namespace aa {

// ---- warp helpers: per-contig phases run one warp per contig.  On the host (tests/emul) a "warp" is one lane.
#if defined(__CUDA_ARCH__)
#define AA_LANES 32
AA_HD int aa_lane() { return (int)(threadIdx.x & 31); }
AA_HD void aa_syncwarp() { __syncwarp(); }
#else
#define AA_LANES 1
inline int aa_lane() { return 0; }
inline void aa_syncwarp() {}
#endif

// ---- constants (paf_data.hpp:21-29, paf_data.cpp:729) ------------------------------------------
constexpr int64_t SV_BASELINE = 1000000;
constexpr int64_t SV_TRANS_PENALTY = 2000;
constexpr int64_t SV_INV_PENALTY = 500;
constexpr int64_t SV_FRONT_END = 2;
constexpr int64_t REF_NEG_PENALTY = 2;
// run capacity of the enumeration's sorted front (entries of 32 B in shared memory)
#ifndef AA_RC_SLOTS
#define AA_RC_SLOTS 512
#endif
#ifndef AA_FCAP
#define AA_FCAP 512  // (1024 measured the same or slightly slower on C1 / C2 / C5; 16 KB lets more contigs be resident)
#endif
constexpr size_t HEAP2_FIXED_BYTES = 16 * 1024;  // f_heaps_chain: saved spines + op ring (>= sizeof(ChainSmem))
constexpr int32_t HEAP_CHUNK = 4096;  // leftist-heap nodes handed to a contig per arena grab
constexpr int64_t I64_MAX = 0x7fffffffffffffffLL;

// ---- distance types -----------------------------------------------------------------------------
// PafDistance (paf_data.hpp:121-189) in CALC_SUM mode only ever looks at qry+ref, so the k-walk
// pipeline carries the sum.  aux: reach flag in d[], unused elsewhere.
struct D4 {
    int64_t sum;
    int32_t anom, nz, tot, aux;
};
// QRY_SCORE mode (the gap-filling DP) needs qry and ref apart.
struct D5 {
    int64_t qry, ref;
    int32_t anom, nz, tot, pad;
};
AA_HD int64_t den(int32_t t) { return t ? (int64_t)t : 1; }
// operator< in CALC_SUM_MODE (paf_data.hpp:142-159); MAX never reaches here (callers test reach)
AA_HD bool less4(const D4 &a, const D4 &b) {
    if (a.sum != b.sum) return a.sum < b.sum;
    if (a.anom != b.anom) return a.anom < b.anom;
    return (int64_t)a.nz * den(b.tot) > (int64_t)b.nz * den(a.tot);
}
AA_HD bool less5(const D5 &a, const D5 &b) {  // QRY_SCORE_MODE
    if (a.qry != b.qry) return a.qry < b.qry;
    if (a.ref != b.ref) return a.ref < b.ref;
    if (a.anom != b.anom) return a.anom < b.anom;
    return (int64_t)a.nz * den(b.tot) > (int64_t)b.nz * den(a.tot);
}

// ---- edge record: 16 B ----------------------------------------------------------------------------
// dst_fl: bits 0..26 contig-local destination vertex, 27..28 anom, 29 qul_nonzero, 30 qul_total
struct Edge {
    int64_t qry;
    int32_t ref;
    uint32_t dst_fl;
};
constexpr uint32_t DST_MASK = (1u << 27) - 1;
AA_HD int32_t e_dst(const Edge &e) { return (int32_t)(e.dst_fl & DST_MASK); }
AA_HD int32_t e_anom(const Edge &e) { return (int32_t)((e.dst_fl >> 27) & 3u); }
AA_HD int32_t e_nz(const Edge &e) { return (int32_t)((e.dst_fl >> 29) & 1u); }
AA_HD int32_t e_tot(const Edge &e) { return (int32_t)((e.dst_fl >> 30) & 1u); }

// persistent leftist-heap node: 32 B, the (u,v) payload lives in hn_eid[]
struct __attribute__((aligned(16))) HNode {
    int64_t sum;
    int32_t anom, nz, tot;
    int32_t left, right;
    int16_t rank;   // node_rank of leftist_heap.hpp
    int16_t lrank;  // rank of the left child (saves a dependent load when the node is copied)
};
struct __attribute__((aligned(16))) V16 {
    int64_t a, b;
};
AA_HD HNode hn_load(const HNode *p) {  // two 128-bit loads
    union {
        HNode n;
        V16 v[2];
    } u;
    const V16 *q = reinterpret_cast<const V16 *>(p);
    u.v[0] = q[0];
    u.v[1] = q[1];
    return u.n;
}
AA_HD void hn_store(HNode *p, const HNode &n) {  // two 128-bit stores
    union {
        HNode n;
        V16 v[2];
    } u;
    u.n = n;
    V16 *q = reinterpret_cast<V16 *>(p);
    q[0] = u.v[0];
    q[1] = u.v[1];
}
// sidetrack key of an edge, precomputed by a parallel pass: c = w + d[v] - d[u] (k_shortest_walks.hpp:206)
struct __attribute__((aligned(8))) SKey {
    int64_t sum;
    int32_t anom, nz, tot;
    int32_t use;  // 0: not inserted (head unreachable, or the tree edge u -> best[u])
};
// priority-queue entry of the enumeration: 32 B
struct __attribute__((aligned(16))) PQEnt {
    int64_t sum;
    int32_t anom, nz, tot;
    int32_t node;  // heap node id (allocation order == the reference's pointer order)
    int32_t idx;   // entry index
    int32_t pad;
};

// packed reverse-graph record: everything the relax step needs about one in-edge, 16 B
struct __attribute__((aligned(16))) RevRec {
    int64_t sum;   // qry + ref
    int32_t src;   // contig-local source vertex
    uint32_t fl;   // bits 0..1 anom, 2 qul_nonzero, 3 qul_total, 4..31 out-degree of src (saturating; its initial Kahn count)
};
constexpr uint32_t RR_DEG_SAT = (1u << 28) - 1u;
// relax state of one vertex, 32 B: d[v] (CALC_SUM view), best[v], remaining out-degree, min anom to dest
struct __attribute__((aligned(16))) VState {
    int64_t sum;
    int32_t anom, nz, tot;
    int32_t best;
    int32_t cnt;
    int32_t amin_reach;  // bit 0: reaches dest; bits 1..: min anom sum to dest
};
// distance from a segment's upper articulation vertex to dest: added to the segment-local states
struct __attribute__((aligned(16))) SegShift {
    int64_t sum;
    int32_t anom, nz, tot;
    int32_t amin_reach;  // bit 0: the articulation vertex reaches dest; bits 1..: its min anom sum
    int32_t pad[2];
};
// one mapq-ratio tie-break of a segment run: with the true counts (N, T) of the upper articulation vertex the compare
// cand < cur reads c0 + T * c1 + N * c2 > 0 (cross-multiplied ratios, paf_data.hpp:152-158); the run stands iff that
// still gives `out`
struct __attribute__((aligned(8))) SegCon {
    int64_t c0;
    int32_t c1, c2;
    int32_t out, pad;
};
constexpr int32_t SEG_BLOCKS = 256;
constexpr int32_t SEG_MAXCON = 16;
// per edge (u,v): root of v's sidetrack heap and its key, so that a pop needs one load instead of three
struct __attribute__((aligned(16))) ENext {
    int64_t sum;
    int32_t anom, nz, tot;
    int32_t hv;     // hroot[v] or -1
    int32_t hrank;  // its rank in the sequential allocation order (device path)
    int32_t pad;
};
// per heap node: everything a pop of this node needs (children, next heap root, their key deltas and order keys), gathered
// by a parallel pass once the heaps are built, so that the enumeration's dependent chain is ONE load per pop instead of
// node -> (children, edge) -> (keys)
struct __attribute__((aligned(16))) XRec {
    int32_t left, right, hv, hkey;  // successor nodes (-1: none); hkey/lkey/rkey: their order keys (id, or allocation rank)
    int64_t lsum, rsum;             // key of the child minus key of the node
    int64_t hsum;                   // key of the next heap's root
    int32_t lanom, ranom;
    int32_t hanom, lnz, ltot, rnz;
    int32_t rtot, hnz, htot, lkey;
    int32_t rkey, pad0, pad1, pad2;
};
static_assert(sizeof(XRec) == 96, "XRec is six 16-byte words");
// ---- BFS order of the shortest-path tree, computed in parallel (Euler tour + list ranking), and the
// flat stream of sidetrack inserts in that order (k_shortest_walks.hpp:196-215) -------------------------------
constexpr uint32_t TOUR_END = 0xffffffffu;
struct __attribute__((aligned(16))) Tour {  // one tour element: D_x = 2x (entering x), U_x = 2x+1 (leaving x)
    uint32_t nxt;  // successor element (global index) or TOUR_END
    int32_t sp;    // suffix count of D elements: preorder rank from the end
    int32_t sd;    // suffix sum of +1 (D) / -1 (U): depth
    int32_t pad;
};
struct __attribute__((aligned(16))) VInfo {  // one tree vertex at its BFS position
    int32_t x;         // contig-local vertex
    uint32_t ins_beg;  // first record of its inserts in ins[] (global index)
    int32_t nins;      // number of inserts; bit 30: has tree children
    int32_t ppos;      // BFS position of its tree parent (-1 for dest)
};
constexpr int32_t VI_KIDS = 1 << 30;
struct __attribute__((aligned(8))) InsKey {  // one sidetrack to insert: key c = w + d[v] - d[u] and its edge
    int64_t sum;
    int32_t anom, nz, tot;
    int32_t eid;  // contig-local edge id
};
struct __attribute__((aligned(16))) HOp {  // one operation of the serial heap builder (f_heaps_chain), 32 B
    int64_t sum;
    int32_t anom, nz, tot;  // insert: the sidetrack key (reservation: anom = number of node ids)
    int32_t eid;
    int32_t ctl;            // HOP_FIRST: first insert of its vertex, HOP_LAST: last one; FIRST also carries, from bit HOP_RESV_SHIFT
                            // up, the node ids to reserve BEFORE this vertex for the tree leaves that precede it in BFS order
    int32_t aux;            // FIRST: chain ordinal (within the contig) of the vertex whose heap is inherited, -1: empty heap
};
constexpr int32_t HOP_FIRST = 1, HOP_LAST = 2, HOP_RESV_SHIFT = 3;
constexpr int32_t LEAF_MAX_INS = 32;  // a leaf's reserved ids (32 per insert) fit one arena chunk
struct CandRec {  // one candidate pair (i partially overlaps j): cut point, 48 B
    int32_t i, j;        // contig-local sorted block indices
    int64_t pe_q, pe_r;  // edited_loc_pre_end[i][j]
    int64_t st_q, st_r;  // edited_loc_str[i][j]
};

struct Task {  // one edge_path_to_paf_path call
    int32_t ctg, walk;
    int32_t call;   // call order inside the contig (OC11)
    int32_t group;  // 0 = ties of walk 0; g>=1 = g-th alt answer group
};

// ---- workspace: every array of the pipeline (device pointers) ----------------------------------------
struct Ws {
    // batch
    int64_t C, B, R;
    int32_t nsl, K;
    const int64_t *ctg_off;  // [C+1]
    // original-order input
    const int64_t *in_qs, *in_qe, *in_rs, *in_re, *in_qtot;
    const int32_t *in_chr;
    const uint8_t *in_fwd, *in_mapq;
    const int64_t *run_off, *run_ql, *run_qr, *run_rl;
    const int32_t *perm;  // [B] sorted position -> original position inside its contig
    // sorted blocks
    int32_t *blk_ctg;  // [B]
    int64_t *qs, *qe, *rs, *re, *qtot;
    int32_t *chr, *orig;
    uint8_t *fwd, *mapq;
    int64_t *run_beg;
    int32_t *run_cnt;
    int32_t *part_l, *part_r;  // [B] contig-local [l, r) of the block's part
    // candidate pairs / pair vertices
    int32_t *cand_cnt;   // [B]
    int64_t *cand_off;   // [B+1]
    CandRec *cand;       // [Ncand]
    int32_t *cand_ok;    // [Ncand]
    int64_t *cand_rank;  // [Ncand+1]
    CandRec *pair;       // [P] compacted (global pair index)
    int64_t *pair_beg;   // [B+1] first global pair index whose pre-block is b
    // vertices / edges
    int64_t *vtx_off;    // [C+1]
    int32_t *deg;        // [Vtot]
    int64_t *eoff;       // [Vtot+1]
    Edge *edge;          // [E]
    int32_t *e_src;      // [E] contig-local source
    uint32_t *rkey_in, *rkey;   // [E] global destination vertex (sort key) before / after
    uint32_t *rval_in, *rev_eid;  // [E] global edge id, sorted by destination, stable in source order
    int64_t *rev_off;    // [Vtot+1]
    struct RevRec *rrec; // [E] packed reverse records (source, weight), same order as rev_eid
    struct VState *vs;   // [Vtot] relax state record (device relax only)
    int32_t *cnt2;       // [Vtot] in-degree counters of the forward Kahn pass (runs concurrently with relax)
    // relax / topo
    D4 *d;               // [Vtot]  aux = reachable
    int32_t *best;       // [Vtot]
    int32_t *cnt;        // [Vtot] scratch in/out-degree counters
    int32_t *queue;      // [Vtot] scratch FIFO
    int32_t *order;      // [Vtot] forward Kahn position
    int32_t *topo;       // [Vtot] forward Kahn order (vertex at position)
    int32_t *amin;       // [Vtot] min anom sum to dest
    int64_t *anom_dis;   // [C]
    int32_t *status;     // [C] 0 ok, 1 singleton, 2 unsolvable, 3 heap arena overflow
    // sidetrack heaps
    HNode *hn;           // [Hcap]
    int32_t *hn_eid;     // [Hcap] contig-local edge id (u,v) of the sidetrack
    int64_t Hcap;
    unsigned long long *heap_top;  // arena bump pointer
    int32_t *hroot;      // [Vtot]
    SKey *skey;          // [E] sidetrack keys (parallel pre-pass)
    int32_t *child;      // [E] tree children of v, compacted at rev_off[v] (ascending source)
    int32_t *nchild;     // [Vtot]
    int32_t *nins;       // [Vtot] sidetracks of the vertex that are inserted
    uint32_t *cslot;     // [Vtot] slot of the vertex in its parent's child list (index into child[])
    Tour *tour_a, *tour_b;  // [2*Vtot] Euler tour of every contig's tree, double-buffered for pointer jumping
    uint64_t *bkey_in, *bkey;  // [Vtot] (contig, depth, preorder) sort keys
    uint32_t *bval_in, *bfs_vtx;  // [Vtot] global vertex ids; bfs_vtx[v0 + pos] = vertex at BFS position pos
    int32_t *bfspos;     // [Vtot] BFS position of a tree vertex, -1 outside the tree
    int32_t *ntree;      // [C] vertices in the tree
    VInfo *vinfo;        // [Vtot] by BFS position
    int32_t *ins_cnt;    // [Vtot+1] inserts per BFS position
    int64_t *ins_off;    // [Vtot+2]
    InsKey *ins;         // [Nins] inserts of every contig, BFS order then edge order
    int32_t *root_at;    // [Vtot] heap root by BFS position
    int32_t key_bits;    // bits of the depth / preorder fields of the sort key
    // segment-parallel relax of chain-like contigs (see f_relax_seg_warp)
    int64_t *seg_boff;   // [C+1] first bucket (SEG_BLOCKS sorted blocks) of every contig
    int32_t *seg_bnd;    // [TB] first articulation block of the bucket (contig-local) or -1; bucket 0 of a contig: -1
    int32_t *seg_ncon;   // [TB] mapq-ratio tie-breaks the segment's run decided (each one a condition on the true shift)
    struct SegCon *seg_con;  // [TB * SEG_MAXCON] those conditions
    int32_t *seg_seed;   // [TB * 2] (qul_nonzero, qul_total) the run assumed for its upper articulation vertex
    struct SegShift *seg_shift;  // [TB] what relax_unpack adds to the segment's local distances
    int32_t *seg_mode;   // [C] 1: vs[] holds segment-local states
    int32_t seg_guess;   // test knob (AA_SEG_GUESS): 0 = plausible counts, 1 = (0, 1): most tie-break conditions then fail and
                         // the sweep's redo path runs, 2 = also refuse every condition (every tied segment is redone)
    int32_t *topo_redo;  // [C] 1: the segmented forward pass did not cover the contig, redo it in one piece
    int64_t TB;
    // level-synchronous Kahn passes for wide, shallow DAGs (dense contigs)
    int32_t *rmode;      // [C] -1: warp-per-contig Kahn passes, k >= 0: k-th contig of the level-synchronous passes
    int32_t *kl_cnt;     // [Vtot] remaining degree
    unsigned long long *kl_last;  // [Vtot] (contig k, position of the last-finishing neighbour, index in its list)
    int32_t *kl_pos;     // [Vtot] FIFO position (reverse pass)
    uint32_t *kl_front, *kl_next;  // [Vtot] current / next level (global vertex ids)
    int32_t *kl_nnext;   // [1]
    unsigned long long *kl_key_in, *kl_key;  // [Vtot] sort keys of the next level
    uint32_t *kl_val;    // [Vtot] next level, sorted
    int32_t *kl_done;    // [64] vertices positioned so far, per level-mode contig
    int32_t *cdepth;     // [C] depth of the tree
    int32_t *hmode;      // [C] 0: streaming builder (one warp per contig), 1: level-parallel builder (shallow, wide trees)
    int32_t *lvl_overflow;  // arena / spine overflow flag of the level-parallel builder
    // order of the heap nodes: the reference's queue breaks ties by node address = allocation order.  Nodes are
    // built out of that order (leaves and dense contigs in parallel), so every node records (owner's BFS slot, running
    // number inside the owner); after the build that pair is replaced by the node's rank in the sequential order
    unsigned long long *hn_key;  // [Hcap] slot << 32 | number, then the rank
    int32_t *vcnt;       // [Vtot+1] nodes allocated by the vertex at each BFS slot
    int64_t *vbase;      // [Vtot+2] exclusive prefix sum of vcnt
    int32_t *leaf_flag;  // [Vtot+1] 1: tree leaf with inserts of a streaming-mode contig (built by f_heaps_level)
    int64_t *leaf_off;   // [Vtot+2]
    uint32_t *leaf_list; // [n_leaf] their BFS slots
    int32_t *leaf_base;  // [Vtot] first of the 32 * nins node ids the streaming builder reserved for a leaf (keeps ids in order)
    ENext *enext;        // [E] (device enumeration only)
    XRec *xrec;          // [heap_top] expansion records (device enumeration; nullptr when they would not fit)
    int32_t *chunk_ctg;  // [Hcap / 64 + 1] contig that owns each run of 64 node ids (written by the heap builders)
    int64_t *heap_used;  // [C]
    // streaming builder, flat form (f_heaps_chain): the serial warp reads one stream of operations per contig
    HOp *ops;            // [n_ops] inserts of the chain vertices (tree vertices whose heap others inherit)
    int32_t *leaf_need;  // [Vtot+1] node ids a tree leaf reserves (32 per insert), 0 for every other vertex
    int64_t *leaf_lp;    // [Vtot+2] exclusive prefix sum of leaf_need
    int32_t *chain_slot; // [n_chain] BFS slot (global) of every chain vertex
    int32_t *resv_base;  // [n_chain + C] first reserved id of the leaf run before every chain vertex (+ one per contig: the run
                         // behind its last chain vertex), written by the serial warp
    int32_t *op_cnt;     // [Vtot+1] operations of the vertex at each BFS slot
    int64_t *op_off;     // [Vtot+2]
    int32_t *chain_flag; // [Vtot+1] 1: chain vertex
    int64_t *chain_ord;  // [Vtot+2] exclusive prefix sum of chain_flag
    int32_t *owner;      // [Vtot] contig-local BFS slot of the nearest chain vertex among the vertex and its tree ancestors (-1: none)
    int32_t *chain_root; // [n_chain] heap root of every chain vertex, written by the serial warp
    int32_t heap_cache_bits;  // log2 of the on-chip node cache of f_heaps_chain (entries of 40 B); 0: none
    int32_t heaps_variant;    // tuning switch (tools/heap_lab)
    // Two-group pipelining of heaps -> enumeration (aa_pipeline.cuh): the passes between the two are launched once per group,
    // each over all items, and skip the contigs of the other group.  grp == nullptr / grp_sel < 0: no filter.
    const int8_t *grp;        // [C] group of the contig (1: long serial chains, 0: the rest)
    int32_t grp_sel;          // the group this launch works on
    // enumeration
    int64_t *walk_off;   // [C+1] = c*K
    int32_t *n_walk;     // [C]
    D4 *wdist;           // [C*K]
    int32_t *wlast;      // [C*K] path_last_node
    int32_t *ent_node, *ent_prev;  // [C*3K]
    PQEnt *pq;           // [C*3K]
    PQEnt *pq_far;       // [C*3K] second backlog region of the device enumeration (null: one region)
    // tasks
    Task *task;          // [C*2K]
    int32_t *n_task;     // [C]
    int32_t *n_tie;      // [C] tasks of group 0
    int32_t *last_group; // [C]
    int64_t *task_off;   // [C+1] compacted offsets (after scan of n_task)
    Task *tasks;         // [Ntask] compacted
    int64_t *task_cov;   // [Ntask]
    int32_t *task_rows;  // [Ntask]
    int32_t *first_call; // [B] min call index that marked the block (not_alt_vertex_map)
    // per-slot scratch for walk tasks
    int64_t slot_stride;  // elements per slot (max V over contigs)
    int32_t *sc_walk, *sc_up, *sc_side;  // [S*stride]
    D5 *sc_dp;            // [S*stride]
    int32_t *sc_pre;      // [S*stride]
    uint8_t *sc_seen;     // [S*stride]
    unsigned long long *task_next;  // dynamic task counter
    // main chain of every contig = upgrade-automaton states along walk 0 (the shortest-path-tree walk from
    // src), with prefix sums; the other walks only simulate where they leave it (see walk_task_inc)
    int32_t *main_pos;   // [Vtot] position of the vertex on walk 0, -1 when not on it
    int32_t *main_walk;  // [Vtot] vertex at position i of walk 0
    int32_t *m_cs;       // [Vtot] upgraded end point when the automaton stands at position i (-1: no state)
    int64_t *m_cov;      // [Vtot] coverage of the rows closed before position i
    int32_t *m_rows;     // [Vtot] rows closed before position i
    int64_t *m_tot_cov;  // [C]
    int32_t *m_tot_rows; // [C]
    int32_t *m_len;      // [C] number of edges of walk 0 (dest sits at position m_len)
    uint8_t *m_done;     // [Vtot] 1: the resolve pass computed (and emitted) this state's step itself
    // speculative step results, assuming cs == walk vertex (true for >99% of the states)
    int32_t *sp_cs;      // [Vtot] cs after the step
    int32_t *sp_rows;    // [Vtot] rows closed by the step
    int64_t *sp_cov;     // [Vtot] coverage closed by the step
    uint8_t *sp_used;    // [Vtot] walk edges consumed (1/2), 0 = not speculated (DP range too large)
    // rows of the upgraded walk 0, index = row number (emitted in parallel by f_main_rows)
    int32_t *mr_blk;     // [Vtot] contig-local sorted block
    int64_t *mr_qs, *mr_qe, *mr_rs, *mr_re;
    int64_t n_tasks_total;
    int32_t *fb_list;     // [Ntask] tasks the one-thread-per-task pass handed back
    int32_t *fb_n;        // [1]
    int64_t fb_count;
    // selection + output
    int32_t *win_out, *win_alt;  // [C] compacted task index (or -1)
    int32_t *out_cnt, *alt_cnt, *all_cnt;  // [C] rows / rows / paths
    int64_t *out_off, *alt_off, *all_path_off;  // [C+1]
    int32_t *all_task;    // [Npaths] task index of each .all path
    int32_t *all_rows;    // [Npaths]
    int64_t *all_row_off; // [Npaths+1]
    // rows: destination arrays (out / alt / all)
    int32_t *r_idx[3];
    int64_t *r_qs[3], *r_qe[3], *r_rs[3], *r_re[3];
    uint8_t *r_alt[3];
};

struct Ctg {
    int64_t b0;   // first (sorted) block
    int32_t n;    // blocks
    int64_t p0;   // first global pair index
    int32_t P;    // pair vertices
    int32_t V, src, dest;
    int64_t v0;   // first global vertex
};
AA_HD Ctg ctg_view(const Ws &w, int64_t c) {
    Ctg g;
    g.b0 = w.ctg_off[c];
    g.n = (int32_t)(w.ctg_off[c + 1] - g.b0);
    g.p0 = w.pair_beg[g.b0];
    g.P = (int32_t)(w.pair_beg[g.b0 + g.n] - g.p0);
    g.V = g.n + g.P + 2;
    g.src = g.n + g.P;
    g.dest = g.src + 1;
    g.v0 = w.vtx_off[c];
    return g;
}
template <class T>
AA_HD int64_t upper_idx(const T *off, int64_t n, T x) {  // largest c in [0,n) with off[c] <= x
    int64_t lo = 0, hi = n;
    while (hi - lo > 1) {
        int64_t mid = (lo + hi) >> 1;
        if (off[mid] <= x) lo = mid;
        else hi = mid;
    }
    return lo;
}

// ================================================================================================
// phase: gather blocks into sorted order
AA_HDN void f_gather(const Ws &w, int64_t b) {
    int64_t c = upper_idx(w.ctg_off, w.C, b);
    int64_t b0 = w.ctg_off[c];
    int64_t s = b0 + w.perm[b];
    w.blk_ctg[b] = (int32_t)c;
    w.qs[b] = w.in_qs[s];
    w.qe[b] = w.in_qe[s];
    w.rs[b] = w.in_rs[s];
    w.re[b] = w.in_re[s];
    w.qtot[b] = w.in_qtot[s];
    w.chr[b] = w.in_chr[s];
    w.fwd[b] = w.in_fwd[s];
    w.mapq[b] = w.in_mapq[s];
    w.orig[b] = w.perm[b];
    w.run_beg[b] = w.run_off[s];
    w.run_cnt[b] = (int32_t)(w.run_off[s + 1] - w.run_off[s]);
    w.first_call[b] = 0x7fffffff;
}

// phase: parts (paf_data.cpp:249-261), one contig per call
AA_HDN void f_parts(const Ws &w, int64_t c) {
    if (aa_lane() != 0) return;
    int64_t b0 = w.ctg_off[c];
    int32_t n = (int32_t)(w.ctg_off[c + 1] - b0);
    w.status[c] = (n == 1) ? 1 : 0;
    int64_t part_end = -1;
    int32_t start = 0;
    for (int32_t i = 0; i < n; i++) {
        int64_t s = w.qs[b0 + i];
        if (part_end < s) {
            for (int32_t k = start; k < i; k++) w.part_r[b0 + k] = i;
            start = i;
        }
        w.part_l[b0 + i] = start;
        int64_t e = w.qe[b0 + i];
        if (e > part_end) part_end = e;
    }
    for (int32_t k = start; k < n; k++) w.part_r[b0 + k] = n;
}

// parts, device form: three segmented scans over all blocks of the batch (keys = contig of the block) instead of one warp per
// contig.  pm = exclusive prefix max of the ends; a block is a boundary iff it starts beyond pm (paf_data.cpp:249-261);
// part_l = inclusive prefix max of (boundary ? index : -1); part_r = exclusive suffix min of (boundary ? index : +inf).
AA_HDN void f_parts_bnd(const Ws &w, int64_t i, const int64_t *pm, int32_t *bl, int32_t *br) {
    const int64_t c = w.blk_ctg[i];
    const int64_t b0 = w.ctg_off[c];
    const int32_t li = (int32_t)(i - b0);
    const bool bnd = pm[i] < w.qs[i];
    bl[i] = bnd ? li : -1;
    br[i] = bnd ? li : 0x7fffffff;
    if (li == 0) w.status[c] = (w.ctg_off[c + 1] - b0 == 1) ? 1 : 0;
}
AA_HDN void f_parts_fin(const Ws &w, int64_t i, const int32_t *br) {
    const int64_t c = w.blk_ctg[i];
    const int32_t n = (int32_t)(w.ctg_off[c + 1] - w.ctg_off[c]);
    w.part_r[i] = br[i] < n ? br[i] : n;
}

// qry_partial_overlap for sorted i < j (paf_data.hpp:78-86)
AA_HD bool partial_ij(const Ws &w, int64_t gi, int64_t gj) {
    int64_t is = w.qs[gi], ie = w.qe[gi], js = w.qs[gj], je = w.qe[gj];
    if (is < js) return js <= ie && ie < je;
    if (js < is) return is <= je && je < ie;
    return false;
}
AA_HD bool contains_ij(const Ws &w, int64_t gi, int64_t gj) {  // i contains j (paf_data.hpp:74-77)
    return w.qs[gi] <= w.qs[gj] && w.qe[gj] <= w.qe[gi];
}

// phase: count candidate pairs of block b (paf_data.cpp:297-302)
AA_HDN void f_cand_count(const Ws &w, int64_t b) {
    int64_t c = w.blk_ctg[b];
    int64_t bend = w.ctg_off[c + 1];
    int32_t cnt = 0;
    int64_t ie = w.qe[b];
    for (int64_t j = b + 1; j < bend; j++) {
        if (ie < w.qs[j]) break;
        if (partial_ij(w, b, j)) cnt++;
    }
    w.cand_cnt[b] = cnt;
}

// cut point of a partially overlapping pair: the two-pointer scan of paf_data.cpp:303-376
AA_HDN bool find_cut(const Ws &w, int64_t gi, int64_t gj, CandRec &r) {
    int64_t ai = w.run_beg[gi], bi = w.run_beg[gj];
    int64_t na = w.run_cnt[gi], nb = w.run_cnt[gj];
    int64_t step_a = w.fwd[gi] ? 1 : -1, step_b = w.fwd[gj] ? 1 : -1;
    int64_t min_gap = -1, g_i = -1, g_j = -1;
    int64_t pi = 0, pj = 0;
    while (pi < na && pj < nb) {
        int64_t li = w.run_ql[ai + pi], ri = w.run_qr[ai + pi];
        int64_t lj = w.run_ql[bi + pj], rj = w.run_qr[bi + pj];
        if (li == lj) {
            if (lj == rj) {
                pj++;
                continue;
            }
            r.pe_q = li;
            r.pe_r = w.run_rl[ai + pi];
            r.st_q = lj + 1;
            r.st_r = w.run_rl[bi + pj] + step_b;
            return true;
        }
        if (li < lj) {
            if (lj <= ri + 1) {
                r.pe_q = lj - 1;
                r.pe_r = w.run_rl[ai + pi] + ((lj - 1) - li) * step_a;
                r.st_q = lj;
                r.st_r = w.run_rl[bi + pj];
                return true;
            }
            int64_t gap = lj - (ri + 1);
            if (min_gap == -1 || gap < min_gap) {
                min_gap = gap;
                g_i = pi;
                g_j = pj;
            }
            pi++;
        } else {
            if (li <= rj - 1) {
                r.pe_q = li;
                r.pe_r = w.run_rl[ai + pi];
                r.st_q = li + 1;
                r.st_r = w.run_rl[bi + pj] + (li + 1 - lj) * step_b;
                return true;
            }
            pj++;
        }
    }
    if (min_gap == -1) return false;  // the release build drops such a pair (paf_data.cpp:373-375)
    int64_t li = w.run_ql[ai + g_i], ri = w.run_qr[ai + g_i];
    r.pe_q = ri;
    r.pe_r = w.run_rl[ai + g_i] + (ri - li) * step_a;
    r.st_q = w.run_ql[bi + g_j];
    r.st_r = w.run_rl[bi + g_j];
    return true;
}

// phase: cut points of block b's candidates
AA_HDN void f_cuts(const Ws &w, int64_t b) {
    int64_t c = w.blk_ctg[b];
    int64_t b0 = w.ctg_off[c], bend = w.ctg_off[c + 1];
    int64_t slot = w.cand_off[b];
    int64_t ie = w.qe[b];
    for (int64_t j = b + 1; j < bend; j++) {
        if (ie < w.qs[j]) break;
        if (!partial_ij(w, b, j)) continue;
        CandRec r;
        r.i = (int32_t)(b - b0);
        r.j = (int32_t)(j - b0);
        r.pe_q = r.pe_r = r.st_q = r.st_r = -1;
        bool ok = find_cut(w, b, j, r);
        w.cand[slot] = r;
        w.cand_ok[slot] = ok ? 1 : 0;
        slot++;
    }
}

// phase: compact valid candidates into pair vertices (discovery order i^, j^ : OC2); item = slot or block
AA_HDN void f_compact(const Ws &w, int64_t k, int64_t ncand) {
    if (k < ncand && w.cand_ok[k]) w.pair[w.cand_rank[k]] = w.cand[k];
    if (k <= w.B) w.pair_beg[k] = w.cand_rank[w.cand_off[k]];
}
AA_HDN void f_vtx_off(const Ws &w, int64_t c) {  // c in [0, C]
    int64_t b = w.ctg_off[c];
    w.vtx_off[c] = b + w.pair_beg[b] + 2 * c;
    if (c < w.C) w.walk_off[c] = c * (int64_t)w.K;
}

// ---- vertex view (Internal_Vertex, paf_data.cpp:392-411) ---------------------------------------------
struct VV {
    int64_t cur;  // global sorted block index of cur_idx
    int64_t qs, qe, rs, re;
};
AA_HD VV vv_single(const Ws &w, int64_t gb) { return VV{gb, w.qs[gb], w.qe[gb], w.rs[gb], w.re[gb]}; }
AA_HD VV vv_pair(const Ws &w, const Ctg &g, const CandRec &p) {
    int64_t gj = g.b0 + p.j;
    return VV{gj, p.st_q, w.qe[gj], p.st_r, w.re[gj]};
}
AA_HD int64_t ref_abs(int64_t x) { return x < 0 ? -x * REF_NEG_PENALTY : x; }

// get_score (paf_data.cpp:449-521). rp != nullptr when the right vertex is a pair vertex.
AA_HD Edge score(const Ws &w, VV l, const VV &r, const CandRec *rp, int32_t dst) {
    if (rp) {
        l.qe = rp->pe_q;
        l.re = rp->pe_r;
    }
    int64_t ref_diff = 0;
    uint32_t anom = 0;
    bool lf = w.fwd[l.cur] != 0, rf = w.fwd[r.cur] != 0;
    if (w.chr[l.cur] == w.chr[r.cur]) {
        if (lf == rf) {
            int64_t gap = lf ? r.rs - (l.re + 1) : l.re - (r.rs + 1);
            ref_diff = ref_abs(gap);
        } else {
            anom = 1;
            ref_diff = SV_INV_PENALTY + (lf ? ref_abs(r.re - (l.re + 1)) : ref_abs(r.rs - (l.rs + 1)));
        }
        if (ref_diff > SV_BASELINE) {
            anom += 1;
            ref_diff = SV_BASELINE;
        }
    } else {
        anom = 1;
        ref_diff = SV_TRANS_PENALTY;
    }
    Edge e;
    e.qry = r.qs - l.qe - 1;
    e.ref = (int32_t)ref_diff;
    e.dst_fl = (uint32_t)dst | (anom << 27) | ((w.mapq[r.cur] ? 1u : 0u) << 29) | (1u << 30);
    return e;
}

// index_of_vtx[i][j] for a pair vertex, -1 when (i,j) is no vertex.  Pairs of i are contiguous, j ascending.
AA_HD int32_t pair_lookup(const Ws &w, const Ctg &g, int32_t i, int32_t j, int64_t &cursor) {
    int64_t end = w.pair_beg[g.b0 + i + 1];
    while (cursor < end && w.pair[cursor].j < j) cursor++;
    if (cursor < end && w.pair[cursor].j == j) return g.n + (int32_t)(cursor - g.p0);
    return -1;
}

// Enumerate the out-edges of contig-local vertex v in the reference's adjacency order (OC3).
// make_Graph (paf_data.cpp:531-696) restated per source vertex; Emit is called once per edge.
template <class Emit>
AA_HDN void enum_edges(const Ws &w, const Ctg &g, int32_t v, Emit &emit) {
    const bool nsl = w.nsl != 0;
    const int64_t b0 = g.b0;
    const int32_t n = g.n;
    if (v == g.dest) return;
    if (v == g.src) {  // paf_data.cpp:540-563
        int32_t r = w.part_r[b0];
        int64_t min_qe = I64_MAX;
        for (int32_t i = 0; i < r; i++) {
            int64_t s = w.qs[b0 + i];
            if (nsl) {
                if (min_qe < s) break;
                int64_t e = w.qe[b0 + i];
                if (e < min_qe) min_qe = e;
            }
            Edge ed;
            ed.qry = s * SV_FRONT_END;
            ed.ref = 0;
            ed.dst_fl = (uint32_t)i | ((w.mapq[b0 + i] ? 1u : 0u) << 29) | (1u << 30);
            emit(ed);
        }
        return;
    }
    const int32_t last_l = w.part_l[b0 + n - 1];
    const int64_t max_qs = w.qs[b0 + n - 1];
    int32_t j;        // cur_idx of the source vertex
    VV lv;
    const CandRec *sp = nullptr;
    if (v < n) {
        j = v;
        lv = vv_single(w, b0 + v);
    } else {
        sp = &w.pair[g.p0 + (v - n)];
        j = sp->j;
        lv = vv_pair(w, g, *sp);
    }
    const int64_t gj = b0 + j;
    const int32_t r = w.part_r[gj];
    const int64_t j_qe = w.qe[gj];
    // -> dest (paf_data.cpp:565-595)
    if (w.part_l[gj] == last_l && !(nsl && j_qe < max_qs)) {
        Edge ed;
        ed.qry = (w.qtot[gj] - j_qe - 1) * SV_FRONT_END;
        ed.ref = 0;
        ed.dst_fl = (uint32_t)g.dest;
        emit(ed);
    }
    // inside the part (paf_data.cpp:598-651)
    {
        int64_t min_after = I64_MAX;
        int64_t cursor = w.pair_beg[gj];  // pairs (j, k)
        for (int32_t k = j + 1; k < r; k++) {
            const int64_t gk = b0 + k;
            if (!sp && contains_ij(w, gj, gk)) continue;  // only the (i,i) row skips contained blocks (paf_data.cpp:605)
            const int64_t k_qs = w.qs[gk];
            if (nsl) {
                if (min_after < k_qs) break;
                if (j_qe < k_qs) {
                    int64_t e = w.qe[gk];
                    if (e < min_after) min_after = e;
                }
            }
            if (j_qe < k_qs) {
                emit(score(w, lv, vv_single(w, gk), nullptr, k));  // -> (k,k)
            } else {
                // -> (j,k): needs the pair vertex and lft.qry_str < rht.qry_str (paf_data.cpp:433-436)
                int32_t pid = pair_lookup(w, g, j, k, cursor);
                if (pid >= 0) {
                    const CandRec &rp = w.pair[g.p0 + (pid - n)];
                    if (lv.qs < rp.st_q) emit(score(w, lv, vv_pair(w, g, rp), &rp, pid));
                }
            }
        }
    }
    // to the next part (paf_data.cpp:653-695)
    if (r < n) {
        int32_t r2 = w.part_r[b0 + r];
        int64_t mn = I64_MAX;
        for (int32_t k = r; k < r2; k++) {
            const int64_t gk = b0 + k;
            if (nsl) {
                int64_t k_qs = w.qs[gk];
                if (mn < k_qs) break;
                if (j_qe < k_qs) {
                    int64_t e = w.qe[gk];
                    if (e < mn) mn = e;
                }
            }
            emit(score(w, lv, vv_single(w, gk), nullptr, k));
        }
    }
}

struct EmitCount {
    int32_t n = 0;
    AA_HD void operator()(const Edge &) { n++; }
};
struct EmitFill {
    Edge *edge;
    int32_t *src;
    uint32_t *key;
    uint32_t *val;
    int64_t at;
    int32_t u;
    uint32_t v0;
    AA_HD void operator()(const Edge &e) {
        edge[at] = e;
        src[at] = u;
        key[at] = v0 + (e.dst_fl & DST_MASK);
        val[at] = (uint32_t)at;
        at++;
    }
};

// phase: out-degree of global vertex gv
AA_HDN void f_degree(const Ws &w, int64_t gv) {
    int64_t c = upper_idx(w.vtx_off, w.C, gv);
    if (w.status[c] != 0) {
        w.deg[gv] = 0;
        return;
    }
    Ctg g = ctg_view(w, c);
    EmitCount ec;
    enum_edges(w, g, (int32_t)(gv - g.v0), ec);
    w.deg[gv] = ec.n;
}
// phase: fill the out-edges of global vertex gv
AA_HDN void f_fill(const Ws &w, int64_t gv) {
    int64_t c = upper_idx(w.vtx_off, w.C, gv);
    if (w.status[c] != 0) return;
    Ctg g = ctg_view(w, c);
    EmitFill ef{w.edge, w.e_src, w.rkey_in, w.rval_in, w.eoff[gv], (int32_t)(gv - g.v0), (uint32_t)g.v0};
    enum_edges(w, g, (int32_t)(gv - g.v0), ef);
}
// phase: reverse CSR offsets from the sorted destination keys
AA_HDN void f_rev_off(const Ws &w, int64_t gv, int64_t E) {  // gv in [0, Vtot]
    int64_t lo = 0, hi = E;  // first position with key >= gv
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if ((int64_t)w.rkey[mid] < gv) lo = mid + 1;
        else hi = mid;
    }
    w.rev_off[gv] = lo;
}

// ================================================================================================
// phase: reverse relax.  Kahn FIFO order on the reverse graph, seeds = vertices without out-edges in
// ascending id, first strict improvement wins (k_shortest_walks.hpp:132-175); plus the min-anom DP
// that replaces Dial's BFS (only anom_dis[dest] is consumed, paf_data.cpp:713,1615).
AA_HDN void f_relax(const Ws &w, int64_t c) {
    if (aa_lane() != 0) return;
    if (w.status[c] != 0) return;
    Ctg g = ctg_view(w, c);
    const int64_t v0 = g.v0;
    const int64_t e0 = w.eoff[v0];
    int32_t *q = w.queue + v0;
    int32_t tail = 0;
    for (int32_t v = 0; v < g.V; v++) {
        int32_t od = (int32_t)(w.eoff[v0 + v + 1] - w.eoff[v0 + v]);
        w.cnt[v0 + v] = od;
        D4 z;
        z.sum = 0;
        z.anom = z.nz = z.tot = 0;
        z.aux = 0;
        w.d[v0 + v] = z;
        w.best[v0 + v] = -1;
        w.amin[v0 + v] = 0x3fffffff;
        if (od == 0) q[tail++] = v;
    }
    w.d[v0 + g.dest].aux = 1;
    w.amin[v0 + g.dest] = 0;
    for (int32_t head = 0; head < tail; head++) {
        int32_t v = q[head];
        D4 dv = w.d[v0 + v];
        int32_t av = w.amin[v0 + v];
        int64_t ra = w.rev_off[v0 + v], rb = w.rev_off[v0 + v + 1];
        for (int64_t k = ra; k < rb; k++) {
            int64_t eid = w.rev_eid[k];
            int32_t x = w.e_src[eid];
            if (dv.aux) {
                Edge e = w.edge[eid];
                D4 cand;
                cand.sum = dv.sum + e.qry + e.ref;
                cand.anom = dv.anom + e_anom(e);
                cand.nz = dv.nz + e_nz(e);
                cand.tot = dv.tot + e_tot(e);
                cand.aux = 1;
                D4 cur = w.d[v0 + x];
                if (!cur.aux || less4(cand, cur)) {
                    w.d[v0 + x] = cand;
                    w.best[v0 + x] = v;
                }
                int32_t na = av + e_anom(e);
                if (na < w.amin[v0 + x]) w.amin[v0 + x] = na;
            }
            if (--w.cnt[v0 + x] == 0) q[tail++] = x;
        }
    }
    (void)e0;
    w.anom_dis[c] = w.amin[v0 + g.src];
    if (!w.d[v0 + g.src].aux) w.status[c] = 2;
}

// ---- parallel helpers around the warp-cooperative relax / topo (device path) ----------------------------------
AA_HDN void f_rev_pack(const Ws &w, int64_t k) {  // one reverse slot
    const int64_t eid = w.rev_eid[k];
    const Edge e = w.edge[eid];
    RevRec r;
    r.sum = e.qry + e.ref;
    r.src = w.e_src[eid];
    r.fl = (uint32_t)e_anom(e) | ((uint32_t)e_nz(e) << 2) | ((uint32_t)e_tot(e) << 3);
    const int64_t gs = upper_idx(w.eoff, w.vtx_off[w.C], eid);  // global source vertex: the one whose edge range holds eid
    const int64_t deg = w.eoff[gs + 1] - w.eoff[gs];
    r.fl |= (uint32_t)(deg < (int64_t)RR_DEG_SAT ? deg : (int64_t)RR_DEG_SAT) << 4;
    w.rrec[k] = r;
}
AA_HDN void f_relax_init(const Ws &w, int64_t gv) {  // one vertex
    VState s;
    s.sum = 0;
    s.anom = s.nz = s.tot = 0;
    s.best = -1;
    s.cnt = (int32_t)(w.eoff[gv + 1] - w.eoff[gv]);
    s.amin_reach = 0x3fffffff << 1;
    const int64_t c = upper_idx(w.vtx_off, w.C, gv);
    if (gv == w.vtx_off[c + 1] - 1) s.amin_reach = 1;  // dest: reaches itself, anom 0
    w.vs[gv] = s;
    w.cnt2[gv] = (int32_t)(w.rev_off[gv + 1] - w.rev_off[gv]);
}
// bucket that owns the segment of sorted block i of contig c (largest boundary <= i)
AA_HD int64_t seg_owner(const Ws &w, int64_t c, int32_t i) {
    const int64_t bo = w.seg_boff[c];
    int64_t m = i / SEG_BLOCKS;
    while (m > 0) {
        const int32_t b = w.seg_bnd[bo + m];
        if (b >= 0 && b <= i) break;
        m--;
    }
    return bo + m;
}
// local state of a segment + the distance of its upper articulation vertex = the state the unsegmented pass computes
AA_HD VState seg_compose(VState s, const SegShift &sh) {
    if (!(s.amin_reach & 1) || !(sh.amin_reach & 1)) {
        s.sum = 0;
        s.anom = s.nz = s.tot = 0;
        s.best = -1;
        s.amin_reach = 0x3fffffff << 1;
        return s;
    }
    s.sum += sh.sum;
    s.anom += sh.anom;
    s.nz += sh.nz;
    s.tot += sh.tot;
    s.amin_reach = (((s.amin_reach >> 1) + (sh.amin_reach >> 1)) << 1) | 1;
    return s;
}
AA_HDN void f_relax_unpack(const Ws &w, int64_t gv) {  // one vertex: VState -> d / best
    const int64_t c = upper_idx(w.vtx_off, w.C, gv);
    if (w.rmode[c] >= 0) return;  // written directly by the level-synchronous pass
    VState s = w.vs[gv];
    if (w.seg_mode && w.seg_mode[c]) {
        const Ctg g = ctg_view(w, c);
        const int32_t v = (int32_t)(gv - g.v0);
        const int32_t blk = v < g.n ? v : v < g.n + g.P ? w.pair[g.p0 + (v - g.n)].j : v == g.src ? 0 : g.n - 1;
        s = seg_compose(s, w.seg_shift[seg_owner(w, c, blk)]);
    }
    D4 d;
    d.sum = s.sum;
    d.anom = s.anom;
    d.nz = s.nz;
    d.tot = s.tot;
    d.aux = s.amin_reach & 1;
    w.d[gv] = d;
    w.best[gv] = s.best;
}

// phase: forward Kahn order (paf_data.cpp:742-746)
AA_HDN void f_topo(const Ws &w, int64_t c) {
    if (aa_lane() != 0) return;
    if (w.status[c] != 0) return;
    Ctg g = ctg_view(w, c);
    const int64_t v0 = g.v0;
    int32_t *q = w.topo + v0;
    int32_t tail = 0;
    for (int32_t v = 0; v < g.V; v++) {
        int32_t id = (int32_t)(w.rev_off[v0 + v + 1] - w.rev_off[v0 + v]);
        w.cnt[v0 + v] = id;
        if (id == 0) q[tail++] = v;
    }
    for (int32_t head = 0; head < tail; head++) {
        int32_t u = q[head];
        w.order[v0 + u] = head;
        int64_t ea = w.eoff[v0 + u], eb = w.eoff[v0 + u + 1];
        for (int64_t k = ea; k < eb; k++) {
            int32_t x = e_dst(w.edge[k]);
            if (--w.cnt[v0 + x] == 0) q[tail++] = x;
        }
    }
}

// ---- level-synchronous Kahn passes (dense contigs: tens of levels, hundreds of vertices per level) -------------
// The reference's FIFO order (k_shortest_walks.hpp:132-156) is: levels in order (level = longest path to a seed),
// and inside a level the order in which the vertices became ready, i.e. by (FIFO position of the neighbour that
// finished last, index in that neighbour's list).  Both are available without a queue: every list entry of a
// finished vertex does an atomicMax of that pair on its target, the targets whose degree reaches zero form the
// next level, one radix sort of their pairs gives their positions.  REV = reverse graph (relax), else forward.
constexpr int KL_POSB = 27;  // bits of a position / list index
AA_HD unsigned long long kl_key(int32_t k, int32_t pos, int32_t j) {
    return ((unsigned long long)(uint32_t)k << (2 * KL_POSB)) | ((unsigned long long)(uint32_t)pos << KL_POSB) | (unsigned long long)(uint32_t)j;
}
AA_HD int32_t aa_atomic_dec(int32_t *p) {  // returns the old value
#if defined(__CUDA_ARCH__)
    return atomicSub(p, 1);
#else
    return (*p)--;
#endif
}
AA_HD void aa_atomic_max64(unsigned long long *p, unsigned long long v) {
#if defined(__CUDA_ARCH__)
    atomicMax(p, v);
#else
    if (v > *p) *p = v;
#endif
}
AA_HD int32_t aa_atomic_inc(int32_t *p) {
#if defined(__CUDA_ARCH__)
    return atomicAdd(p, 1);
#else
    return (*p)++;
#endif
}
template <bool REV>
AA_HDN void f_kl_init(const Ws &w, int64_t gv) {
    const int64_t c = upper_idx(w.vtx_off, w.C, gv);
    const int32_t k = w.rmode[c];
    if (k < 0 || w.status[c] != 0) return;
    const int32_t deg = REV ? (int32_t)(w.eoff[gv + 1] - w.eoff[gv]) : (int32_t)(w.rev_off[gv + 1] - w.rev_off[gv]);
    w.kl_cnt[gv] = deg;
    w.kl_last[gv] = 0;
    if (deg == 0) {  // seeds in ascending id (k_shortest_walks.hpp:139-141)
        w.kl_last[gv] = kl_key(k, 0, (int32_t)(gv - w.vtx_off[c]));
        w.kl_next[aa_atomic_inc(w.kl_nnext)] = (uint32_t)gv;
    }
}
AA_HDN void f_kl_keys(const Ws &w, int64_t i) { w.kl_key_in[i] = w.kl_last[w.kl_next[i]]; }
// sorted slot i of the new level -> FIFO position
template <bool REV>
AA_HDN void f_kl_assign(const Ws &w, int64_t i, int64_t n) {
    const unsigned long long key = w.kl_key[i];
    const int32_t k = (int32_t)(key >> (2 * KL_POSB));
    const unsigned long long want = (unsigned long long)(uint32_t)k << (2 * KL_POSB);
    int64_t lo = 0, hi = n;  // first slot of contig k in this level
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (w.kl_key[mid] < want) lo = mid + 1;
        else hi = mid;
    }
    const int64_t gv = w.kl_val[i];
    const int64_t c = upper_idx(w.vtx_off, w.C, gv);
    const int64_t v0 = w.vtx_off[c];
    const int32_t pos = w.kl_done[k] + (int32_t)(i - lo);
    w.kl_front[i] = (uint32_t)gv;
    if (REV) {
        w.kl_pos[gv] = pos;
        w.queue[v0 + pos] = (int32_t)(gv - v0);
    } else {
        w.order[gv] = pos;
        w.topo[v0 + pos] = (int32_t)(gv - v0);
    }
}
AA_HDN void f_kl_count(const Ws &w, int64_t k, int64_t n) {  // level-mode contig k: vertices of this level
    const unsigned long long a = (unsigned long long)(uint32_t)k << (2 * KL_POSB), b = (unsigned long long)(uint32_t)(k + 1) << (2 * KL_POSB);
    int64_t lo = 0, hi = n, lo2 = 0, hi2 = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (w.kl_key[mid] < a) lo = mid + 1;
        else hi = mid;
    }
    while (lo2 < hi2) {
        const int64_t mid = (lo2 + hi2) >> 1;
        if (w.kl_key[mid] < b) lo2 = mid + 1;
        else hi2 = mid;
    }
    w.kl_done[k] += (int32_t)(lo2 - lo);
}
// one finished vertex (front slot i): its list entries count their targets down; lanes stride the list
template <bool REV>
AA_HDN void f_kl_expand(const Ws &w, int64_t i) {
    const int64_t gv = w.kl_front[i];
    const int64_t c = upper_idx(w.vtx_off, w.C, gv);
    const int64_t v0 = w.vtx_off[c];
    const int32_t k = w.rmode[c];
    const int32_t pos = REV ? w.kl_pos[gv] : w.order[gv];
    const int64_t a = REV ? w.rev_off[gv] : w.eoff[gv], b = REV ? w.rev_off[gv + 1] : w.eoff[gv + 1];
    for (int64_t e = a + aa_lane(); e < b; e += AA_LANES) {
        const int64_t gx = v0 + (REV ? w.e_src[w.rev_eid[e]] : e_dst(w.edge[e]));
        aa_atomic_max64(&w.kl_last[gx], kl_key(k, pos, (int32_t)(e - a)));
        if (aa_atomic_dec(&w.kl_cnt[gx]) == 1) w.kl_next[aa_atomic_inc(w.kl_nnext)] = (uint32_t)gx;
    }
}
// relax of one vertex of the new level (reverse pass): minimum over its out-edges of d[head] + w, the first popped
// head winning among equals (strict '<' on arrival order, k_shortest_walks.hpp:168-172); min-anom DP beside it
AA_HDN void f_kl_pull(const Ws &w, int64_t i) {
    const int64_t gv = w.kl_front[i];
    const int64_t c = upper_idx(w.vtx_off, w.C, gv);
    const int64_t v0 = w.vtx_off[c];
    const int32_t x = (int32_t)(gv - v0), dest = (int32_t)(w.vtx_off[c + 1] - v0) - 1;
    D4 best;
    best.sum = 0;
    best.anom = best.nz = best.tot = 0;
    best.aux = x == dest ? 1 : 0;
    int32_t bv = -1, bpos = 0x7fffffff, am = x == dest ? 0 : 0x3fffffff;
    for (int64_t e = w.eoff[gv] + aa_lane(); e < w.eoff[gv + 1]; e += AA_LANES) {
        const Edge ed = w.edge[e];
        const int32_t v = e_dst(ed);
        const D4 dv = w.d[v0 + v];
        if (!dv.aux) continue;
        D4 cand;
        cand.sum = dv.sum + ed.qry + ed.ref;
        cand.anom = dv.anom + e_anom(ed);
        cand.nz = dv.nz + e_nz(ed);
        cand.tot = dv.tot + e_tot(ed);
        cand.aux = 1;
        const int32_t pv = w.kl_pos[v0 + v];
        if (!best.aux || less4(cand, best) || (!less4(best, cand) && pv < bpos)) {
            best = cand;
            bv = v;
            bpos = pv;
        }
        const int32_t na = w.amin[v0 + v] + e_anom(ed);
        if (na < am) am = na;
    }
#if defined(__CUDA_ARCH__)
    for (int32_t d = 16; d > 0; d >>= 1) {  // combine the lanes
        D4 o;
        o.sum = __shfl_down_sync(0xffffffffu, best.sum, d);
        o.anom = __shfl_down_sync(0xffffffffu, best.anom, d);
        o.nz = __shfl_down_sync(0xffffffffu, best.nz, d);
        o.tot = __shfl_down_sync(0xffffffffu, best.tot, d);
        o.aux = __shfl_down_sync(0xffffffffu, best.aux, d);
        const int32_t ov = __shfl_down_sync(0xffffffffu, bv, d), op = __shfl_down_sync(0xffffffffu, bpos, d);
        const int32_t oa = __shfl_down_sync(0xffffffffu, am, d);
        if (o.aux && (!best.aux || less4(o, best) || (!less4(best, o) && op < bpos))) {
            best = o;
            bv = ov;
            bpos = op;
        }
        if (oa < am) am = oa;
    }
    if (aa_lane() != 0) return;
#endif
    if (!best.aux) {
        best.sum = 0;
        best.anom = best.nz = best.tot = 0;
    }
    w.d[gv] = best;
    w.best[gv] = bv;
    w.amin[gv] = am;
}
AA_HDN void f_kl_finish(const Ws &w, int64_t c) {  // anom_dis[dest] of the reference = min anom from src (paf_data.cpp:713)
    if (w.rmode[c] < 0 || w.status[c] != 0) return;
    const int64_t gs = w.vtx_off[c + 1] - 2;
    w.anom_dis[c] = w.amin[gs];
    if (!w.d[gs].aux) w.status[c] = 2;
}
AA_HDN void f_ctg_edges(const Ws &w, int64_t c, int64_t *out) { out[c] = w.eoff[w.vtx_off[c]]; }  // c in [0, C]

// ---- persistent leftist heap (leftist_heap.hpp:29-40) -------------------------------------------------
struct HeapAlloc {
    int64_t cur, end;
    int64_t used;
    bool overflow;
};
AA_HD int32_t heap_new(const Ws &w, HeapAlloc &ha) {
    if (ha.cur == ha.end) {
#if defined(__CUDA_ARCH__)
        unsigned long long at = atomicAdd(w.heap_top, (unsigned long long)HEAP_CHUNK);
#else
        unsigned long long at = *w.heap_top;
        *w.heap_top = at + HEAP_CHUNK;
#endif
        if ((int64_t)at + HEAP_CHUNK > w.Hcap) {
            ha.overflow = true;
            return -1;
        }
        ha.cur = (int64_t)at;
        ha.end = ha.cur + HEAP_CHUNK;
    }
    ha.used++;
    return (int32_t)(ha.cur++);
}
AA_HD D4 hkey(const HNode &h) {
    D4 k;
    k.sum = h.sum;
    k.anom = h.anom;
    k.nz = h.nz;
    k.tot = h.tot;
    k.aux = 0;
    return k;
}
// parallel pre-pass over vertices: sidetrack keys of u's out-edges and u's tree children
// (children of u = sources x of u's in-edges with best[x] == u, ascending x = reverse-list order)
AA_HDN void f_heap_prep(const Ws &w, int64_t gv) {
    const int64_t c = upper_idx(w.vtx_off, w.C, gv);
    w.nins[gv] = 0;
    w.nchild[gv] = 0;
    if (w.status[c] != 0 && w.status[c] != 3) return;
    const int64_t v0 = w.vtx_off[c];
    const D4 du = w.d[gv];
    const int32_t bu = w.best[gv];
    const int64_t ea = w.eoff[gv], eb = w.eoff[gv + 1];
    int32_t ni = 0;
    for (int64_t k = ea; k < eb; k++) {
        const Edge e = w.edge[k];
        const int32_t v = e_dst(e);
        const D4 dv = w.d[v0 + v];
        SKey sk;
        sk.sum = e.qry + e.ref + dv.sum - du.sum;
        sk.anom = e_anom(e) + dv.anom - du.anom;
        sk.nz = e_nz(e) + dv.nz - du.nz;
        sk.tot = e_tot(e) + dv.tot - du.tot;
        // unreachable heads are skipped; the tree edge has c == IDENTITY and is skipped once (it is unique)
        sk.use = (du.aux && dv.aux && v != bu) ? 1 : 0;
        ni += sk.use;
        w.skey[k] = sk;
    }
    w.nins[gv] = ni;
    const int32_t u = (int32_t)(gv - v0);
    const int64_t ra = w.rev_off[gv], rb = w.rev_off[gv + 1];
    int32_t n = 0;
    if (du.aux)
        for (int64_t k = ra; k < rb; k++) {
            const int32_t x = w.e_src[w.rev_eid[k]];
            if (w.best[v0 + x] == u) {
                w.cslot[v0 + x] = (uint32_t)(ra + n);
                w.child[ra + n++] = x;
            }
        }
    w.nchild[gv] = n;
}
// ---- BFS order of the tree without walking it: Euler tour, list ranking, sort by (depth, preorder) --------
// BFS (FIFO from dest, children in ascending id: k_shortest_walks.hpp:191-200) visits the tree level by level
// and, inside a level, in left-to-right order, which is the DFS preorder restricted to that level.
AA_HD bool in_tree(const Ws &w, int64_t c, int64_t gv, int32_t x, int32_t dest) {
    if (w.status[c] != 0 && w.status[c] != 3) return false;
    return w.d[gv].aux != 0 && (x == dest || w.best[gv] >= 0);
}
AA_HDN void f_tour_build(const Ws &w, int64_t gv) {
    const int64_t c = upper_idx(w.vtx_off, w.C, gv);
    const int64_t v0 = w.vtx_off[c];
    const int32_t x = (int32_t)(gv - v0), dest = (int32_t)(w.vtx_off[c + 1] - v0) - 1;
    Tour dn, up;
    dn.pad = up.pad = 0;
    if (!in_tree(w, c, gv, x, dest)) {
        dn.nxt = up.nxt = TOUR_END;
        dn.sp = dn.sd = up.sp = up.sd = 0;
    } else {
        dn.sp = 1;
        dn.sd = 1;
        dn.nxt = w.nchild[gv] > 0 ? (uint32_t)(2 * (v0 + w.child[w.rev_off[gv]])) : (uint32_t)(2 * gv + 1);
        up.sp = 0;
        up.sd = -1;
        if (x == dest) {
            up.nxt = TOUR_END;
        } else {
            const int64_t gp = v0 + w.best[gv];
            const int64_t slot = (int64_t)w.cslot[gv] + 1;
            up.nxt = slot < w.rev_off[gp] + w.nchild[gp] ? (uint32_t)(2 * (v0 + w.child[slot])) : (uint32_t)(2 * gp + 1);
        }
    }
    w.tour_a[2 * gv] = dn;
    w.tour_a[2 * gv + 1] = up;
}
AA_HD void tour_jump(const Tour *__restrict__ src, Tour *__restrict__ dst, int64_t i) {  // one pointer-jumping round
    Tour t = src[i];
    if (t.nxt != TOUR_END) {
        const Tour n = src[t.nxt];
        t.sp += n.sp;
        t.sd += n.sd;
        t.nxt = n.nxt;
    }
    dst[i] = t;
}
// sort key of a vertex: (contig, depth, preorder); vertices outside the tree go behind their contig's tree
AA_HDN void f_bfs_key(const Ws &w, int64_t gv, const Tour *__restrict__ tour) {
    const int64_t c = upper_idx(w.vtx_off, w.C, gv);
    const int64_t v0 = w.vtx_off[c];
    const int32_t V = (int32_t)(w.vtx_off[c + 1] - v0);
    const int32_t x = (int32_t)(gv - v0), dest = V - 1;
    const int kb = w.key_bits;
    uint64_t depth, pre;
    if (in_tree(w, c, gv, x, dest)) {
        const Tour me = tour[2 * gv], root = tour[2 * (v0 + dest)];
        depth = (uint64_t)(1 - me.sd);
        pre = (uint64_t)(root.sp - me.sp);
        if (x == dest) w.ntree[c] = root.sp;
#if defined(__CUDA_ARCH__)
        if (w.nchild[gv] == 0) atomicMax(&w.cdepth[c], (int32_t)depth);  // leaves suffice for the maximum
#else
        if ((int32_t)depth > w.cdepth[c]) w.cdepth[c] = (int32_t)depth;
#endif
    } else {
        depth = ((uint64_t)1 << kb) - 1;
        pre = (uint64_t)x;
        if (x == dest) w.ntree[c] = 0;
    }
    w.bkey_in[gv] = ((uint64_t)c << (2 * kb)) | (depth << kb) | pre;
    w.bval_in[gv] = (uint32_t)gv;
}
AA_HDN void f_lvl_off(const Ws &w, int64_t item, const int32_t *ctgs, int32_t per, int64_t *out) {
    const int64_t c = ctgs[item / per];
    const int64_t d = item % per + 1;  // depths 1 .. per
    const int kb = w.key_bits;
    const uint64_t want = ((uint64_t)c << (2 * kb)) | ((uint64_t)d << kb);
    int64_t lo = w.vtx_off[c], hi = w.vtx_off[c] + w.ntree[c];
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (w.bkey[mid] < want) lo = mid + 1;
        else hi = mid;
    }
    out[item] = lo;
}
AA_HDN void f_bfs_pos(const Ws &w, int64_t i) {  // i = sorted slot; contigs keep their vertex ranges
    const int64_t c = upper_idx(w.vtx_off, w.C, i);
    const int64_t v0 = w.vtx_off[c];
    const int64_t gv = w.bfs_vtx[i];
    w.bfspos[gv] = (i - v0) < w.ntree[c] ? (int32_t)(i - v0) : -1;
}
AA_HDN void f_vinfo(const Ws &w, int64_t i) {
    const int64_t c = upper_idx(w.vtx_off, w.C, i);
    const int64_t v0 = w.vtx_off[c];
    const int32_t dest = (int32_t)(w.vtx_off[c + 1] - v0) - 1;
    VInfo vi;
    vi.x = -1;
    vi.ins_beg = 0;
    vi.nins = 0;
    vi.ppos = -1;
    int32_t cnt = 0;
    if (i - v0 < w.ntree[c]) {
        const int64_t gv = w.bfs_vtx[i];
        vi.x = (int32_t)(gv - v0);
        cnt = w.nins[gv];
        vi.nins = cnt | (w.nchild[gv] > 0 ? VI_KIDS : 0);
        vi.ppos = vi.x == dest ? -1 : w.bfspos[v0 + w.best[gv]];
    }
    w.vinfo[i] = vi;
    w.ins_cnt[i] = cnt;
}
AA_HDN void f_ins_fill(const Ws &w, int64_t i) {  // the inserts of the vertex at sorted slot i, in edge order
    const int64_t c = upper_idx(w.vtx_off, w.C, i);
    const int64_t v0 = w.vtx_off[c];
    int64_t o = w.ins_off[i];
    w.vinfo[i].ins_beg = (uint32_t)o;
    if (i - v0 >= w.ntree[c] || w.ins_cnt[i] == 0) return;
    const int64_t gv = w.bfs_vtx[i];
    const int64_t e0 = w.eoff[v0];
    const int64_t ea = w.eoff[gv], eb = w.eoff[gv + 1];
    for (int64_t k = ea; k < eb; k++) {
        const SKey sk = w.skey[k];
        if (!sk.use) continue;
        InsKey ik;
        ik.sum = sk.sum;
        ik.anom = sk.anom;
        ik.nz = sk.nz;
        ik.tot = sk.tot;
        ik.eid = (int32_t)(k - e0);
        w.ins[o++] = ik;
    }
}

// ---- flat operation stream of the serial heap builder (f_heaps_chain) ---------------------------------------------
// Tree vertices fall in three classes: CHAIN (has inserts and either tree children or more than LEAF_MAX_INS inserts: its heap
// is inherited or too big for the leaf builder), LEAF (has inserts, no children: built off the chain by f_heaps_level) and the
// rest (no inserts: the heap is the parent's).  The serial warp sees only the inserts of the chain vertices, in BFS order, and
// one id reservation per leaf (so that node ids stay in allocation order); which heap a vertex starts from is
// resolved here, in parallel, by pointer jumping to the nearest chain ancestor.
AA_HDN void f_ops_class(const Ws &w, int64_t i) {
    const int64_t c = upper_idx(w.vtx_off, w.C, i);
    const int64_t v0 = w.vtx_off[c];
    int32_t cls = 0, n = 0, own = -1;
    if (w.hmode[c] == 0 && (w.status[c] == 0 || w.status[c] == 3) && i - v0 < w.ntree[c]) {
        const VInfo vi = w.vinfo[i];
        n = vi.nins & (VI_KIDS - 1);
        if (n > 0) cls = (n <= LEAF_MAX_INS && !(vi.nins & VI_KIDS)) ? 2 : 1;
        own = cls == 1 ? (int32_t)(i - v0) : vi.ppos;
    }
    w.op_cnt[i] = cls == 1 ? n : 0;
    w.leaf_need[i] = cls == 2 ? 32 * n : 0;  // (32 per insert is the most a spine can copy)
    w.chain_flag[i] = cls == 1;
    w.owner[i] = own;
}
AA_HDN void f_owner_jump(const Ws &w, int64_t i) {  // one round; in place (a racing read sees another ancestor on the same path)
    const int32_t o = w.owner[i];
    if (o < 0) return;
    const int64_t c = upper_idx(w.vtx_off, w.C, i);
    const int64_t v0 = w.vtx_off[c];
    if (w.chain_flag[v0 + o]) return;
    w.owner[i] = w.owner[v0 + o];
}
AA_HDN void f_chain_slot(const Ws &w, int64_t i) {
    if (w.chain_flag[i]) w.chain_slot[w.chain_ord[i]] = (int32_t)i;
}
// slot where the leaf run in front of chain ordinal j (global) of the contig starting at v0 begins
AA_HD int64_t run_start_slot(const Ws &w, int64_t v0, int64_t gord) { return gord > w.chain_ord[v0] ? (int64_t)w.chain_slot[gord - 1] : v0; }
AA_HDN void f_ops_fill(const Ws &w, int64_t i) {
    if (!w.chain_flag[i]) return;
    const int64_t c = upper_idx(w.vtx_off, w.C, i);
    const int64_t v0 = w.vtx_off[c];
    const VInfo vi = w.vinfo[i];
    const int64_t o = w.op_off[i];
    const int32_t n = vi.nins & (VI_KIDS - 1);
    int32_t src = -1;
    if (vi.ppos >= 0) {
        const int32_t po = w.owner[v0 + vi.ppos];
        if (po >= 0) src = (int32_t)(w.chain_ord[v0 + po] - w.chain_ord[v0]);
    }
    // ids of the leaves between the previous chain vertex and this one (BFS order): reserved in front of this vertex's nodes
    const int64_t resv = w.leaf_lp[i] - w.leaf_lp[run_start_slot(w, v0, w.chain_ord[i])];
    for (int32_t k = 0; k < n; k++) {
        const InsKey ik = w.ins[(int64_t)vi.ins_beg + k];
        HOp op;
        op.sum = ik.sum;
        op.anom = ik.anom;
        op.nz = ik.nz;
        op.tot = ik.tot;
        op.eid = ik.eid;
        op.ctl = (k == 0 ? (HOP_FIRST | (int32_t)(resv << HOP_RESV_SHIFT)) : 0) | (k == n - 1 ? HOP_LAST : 0);
        op.aux = k == 0 ? src : 0;
        w.ops[o + k] = op;
    }
}
AA_HD bool other_group(const Ws &w, int64_t c) { return w.grp_sel >= 0 && w.grp != nullptr && w.grp[c] != w.grp_sel; }
AA_HDN void f_root_fill(const Ws &w, int64_t i) {  // after the serial warp: every tree vertex takes the heap of its owner
    const int64_t c = upper_idx(w.vtx_off, w.C, i);
    if (other_group(w, c)) return;
    const int64_t v0 = w.vtx_off[c];
    if (w.hmode[c] != 0 || w.status[c] != 0 || i - v0 >= w.ntree[c]) return;
    const int32_t o = w.owner[i];
    const int32_t root = o < 0 ? -1 : w.chain_root[w.chain_ord[v0 + o]];
    w.root_at[i] = root;
    w.hroot[v0 + w.vinfo[i].x] = root;
    if (w.leaf_need[i] > 0) {  // a leaf: its ids lie in the run reserved in front of the next chain vertex (or behind the last one)
        const int64_t gord = w.chain_ord[i];
        w.leaf_base[i] = w.resv_base[gord + c] + (int32_t)(w.leaf_lp[i] - w.leaf_lp[run_start_slot(w, v0, gord)]);
    }
}

// insert (key, eid) into the persistent heap rooted at a; returns the new root (or -1 on arena overflow).
// leftist_heap.hpp:29-40 made iterative: walk down the right spine while a->key < k (ties stop: the new key
// goes on top), then copy the spine bottom-up.  Nodes are immutable, 32 B, moved with 128-bit accesses.
AA_HDN int32_t heap_insert(HNode *__restrict__ hn, int32_t *__restrict__ hn_eid, const Ws &w, HeapAlloc &ha, int32_t a,
                           const SKey &k, int32_t eid) {
    HNode spine[48];
    int32_t spine_id[48];
    int32_t depth = 0, arank = 0;
    while (a >= 0) {
        const HNode an = hn_load(hn + a);
        arank = an.rank;
        // a->key < k ? (paf_data.hpp:142-159, CALC_SUM_MODE)
        bool lt;
        if (an.sum != k.sum) lt = an.sum < k.sum;
        else if (an.anom != k.anom) lt = an.anom < k.anom;
        else lt = (int64_t)an.nz * den(k.tot) > (int64_t)k.nz * den(an.tot);
        if (!lt) break;
        spine[depth] = an;
        spine_id[depth] = a;
        depth++;
        a = an.right;
    }
    int32_t r = heap_new(w, ha);
    if (r < 0) return -1;
    HNode nn;
    nn.sum = k.sum;
    nn.anom = k.anom;
    nn.nz = k.nz;
    nn.tot = k.tot;
    nn.left = a;
    nn.right = -1;
    nn.rank = 1;
    nn.lrank = (int16_t)(a >= 0 ? arank : 0);
    hn_store(hn + r, nn);
    hn_eid[r] = eid;
    int32_t rrank = 1;
    while (depth > 0) {
        --depth;
        HNode o = spine[depth];
        int32_t l = o.left, lr = o.lrank, rr = r, rk = rrank;
        if (l < 0 || lr < rk) {  // swap so that the left rank is not smaller
            int32_t t = l;
            l = rr;
            rr = t;
            t = lr;
            lr = rk;
            rk = t;
        }
        const int32_t id = heap_new(w, ha);
        if (id < 0) return -1;
        o.left = l;
        o.right = rr;
        o.lrank = (int16_t)lr;
        o.rank = (int16_t)(rr >= 0 ? rk + 1 : 0);
        hn_store(hn + id, o);
        hn_eid[id] = hn_eid[spine_id[depth]];
        r = id;
        rrank = o.rank;
    }
    return r;
}

#if defined(__CUDACC__)
__device__ void f_heaps_level(const Ws &w, int64_t slot);  // level-parallel builder, defined with the warp kernels below
#endif
// phase: sidetrack heaps (k_shortest_walks.hpp:191-215): the tree vertices in BFS order, each inserting its sidetracks into the
// heap it inherits from its tree parent.
// Sequential forms of the flat builder (host emulation): the operation stream of one contig, then one leaf.
AA_HD bool heap_reserve_host(const Ws &w, HeapAlloc &ha, int64_t n, int32_t &base) {  // n contiguous ids, not counted as used
    if (ha.cur + n > ha.end) {
        const int64_t chunks = (n + HEAP_CHUNK - 1) / HEAP_CHUNK;
        const unsigned long long at = *w.heap_top;
        *w.heap_top = at + (unsigned long long)(chunks * HEAP_CHUNK);
        if ((int64_t)at + chunks * HEAP_CHUNK > w.Hcap) {
            ha.overflow = true;
            return false;
        }
        ha.cur = (int64_t)at;
        ha.end = ha.cur + chunks * HEAP_CHUNK;
    }
    base = (int32_t)ha.cur;
    ha.cur += n;
    return true;
}
AA_HDN void f_heaps_ops(const Ws &w, int64_t c) {
    if (aa_lane() != 0) return;
    if (w.status[c] != 0 && w.status[c] != 3) return;
    if (w.hmode[c] != 0) return;
    const int64_t v0 = w.vtx_off[c];
    const int32_t nt = w.ntree[c];
    const HOp *ops = w.ops + w.op_off[v0];
    const int64_t nops = w.op_off[v0 + nt] - w.op_off[v0];
    const int64_t cbase = w.chain_ord[v0];
    int32_t *chain_root = w.chain_root + cbase;
    int32_t *resv_base = w.resv_base + cbase + c;
    HeapAlloc ha;
    ha.cur = ha.end = 0;
    ha.used = 0;
    ha.overflow = false;
    int32_t root = -1, cv = 0;
    for (int64_t i = 0; i < nops && !ha.overflow; i++) {
        const HOp op = ops[i];
        if (op.ctl & HOP_FIRST) {
            root = op.aux < 0 ? -1 : chain_root[op.aux];
            const int64_t resv = op.ctl >> HOP_RESV_SHIFT;
            if (resv > 0 && !heap_reserve_host(w, ha, resv, resv_base[cv])) break;
        }
        SKey sk;
        sk.sum = op.sum;
        sk.anom = op.anom;
        sk.nz = op.nz;
        sk.tot = op.tot;
        sk.use = 1;
        root = heap_insert(w.hn, w.hn_eid, w, ha, root, sk, op.eid);
        if (root < 0) break;
        if (op.ctl & HOP_LAST) chain_root[cv++] = root;
    }
    if (!ha.overflow) {  // the leaves behind the last chain vertex
        const int64_t tail = w.leaf_lp[v0 + nt] - w.leaf_lp[run_start_slot(w, v0, cbase + cv)];
        if (tail > 0) heap_reserve_host(w, ha, tail, resv_base[cv]);
    }
    w.heap_used[c] = ha.used;
    w.status[c] = ha.overflow ? 3 : 0;
}
AA_HDN void f_heaps_leaf_host(const Ws &w, int64_t slot) {  // a leaf of a streaming-mode contig, ids from its reservation
    if (aa_lane() != 0) return;
    const int64_t c = upper_idx(w.vtx_off, w.C, slot);
    if (w.status[c] != 0) return;
    const int64_t v0 = w.vtx_off[c];
    const VInfo vi = w.vinfo[slot];
    const int32_t n = vi.nins & (VI_KIDS - 1);
    int32_t root = vi.ppos < 0 ? -1 : w.root_at[v0 + vi.ppos];
    HeapAlloc ha;
    ha.cur = w.leaf_base[slot];
    ha.end = ha.cur + 32 * (int64_t)n;
    ha.used = 0;
    ha.overflow = false;
    for (int32_t k = 0; k < n; k++) {
        const InsKey ik = w.ins[(int64_t)vi.ins_beg + k];
        SKey sk;
        sk.sum = ik.sum;
        sk.anom = ik.anom;
        sk.nz = ik.nz;
        sk.tot = ik.tot;
        sk.use = 1;
        root = heap_insert(w.hn, w.hn_eid, w, ha, root, sk, ik.eid);
    }
    w.root_at[slot] = root;
    w.hroot[v0 + vi.x] = root;
    w.heap_used[c] += ha.used;
}

#if defined(__CUDA_ARCH__)
// ---- warp-cooperative Kahn passes (device only; the host emulation runs f_relax / f_topo) ----------------------
// One warp per contig.  The FIFO is kept in a shared-memory ring (plus global memory when it overflows);
// the edges of the popped vertex are spread over the lanes; vertices that become ready are appended in
// lane order, which is the list order the reference pushes them in (k_shortest_walks.hpp:149-153).
constexpr int32_t KRING = 1024;
struct KahnSmem {
    int32_t ring[KRING];
};
__device__ __forceinline__ int32_t kahn_pop(const KahnSmem &sm, const int32_t *q, int32_t head, int32_t tail) {
    return (tail - head <= KRING) ? sm.ring[head & (KRING - 1)] : q[head];
}
// ordered append of the lanes with `ready` (lane order); returns the new tail
__device__ __forceinline__ int32_t kahn_push(KahnSmem &sm, int32_t *q, int32_t tail, bool ready, int32_t x) {
    const uint32_t m = __ballot_sync(0xffffffffu, ready);
    if (ready) {
        const int32_t pos = tail + __popc(m & ((1u << (threadIdx.x & 31)) - 1u));
        q[pos] = x;
        sm.ring[pos & (KRING - 1)] = x;
    }
    __syncwarp();
    return tail + __popc(m);
}
// reverse Kahn + relaxation + min-anom DP (k_shortest_walks.hpp:132-175; paf_data.cpp:705-713)
// The FIFO ring carries everything a pop needs (final distance of the vertex, its in-edge range), so a pop reads
// shared memory only; the in-edge records of every touched source are pulled into L1 ahead of its own pop.
constexpr int32_t RRING = 256;
struct RelaxSmem {
    int64_t sum[RRING];
    int32_t v[RRING], anom[RRING], nz[RRING], tot[RRING], amin_reach[RRING], deg[RRING];
    uint32_t ra[RRING];
};
static_assert(sizeof(RelaxSmem) <= 10 * 1024, "RelaxSmem does not fit its shared-memory allotment");
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ void f_relax_warp(const Ws &w, int64_t c, void *scratch) {
    RelaxSmem &sm = *reinterpret_cast<RelaxSmem *>(scratch);
    const uint32_t FULL = 0xffffffffu;
    const int32_t lane = (int32_t)(threadIdx.x & 31);
    if (w.status[c] != 0) return;
    if (w.rmode[c] >= 0) return;  // level-synchronous pass
    Ctg g = ctg_view(w, c);
    const int64_t v0 = g.v0;
    VState *__restrict__ vs = w.vs + v0;
    const RevRec *__restrict__ rrec = w.rrec;
    const int64_t *__restrict__ rev_off = w.rev_off + v0;
    int32_t *__restrict__ q = w.queue + v0;
    int32_t head = 0, tail = 0;
    // ordered append of the lanes with `ready`; the entry carries the vertex's final state and in-edge range
    auto push = [&](bool ready, int32_t x, const VState &sx, int64_t xa, int64_t xb) {
        const uint32_t m = __ballot_sync(FULL, ready);
        if (ready) {
            const int32_t pos = tail + __popc(m & ((1u << lane) - 1u));
            const int32_t s = pos & (RRING - 1);
            q[pos] = x;
            sm.v[s] = x;
            sm.sum[s] = sx.sum;
            sm.anom[s] = sx.anom;
            sm.nz[s] = sx.nz;
            sm.tot[s] = sx.tot;
            sm.amin_reach[s] = sx.amin_reach;
            sm.ra[s] = (uint32_t)xa;
            sm.deg[s] = (int32_t)(xb - xa);
        }
        __syncwarp();
        tail += __popc(m);
    };
    // seeds: vertices without out-edges, ascending id (k_shortest_walks.hpp:139-141)
    for (int32_t vb = 0; vb < g.V; vb += 32) {
        const int32_t v = vb + lane;
        VState sx;
        sx.sum = 0;
        sx.anom = sx.nz = sx.tot = sx.best = 0;
        sx.cnt = 1;
        sx.amin_reach = 0;
        int64_t xa = 0, xb = 0;
        if (v < g.V) {
            sx = vs[v];
            xa = rev_off[v];
            xb = rev_off[v + 1];
        }
        push(v < g.V && sx.cnt == 0, v, sx, xa, xb);
    }
    while (head < tail) {
        VState sv;
        int32_t v;
        int64_t ra, rb;
        if (tail - head <= RRING) {
            const int32_t s = head & (RRING - 1);
            v = sm.v[s];
            sv.sum = sm.sum[s];
            sv.anom = sm.anom[s];
            sv.nz = sm.nz[s];
            sv.tot = sm.tot[s];
            sv.amin_reach = sm.amin_reach[s];
            ra = sm.ra[s];
            rb = ra + sm.deg[s];
        } else {  // the ring wrapped over this entry
            v = q[head];
            sv = vs[v];
            ra = rev_off[v];
            rb = rev_off[v + 1];
        }
        head++;
        const bool vreach = (sv.amin_reach & 1) != 0;
        const int32_t av = sv.amin_reach >> 1;
        for (int64_t kb = ra; kb < rb; kb += 32) {
            const int64_t k = kb + lane;
            bool ready = false;
            int32_t x = 0;
            VState sx;
            sx.sum = 0;
            sx.anom = sx.nz = sx.tot = sx.best = sx.cnt = sx.amin_reach = 0;
            int64_t xa = 0, xb = 0;
            if (k < rb) {
                const RevRec r = rrec[k];
                x = r.src;
                sx = vs[x];
                xa = rev_off[x];
                xb = rev_off[x + 1];
                if (xb > xa) prefetch_l1(rrec + xa);  // x's own pop will start from these records
                if (vreach) {
                    D4 cand, cur;
                    cand.sum = sv.sum + r.sum;
                    cand.anom = sv.anom + (int32_t)(r.fl & 3u);
                    cand.nz = sv.nz + (int32_t)((r.fl >> 2) & 1u);
                    cand.tot = sv.tot + (int32_t)((r.fl >> 3) & 1u);
                    cur.sum = sx.sum;
                    cur.anom = sx.anom;
                    cur.nz = sx.nz;
                    cur.tot = sx.tot;
                    int32_t am = sx.amin_reach >> 1;
                    const int32_t na = av + (int32_t)(r.fl & 3u);
                    if (!(sx.amin_reach & 1) || less4(cand, cur)) {  // strict: the first relaxer wins among equals
                        sx.sum = cand.sum;
                        sx.anom = cand.anom;
                        sx.nz = cand.nz;
                        sx.tot = cand.tot;
                        sx.best = v;
                    }
                    if (na < am) am = na;
                    sx.amin_reach = (am << 1) | 1;
                }
                sx.cnt -= 1;
                vs[x] = sx;
                ready = sx.cnt == 0;
            }
            push(ready, x, sx, xa, xb);
        }
    }
    if (lane == 0) {
        const VState ss = vs[g.src];
        w.anom_dis[c] = ss.amin_reach >> 1;
        if (!(ss.amin_reach & 1)) w.status[c] = 2;
    }
}
// ---- relax with the open-vertex state on chip --------------------------------------------------------------------
// A vertex is *open* from the first time one of its out-neighbours is popped until its own count reaches zero.  Its
// state is revisited once per out-edge, and a store to global memory invalidates the line in L1, so keeping the open
// states in HBM costs one L2 round trip per pop on the critical path.  In a chain-like contig the open vertices are
// the few vertices of the parts next to the frontier, and vertex ids are local (singles in sorted block order, pair
// vertices in (i, j) discovery order), so two direct-mapped shared-memory tables (singles / pairs and src / dest,
// indexed by id mod RC_SLOTS) hold them without conflicts.  A first touch initialises the entry from the in-edge
// record itself (RevRec carries the source's initial count), a vertex whose count reaches zero is written to vs[]
// once, with its final state, and leaves the table.  Two open vertices that map to one slot make the warp give up:
// the contig is flagged (status 4) and redone by the global-memory form above (f_relax_redo_warp).
// The first 32 in-edge records of the next queue entry are loaded one pop ahead whenever the queue holds one.
constexpr int32_t RC_SLOTS = AA_RC_SLOTS;  // per table
struct __attribute__((aligned(16))) RCEnt {
    int64_t sum;
    int32_t anom, nz;
    int32_t tot, best, cnt, amin_reach;
    uint32_t ra;
    int32_t deg;
    int32_t tag;  // vertex held by the slot, -1: free
    int32_t pad;
};
struct RelaxSmemC {
    RelaxSmem ring;
    RCEnt tab[2 * RC_SLOTS];
};
static_assert(sizeof(RelaxSmemC) <= 10 * 1024 + 2 * RC_SLOTS * 48, "RelaxSmemC does not fit its shared-memory allotment");
__device__ __forceinline__ RevRec rrec_ld(const RevRec *p) {
    const int4 t = __ldg(reinterpret_cast<const int4 *>(p));
    RevRec r;
    r.sum = (int64_t)(((uint64_t)(uint32_t)t.y << 32) | (uint32_t)t.x);
    r.src = t.z;
    r.fl = (uint32_t)t.w;
    return r;
}
struct SegRun {
    int32_t pops;  // vertices this run finalised
    bool clash;    // two open vertices on one table slot (or a saturated count): the run is void
    bool tie;      // two distances tied on (sum, anom) with different qul counts: the choice depends on the shift
};
// One run of the reverse Kahn + relax over the vertices between two articulation blocks lo < hi (lo = -1: down to src,
// hi = g.n: from dest).  A block that is a part of its own is an articulation vertex of the DAG: edges leave a part only
// for the single vertices of the next part (paf_data.cpp:653-695), so every walk from an earlier part to dest passes
// through it, the FIFO holds nothing else when it is popped, and what happens below it depends on what is above only
// through its own distance, which shifts every distance below by the same amount.  Shifts do not change a comparison
// of (sum, anom); they can change the mapq-ratio tie-break (paf_data.hpp:152-158), so a run in the local frame reports
// whether it ever got that far (tie) and the sweep redoes those runs with the true distance of hi as the seed.
// seed_scan: whole contig in one run, seeds = every vertex without out-edges (k_shortest_walks.hpp:139-141).
__device__ SegRun relax_segment(const Ws &w, const Ctg &g, int32_t lo, int32_t hi, bool seed_scan, const VState &seed_state,
                                RelaxSmemC &smc, int32_t *__restrict__ q, int64_t rec_bucket) {
    RelaxSmem &sm = smc.ring;
    const uint32_t FULL = 0xffffffffu;
    const int32_t lane = (int32_t)(threadIdx.x & 31);
    const int64_t v0 = g.v0;
    VState *__restrict__ vs = w.vs + v0;
    const RevRec *__restrict__ rrec = w.rrec;
    const int64_t *__restrict__ rev_off = w.rev_off + v0;
    int32_t head = 0, tail = 0;
    for (int32_t i = lane; i < 2 * RC_SLOTS; i += 32) smc.tab[i].tag = -1;
    __syncwarp();
    auto push = [&](bool ready, int32_t x, const VState &sx, uint32_t xa, int32_t xdeg, bool to_q) {
        const uint32_t m = __ballot_sync(FULL, ready);
        if (ready) {
            const int32_t pos = tail + __popc(m & ((1u << lane) - 1u));
            const int32_t s = pos & (RRING - 1);
            if (to_q) q[pos] = x;
            sm.v[s] = x;
            sm.sum[s] = sx.sum;
            sm.anom[s] = sx.anom;
            sm.nz[s] = sx.nz;
            sm.tot[s] = sx.tot;
            sm.amin_reach[s] = sx.amin_reach;
            sm.ra[s] = xa;
            sm.deg[s] = xdeg;
        }
        __syncwarp();
        tail += __popc(m);
    };
    const bool top = hi == g.n;
    if (seed_scan) {
        for (int32_t vb = 0; vb < g.V; vb += 32) {
            const int32_t v = vb + lane;
            VState sx;
            sx.sum = 0;
            sx.anom = sx.nz = sx.tot = sx.best = 0;
            sx.cnt = 1;
            sx.amin_reach = 0;
            int64_t xa = 0, xb = 0;
            if (v < g.V) {
                sx = vs[v];
                xa = rev_off[v];
                xb = rev_off[v + 1];
            }
            push(v < g.V && sx.cnt == 0, v, sx, (uint32_t)xa, (int32_t)(xb - xa), true);
        }
    } else {  // the vertex above the segment: dest, or the articulation vertex hi with the given state
        const int32_t v = top ? g.dest : hi;
        VState sx = seed_state;
        if (top) sx = vs[v];
        const int64_t xa = rev_off[v], xb = rev_off[v + 1];
        push(lane == 0, v, sx, (uint32_t)xa, (int32_t)(xb - xa), top);
    }
    RevRec pre;
    pre.sum = 0;
    pre.src = 0;
    pre.fl = 0;
    int32_t pre_for = -1;  // queue position whose first 32 records are in `pre`
    bool clash = false, tie = false;
    while (head < tail) {
        VState sv;
        int32_t v;
        int64_t ra, rb;
        if (tail - head <= RRING) {
            const int32_t s = head & (RRING - 1);
            v = sm.v[s];
            sv.sum = sm.sum[s];
            sv.anom = sm.anom[s];
            sv.nz = sm.nz[s];
            sv.tot = sm.tot[s];
            sv.amin_reach = sm.amin_reach[s];
            ra = sm.ra[s];
            rb = ra + sm.deg[s];
        } else {  // the ring wrapped over this entry: its final state is in vs[]
            v = q[head];
            sv = vs[v];
            ra = rev_off[v];
            rb = rev_off[v + 1];
        }
        const bool have = pre_for == head;
        const RevRec r0 = pre;
        head++;
        if (head < tail && tail - head <= RRING) {  // records of the next pop, one pop ahead
            const int32_t s2 = head & (RRING - 1);
            const uint32_t nra = sm.ra[s2];
            if (lane < sm.deg[s2] && sm.v[s2] != lo) pre = rrec_ld(rrec + nra + lane);
            pre_for = head;
        }
        if (v == lo) continue;  // the lower articulation vertex: its in-edges belong to the next segment
        const bool vreach = (sv.amin_reach & 1) != 0;
        const int32_t av = sv.amin_reach >> 1;
        for (int64_t kb = ra; kb < rb; kb += 32) {
            const int64_t k = kb + lane;
            bool ready = false, bad = false;
            int32_t x = 0, xdeg = 0;
            uint32_t xa = 0;
            VState sx;
            sx.sum = 0;
            sx.anom = sx.nz = sx.tot = sx.best = sx.cnt = sx.amin_reach = 0;
            if (k < rb) {
                const RevRec r = (have && kb == ra) ? r0 : rrec_ld(rrec + k);
                x = r.src;
                RCEnt &e = smc.tab[(x < g.n ? 0 : RC_SLOTS) + (x & (RC_SLOTS - 1))];
                const int32_t old = atomicCAS(&e.tag, -1, x);
                if (old == x) {
                    sx.sum = e.sum;
                    sx.anom = e.anom;
                    sx.nz = e.nz;
                    sx.tot = e.tot;
                    sx.best = e.best;
                    sx.cnt = e.cnt;
                    sx.amin_reach = e.amin_reach;
                    xa = e.ra;
                    xdeg = e.deg;
                } else if (old < 0 && (r.fl >> 4) != RR_DEG_SAT) {  // first touch: the state relax_init gave it
                    sx.best = -1;
                    sx.cnt = (int32_t)(r.fl >> 4);
                    sx.amin_reach = 0x3fffffff << 1;
                    const int64_t a = __ldg(rev_off + x), b = __ldg(rev_off + x + 1);
                    xa = (uint32_t)a;
                    xdeg = (int32_t)(b - a);
                } else {
                    bad = true;
                }
                if (!bad) {
                    if (vreach) {
                        D4 cand, cur;
                        cand.sum = sv.sum + r.sum;
                        cand.anom = sv.anom + (int32_t)(r.fl & 3u);
                        cand.nz = sv.nz + (int32_t)((r.fl >> 2) & 1u);
                        cand.tot = sv.tot + (int32_t)((r.fl >> 3) & 1u);
                        cur.sum = sx.sum;
                        cur.anom = sx.anom;
                        cur.nz = sx.nz;
                        cur.tot = sx.tot;
                        int32_t am = sx.amin_reach >> 1;
                        const int32_t na = av + (int32_t)(r.fl & 3u);
                        const bool lt = !(sx.amin_reach & 1) || less4(cand, cur);
                        if ((sx.amin_reach & 1) && cand.sum == cur.sum && cand.anom == cur.anom && (cand.nz != cur.nz || cand.tot != cur.tot)) {
                            tie = true;
                            if (rec_bucket >= 0) {  // the outcome as a condition on the counts of the seed
                                const int32_t slot = atomicAdd(w.seg_ncon + rec_bucket, 1);
                                if (slot < SEG_MAXCON) {
                                    const int64_t anz = cand.nz - seed_state.nz, atot = cand.tot - seed_state.tot;
                                    const int64_t bnz = cur.nz - seed_state.nz, btot = cur.tot - seed_state.tot;
                                    SegCon sc;
                                    sc.c0 = anz * btot - bnz * atot;
                                    sc.c1 = (int32_t)(anz - bnz);
                                    sc.c2 = (int32_t)(btot - atot);
                                    sc.out = lt ? 1 : 0;
                                    sc.pad = 0;
                                    w.seg_con[rec_bucket * SEG_MAXCON + slot] = sc;
                                }
                            }
                        }
                        if (lt) {  // strict: the first relaxer wins among equals
                            sx.sum = cand.sum;
                            sx.anom = cand.anom;
                            sx.nz = cand.nz;
                            sx.tot = cand.tot;
                            sx.best = v;
                        }
                        if (na < am) am = na;
                        sx.amin_reach = (am << 1) | 1;
                    }
                    sx.cnt -= 1;
                    ready = sx.cnt == 0;
                    if (ready) {
                        vs[x] = sx;  // final
                        e.tag = -1;
                    } else {
                        e.sum = sx.sum;
                        e.anom = sx.anom;
                        e.nz = sx.nz;
                        e.tot = sx.tot;
                        e.best = sx.best;
                        e.cnt = sx.cnt;
                        e.amin_reach = sx.amin_reach;
                        e.ra = xa;
                        e.deg = xdeg;
                    }
                }
            }
            if (__any_sync(FULL, bad)) {
                clash = true;
                break;
            }
            push(ready, x, sx, xa, xdeg, true);
        }
        if (clash) break;
    }
    SegRun out;
    out.pops = tail - ((seed_scan || top) ? 0 : 1);
    out.clash = clash;
    out.tie = __any_sync(FULL, tie);
    return out;
}
// the segment of one bucket: [its boundary, the next boundary of the contig)
struct SegSpan {
    int32_t lo, hi;     // articulation blocks (lo = -1: bottom segment, hi = n: top segment)
    int32_t expect;     // vertices the run must finalise
    int64_t qoff;       // its slice of the scratch queue
    bool whole;         // the contig has no boundary
};
__device__ SegSpan seg_span(const Ws &w, const Ctg &g, int64_t bo, int64_t nb, int64_t m) {
    SegSpan sp;
    sp.lo = m == 0 ? -1 : w.seg_bnd[bo + m];
    sp.hi = g.n;
    for (int64_t k = m + 1; k < nb; k++) {
        const int32_t b = w.seg_bnd[bo + k];
        if (b >= 0) {
            sp.hi = b;
            break;
        }
    }
    const int32_t l0 = sp.lo < 0 ? 0 : sp.lo;
    const int64_t pl = w.pair_beg[g.b0 + l0] - g.p0, ph = w.pair_beg[g.b0 + sp.hi] - g.p0;
    sp.expect = (sp.hi - l0) + (int32_t)(ph - pl) + (sp.lo < 0 ? 1 : 0) + (sp.hi == g.n ? 1 : 0);
    sp.qoff = l0 + pl + (sp.lo >= 0 ? 1 : 0);
    sp.whole = sp.lo < 0 && sp.hi == g.n;
    return sp;
}
// first articulation block (a block that is a part of its own) of every bucket but the first of its contig
__device__ void f_seg_bounds_warp(const Ws &w, int64_t bk) {
    const int32_t lane = (int32_t)(threadIdx.x & 31);
    const int64_t c = upper_idx(w.seg_boff, w.C, bk);
    const int64_t m = bk - w.seg_boff[c];
    int32_t res = -1;
    if (m > 0 && w.status[c] == 0 && w.rmode[c] < 0) {
        const int64_t b0 = w.ctg_off[c];
        const int32_t n = (int32_t)(w.ctg_off[c + 1] - b0);
        const int32_t end = (int32_t)((m + 1) * SEG_BLOCKS) < n ? (int32_t)((m + 1) * SEG_BLOCKS) : n;
        for (int32_t base = (int32_t)(m * SEG_BLOCKS); base < end && res < 0; base += 32) {
            const int32_t i = base + lane;
            const bool art = i < end && w.part_r[b0 + i] - w.part_l[b0 + i] == 1;
            const uint32_t mk = __ballot_sync(0xffffffffu, art);
            if (mk) res = base + (__ffs(mk) - 1);
        }
    }
    if (lane == 0) {
        w.seg_bnd[bk] = res;
        w.seg_ncon[bk] = 0;
        if (m == 0) w.seg_mode[c] = (w.status[c] == 0 && w.rmode[c] < 0) ? 1 : 0;  // singletons keep their initial states
    }
}
// pass 1: every segment in the local frame (distance of its upper articulation vertex := 0)
__device__ void f_relax_seg_warp(const Ws &w, int64_t bk, void *scratch) {
    RelaxSmemC &smc = *reinterpret_cast<RelaxSmemC *>(scratch);
    const int32_t lane = (int32_t)(threadIdx.x & 31);
    const int64_t c = upper_idx(w.seg_boff, w.C, bk);
    if (w.status[c] != 0 && w.status[c] != 4) return;
    if (w.rmode[c] >= 0) return;  // level-synchronous pass
    const int64_t bo = w.seg_boff[c], m = bk - bo;
    if (m > 0 && w.seg_bnd[bk] < 0) return;  // this bucket starts no segment
    const Ctg g = ctg_view(w, c);
    const SegSpan sp = seg_span(w, g, bo, w.seg_boff[c + 1] - bo, m);
    // The run assumes plausible qul counts for its upper articulation vertex (about 5/8 hop per block above it, 5/7 of
    // them with mapq != 0), so that most ratio tie-breaks come out as they will with the true counts; every one of them
    // is recorded and checked by the sweep.
    const bool local = !(sp.whole || sp.hi == g.n);
    VState seed;
    seed.sum = 0;
    seed.anom = 0;
    seed.tot = local ? 1 + (int32_t)(((int64_t)(g.n - sp.hi) * 5) >> 3) : 0;
    seed.nz = local ? (int32_t)(((int64_t)seed.tot * 5) / 7) : 0;
    if (local && w.seg_guess) {
        seed.tot = 1;
        seed.nz = 0;
    }
    seed.best = -1;
    seed.cnt = 0;
    seed.amin_reach = 1;
    const SegRun run = relax_segment(w, g, sp.lo, sp.hi, sp.whole, seed, smc, w.queue + g.v0 + sp.qoff - (local ? 1 : 0), local ? bk : -1);
    if (lane == 0) {
        if (run.clash || run.pops != sp.expect) w.status[c] = 4;  // redo the contig in one piece with the states in global memory
        w.seg_seed[2 * bk] = seed.nz;
        w.seg_seed[2 * bk + 1] = seed.tot;
    }
}
// pass 2, one warp per contig: from dest down, hand every segment the distance of its upper articulation vertex; a
// segment whose recorded tie-breaks do not all stand with the true counts is redone with that distance as the seed (its
// states are then global already)
__device__ void f_relax_sweep_warp(const Ws &w, int64_t c, void *scratch) {
    RelaxSmemC &smc = *reinterpret_cast<RelaxSmemC *>(scratch);
    const int32_t lane = (int32_t)(threadIdx.x & 31);
    if (w.status[c] != 0) return;  // singletons; status 4 goes to the redo launch
    if (w.rmode[c] >= 0) return;
    const Ctg g = ctg_view(w, c);
    const int64_t bo = w.seg_boff[c], nb = w.seg_boff[c + 1] - bo;
    SegShift D;
    D.sum = 0;
    D.anom = D.nz = D.tot = 0;
    D.amin_reach = 1;
    D.pad[0] = D.pad[1] = 0;
    SegShift zero = D;
    for (int64_t m = nb - 1; m >= 0; m--) {
        if (m > 0 && w.seg_bnd[bo + m] < 0) continue;
        const SegSpan sp = seg_span(w, g, bo, nb, m);
        SegShift sh = D;  // states of the run + sh = global states
        sh.nz -= w.seg_seed[2 * (bo + m)];
        sh.tot -= w.seg_seed[2 * (bo + m) + 1];
        const int32_t ncon = sp.hi != g.n ? w.seg_ncon[bo + m] : 0;
        bool stands = true;
        if (ncon > 0 && (D.amin_reach & 1)) {  // do the recorded tie-breaks come out the same with the true counts?
            stands = ncon <= SEG_MAXCON && D.tot >= 1 && w.seg_guess < 2;
            for (int32_t k = lane; stands && k < ncon; k += 32) {
                const SegCon sc = w.seg_con[(bo + m) * SEG_MAXCON + k];
                const int64_t L = sc.c0 + (int64_t)D.tot * sc.c1 + (int64_t)D.nz * sc.c2;
                if ((L > 0) != (sc.out != 0)) stands = false;
            }
            stands = __all_sync(0xffffffffu, stands);
        }
        if (!stands) {
            VState seed;
            seed.sum = D.sum;
            seed.anom = D.anom;
            seed.nz = D.nz;
            seed.tot = D.tot;
            seed.best = -1;
            seed.cnt = 0;
            seed.amin_reach = D.amin_reach;
            const SegRun run = relax_segment(w, g, sp.lo, sp.hi, false, seed, smc, w.queue + g.v0 + sp.qoff - 1, -1);
            if (run.clash || run.pops != sp.expect) {
                if (lane == 0) w.status[c] = 4;
                return;
            }
            sh = zero;
            __threadfence_block();
            __syncwarp();
        }
        if (lane == 0) w.seg_shift[bo + m] = sh;
        if (sp.lo >= 0) {
            const VState s = seg_compose(w.vs[g.v0 + sp.lo], sh);
            D.sum = s.sum;
            D.anom = s.anom;
            D.nz = s.nz;
            D.tot = s.tot;
            D.amin_reach = (s.amin_reach & 1) ? s.amin_reach : 0;
        }
    }
    if (lane == 0) {
        const VState ss = seg_compose(w.vs[g.v0 + g.src], w.seg_shift[bo]);
        w.anom_dis[c] = ss.amin_reach >> 1;
        if (!(ss.amin_reach & 1)) w.status[c] = 2;
    }
}
// second launch: the contigs the on-chip form gave up on (status 4) start over from the initial states
__device__ void f_relax_redo_warp(const Ws &w, int64_t c, void *scratch) {
    const int32_t lane = (int32_t)(threadIdx.x & 31);
    if (w.status[c] != 4) return;
    __syncwarp();
    Ctg g = ctg_view(w, c);
    for (int32_t v = lane; v < g.V; v += 32) {
        const int64_t gv = g.v0 + v;
        VState s;
        s.sum = 0;
        s.anom = s.nz = s.tot = 0;
        s.best = -1;
        s.cnt = (int32_t)(w.eoff[gv + 1] - w.eoff[gv]);
        s.amin_reach = v == g.dest ? 1 : (0x3fffffff << 1);
        w.vs[gv] = s;
    }
    if (lane == 0) {
        w.status[c] = 0;
        w.seg_mode[c] = 0;  // vs[] will hold global states
    }
    __threadfence_block();
    __syncwarp();
    f_relax_warp(w, c, scratch);
}
// forward Kahn order (paf_data.cpp:742-746)
__device__ void f_topo_warp(const Ws &w, int64_t c, void *scratch) {
    KahnSmem &sm = *reinterpret_cast<KahnSmem *>(scratch);
    const int32_t lane = (int32_t)(threadIdx.x & 31);
    if (w.status[c] == 1) return;  // singleton; (unsolvable contigs are not known yet: relax runs concurrently)
    if (w.rmode[c] >= 0) return;   // level-synchronous pass
    Ctg g = ctg_view(w, c);
    const int64_t v0 = g.v0;
    int32_t *__restrict__ cnt = w.cnt2 + v0;
    const int64_t *__restrict__ eoff = w.eoff + v0;
    const Edge *__restrict__ edge = w.edge;
    int32_t *__restrict__ q = w.topo + v0;
    int32_t *__restrict__ order = w.order + v0;
    int32_t head = 0, tail = 0;
    for (int32_t vb = 0; vb < g.V; vb += 32) {
        const int32_t v = vb + lane;
        const bool seed = v < g.V && cnt[v] == 0;
        tail = kahn_push(sm, q, tail, seed, v);
    }
    while (head < tail) {
        const int32_t u = kahn_pop(sm, q, head, tail);
        if (lane == 0) order[u] = head;
        head++;
        const int64_t ea = eoff[u], eb = eoff[u + 1];
        for (int64_t kb = ea; kb < eb; kb += 32) {
            const int64_t k = kb + lane;
            bool ready = false;
            int32_t x = 0;
            if (k < eb) {
                x = e_dst(edge[k]);
                const int32_t left = cnt[x] - 1;  // out-edges of u have distinct heads: no conflict inside the warp
                cnt[x] = left;
                ready = left == 0;
            }
            tail = kahn_push(sm, q, tail, ready, x);
        }
    }
}
// forward Kahn order, one warp per segment: an articulation block is alone in the FIFO when it is popped in this direction
// as well (every earlier vertex has a walk to it, so all of them are popped before it becomes ready), which makes the order
// below and above it independent; positions are local pop numbers + the number of vertices before the segment
__device__ void f_topo_seg_warp(const Ws &w, int64_t bk, void *scratch) {
    KahnSmem &sm = *reinterpret_cast<KahnSmem *>(scratch);
    const int32_t lane = (int32_t)(threadIdx.x & 31);
    const int64_t c = upper_idx(w.seg_boff, w.C, bk);
    if (w.status[c] == 1) return;
    if (w.rmode[c] >= 0) return;  // level-synchronous pass
    const int64_t bo = w.seg_boff[c], m = bk - bo;
    if (m > 0 && w.seg_bnd[bk] < 0) return;  // this bucket starts no segment
    const Ctg g = ctg_view(w, c);
    const SegSpan sp = seg_span(w, g, bo, w.seg_boff[c + 1] - bo, m);
    const int64_t v0 = g.v0;
    const int32_t base = sp.lo < 0 ? 0 : (int32_t)sp.qoff;  // src, the blocks before lo and their pair vertices
    int32_t *__restrict__ cnt = w.cnt2 + v0;
    const int64_t *__restrict__ eoff = w.eoff + v0;
    const Edge *__restrict__ edge = w.edge;
    int32_t *__restrict__ q = w.topo + v0 + base;
    int32_t *__restrict__ order = w.order + v0;
    const int32_t stop = sp.hi == g.n ? -1 : sp.hi;  // popped here, numbered and expanded by the segment it starts
    int32_t head = 0, tail = 0;
    if (sp.whole) {
        for (int32_t vb = 0; vb < g.V; vb += 32) {
            const int32_t v = vb + lane;
            tail = kahn_push(sm, q, tail, v < g.V && cnt[v] == 0, v);
        }
    } else {
        tail = kahn_push(sm, q, tail, lane == 0, sp.lo < 0 ? g.src : sp.lo);
    }
    while (head < tail) {
        const int32_t u = kahn_pop(sm, q, head, tail);
        if (u == stop) {
            head++;
            continue;
        }
        if (lane == 0) order[u] = base + head;
        head++;
        const int64_t ea = eoff[u], eb = eoff[u + 1];
        for (int64_t kb = ea; kb < eb; kb += 32) {
            const int64_t k = kb + lane;
            bool ready = false;
            int32_t x = 0;
            if (k < eb) {
                x = e_dst(edge[k]);
                const int32_t left = cnt[x] - 1;  // out-edges of u have distinct heads: no conflict inside the warp
                cnt[x] = left;
                ready = left == 0;
            }
            tail = kahn_push(sm, q, tail, ready, x);
        }
    }
    if (lane == 0 && tail - (stop >= 0 ? 1 : 0) != sp.expect) w.topo_redo[c] = 1;
}
__device__ void f_topo_redo_warp(const Ws &w, int64_t c, void *scratch) {
    if (!w.topo_redo[c]) return;
    const int32_t lane = (int32_t)(threadIdx.x & 31);
    const Ctg g = ctg_view(w, c);
    for (int32_t v = lane; v < g.V; v += 32) w.cnt2[g.v0 + v] = (int32_t)(w.rev_off[g.v0 + v + 1] - w.rev_off[g.v0 + v]);
    __threadfence_block();
    __syncwarp();
    f_topo_warp(w, c, scratch);
}
#endif
AA_HDN void f_topo_seg_any(const Ws &w, int64_t bk, void *scratch) {
#if defined(__CUDA_ARCH__)
    f_topo_seg_warp(w, bk, scratch);
#else
    (void)w;
    (void)bk;
    (void)scratch;
#endif
}
AA_HDN void f_topo_redo_any(const Ws &w, int64_t c, void *scratch) {
#if defined(__CUDA_ARCH__)
    f_topo_redo_warp(w, c, scratch);
#else
    (void)w;
    (void)c;
    (void)scratch;
#endif
}
AA_HDN void f_relax_any(const Ws &w, int64_t c, void *scratch) {  // host emulation: the sequential form, whole contig
#if defined(__CUDA_ARCH__)
    f_relax_warp(w, c, scratch);
#else
    (void)scratch;
    f_relax(w, c);
#endif
}
AA_HDN void f_seg_bounds_any(const Ws &w, int64_t bk) {
#if defined(__CUDA_ARCH__)
    f_seg_bounds_warp(w, bk);
#else
    (void)w;
    (void)bk;
#endif
}
AA_HDN void f_relax_seg_any(const Ws &w, int64_t bk, void *scratch) {
#if defined(__CUDA_ARCH__)
    f_relax_seg_warp(w, bk, scratch);
#else
    (void)w;
    (void)bk;
    (void)scratch;
#endif
}
AA_HDN void f_relax_sweep_any(const Ws &w, int64_t c, void *scratch) {
#if defined(__CUDA_ARCH__)
    f_relax_sweep_warp(w, c, scratch);
#else
    (void)w;
    (void)c;
    (void)scratch;
#endif
}
AA_HDN void f_relax_redo_any(const Ws &w, int64_t c, void *scratch) {
#if defined(__CUDA_ARCH__)
    f_relax_redo_warp(w, c, scratch);
#else
    (void)w;
    (void)c;
    (void)scratch;
#endif
}
AA_HDN void f_topo_any(const Ws &w, int64_t c, void *scratch) {
#if defined(__CUDA_ARCH__)
    f_topo_warp(w, c, scratch);
#else
    (void)scratch;
    f_topo(w, c);
#endif
}
constexpr size_t KAHN_SMEM_BYTES = 4 * 1024;
constexpr size_t RELAX_SMEM_BYTES = 10 * 1024;        // global-state form (redo launch)
constexpr size_t RELAX_SMEM_C_BYTES = 10 * 1024 + (size_t)2 * AA_RC_SLOTS * 48;  // ring + open-vertex tables

#if defined(__CUDA_ARCH__)
// ---- warp-cooperative sidetrack heaps (device only; the host emulation runs f_heaps above) --------------------
// One warp per contig streams two flat arrays prepared in parallel: the tree vertices in BFS order (vinfo) and
// their inserts (ins).  Nothing on the serial chain chases the tree: both streams are read 32 records at a time,
// one batch ahead.  The known prefix of the current heap's right spine lives in REGISTERS, spine level i on
// lane i, so that one insert (leftist_heap.hpp:29-40) is a handful of warp-wide steps:
//   descent      = one key compare per lane + ballot (the next unknown spine node is loaded ahead of need)
//   rank update  = a suffix scan over the lanes: level q maps the rank r of its new right child to
//                  f_q(r) = left ? min(left.rank, r) + 1 : 0, i.e. min(a, r + b); such maps compose as
//                  (min(a2, a1 + b2), b1 + b2), so log2(p) shuffle rounds give every level its incoming rank
//   path copying = every lane writes its own new node (ids are consecutive: N first, then bottom-up, the
//                  allocation order of leftist_heap.hpp:29-40, which is what the PQ tie-break sees)
// Spines of finished vertices with children are kept in shared memory (keyed by root id): BFS visits siblings
// and then their children, all of which start from a heap built a few vertices earlier.
constexpr int32_t SPMAX = 32;
constexpr int32_t NSAVE = 8;
struct __attribute__((aligned(8))) IdEid {
    int32_t id, eid;
};
__device__ __forceinline__ bool key_lt(const HNode &an, const InsKey &k) {  // a->key < k (paf_data.hpp:142-159)
    if (an.sum != k.sum) return an.sum < k.sum;
    if (an.anom != k.anom) return an.anom < k.anom;
    return (int64_t)an.nz * den(k.tot) > (int64_t)k.nz * den(an.tot);
}
__device__ __forceinline__ InsKey ins_ld(const InsKey *p) {  // 24 B, 8-byte aligned: three 64-bit loads
    union {
        InsKey k;
        int64_t v[3];
    } u;
    const int64_t *q = reinterpret_cast<const int64_t *>(p);
    u.v[0] = q[0];
    u.v[1] = q[1];
    u.v[2] = q[2];
    return u.k;
}
__device__ __forceinline__ VInfo vinfo_ld(const VInfo *p) {
    union {
        VInfo v;
        V16 q;
    } u;
    u.q = *reinterpret_cast<const V16 *>(p);
    return u.v;
}
// ---- the working spine of one warp and the two halves of an insert (shared by the two heap builders) --------
struct Spine {
    HNode nd;        // spine level `lane` (valid below L)
    int32_t nd_id, nd_eid;
    HNode nx;        // the first unknown spine node, loaded ahead of need (valid when next >= 0)
    int32_t nx_eid;
    int32_t L, next;
};
__device__ __forceinline__ void spine_reset(Spine &sp, const HNode *__restrict__ hn, const int32_t *__restrict__ hn_eid,
                                            int32_t root) {  // nothing known yet: the spine starts at `root`
    sp.L = 0;
    sp.next = root;
    if (root >= 0) {
        sp.nx = hn_load(hn + root);
        sp.nx_eid = hn_eid[root];
    }
}
// descent: first spine level whose key is not < k (-1: the spine would exceed the 32 lanes)
__device__ __forceinline__ int32_t spine_descend(Spine &sp, const HNode *__restrict__ hn, const int32_t *__restrict__ hn_eid,
                                                 const InsKey &k) {
    const uint32_t FULL = 0xffffffffu;
    const int32_t lane = (int32_t)(threadIdx.x & 31);
    for (;;) {
        // the sum decides unless some level ties on it (then the full PafDistance order is evaluated)
        const bool known = lane < sp.L;
        uint32_t stop = __ballot_sync(FULL, known && sp.nd.sum >= k.sum);
        if (__ballot_sync(FULL, known && sp.nd.sum == k.sum)) stop = __ballot_sync(FULL, known && !key_lt(sp.nd, k));
        if (stop) return __ffs(stop) - 1;
        if (sp.next < 0) return sp.L;
        if (sp.L >= SPMAX - 1) return -1;  // cannot happen below 2^31 nodes per heap; fail loudly rather than corrupt
        if (lane == sp.L) {
            sp.nd = sp.nx;
            sp.nd_id = sp.next;
            sp.nd_eid = sp.nx_eid;
        }
        sp.L++;
        sp.next = sp.nx.right;
        if (sp.next >= 0) {
            sp.nx = hn_load(hn + sp.next);  // same address on every lane: one transaction
            sp.nx_eid = hn_eid[sp.next];
        }
    }
}
// path copy: N = nbase, the copy of level q = nbase + (p - q); returns the new root
template <bool KEYS>
__device__ __forceinline__ int32_t spine_apply(Spine &sp, HNode *__restrict__ hn, int32_t *__restrict__ hn_eid,
                                               unsigned long long *__restrict__ hn_key, const InsKey &k, int32_t p, int32_t nbase,
                                               unsigned long long keybase /* owner slot << 32 | nodes it has so far */) {
    const uint32_t FULL = 0xffffffffu;
    const int32_t lane = (int32_t)(threadIdx.x & 31);
    const int32_t BIG = 1 << 24;
    // ranks: suffix scan of the per-level maps r -> min(a, r + b) over levels < p
    int32_t fa = BIG, fb = 0;
    if (lane < p) {
        fa = sp.nd.left < 0 ? 0 : (int32_t)sp.nd.lrank + 1;
        fb = sp.nd.left < 0 ? BIG : 1;
    }
    for (int32_t d = 1; d < p; d <<= 1) {
        const int32_t oa = __shfl_down_sync(FULL, fa, d);
        const int32_t ob = __shfl_down_sync(FULL, fb, d);
        if (lane + d < 32) {
            fa = min(fa, oa + fb);
            fb = fb + ob;
        }
    }
    const int32_t my_rank = min(fa, 1 + fb);           // rank of this level's copy
    int32_t rin = __shfl_down_sync(FULL, my_rank, 1);  // rank of its new right child
    if (lane >= p - 1) rin = 1;                        // level p-1 gets N (rank 1)
    const bool my_swap = lane < p && (sp.nd.left < 0 || (int32_t)sp.nd.lrank < rin);
    const uint32_t swaps = __ballot_sync(FULL, my_swap);
    const int32_t sstar = swaps ? __ffs(swaps) - 1 : -1;
    // the stop node (level p, if any) becomes N's left child
    const int32_t stop_id = __shfl_sync(FULL, sp.nd_id, p & 31);
    const int32_t stop_rank = __shfl_sync(FULL, (int32_t)sp.nd.rank, p & 31);
    if (lane < p) {
        const int32_t ch = nbase + (p - 1 - lane);
        if (my_swap) {
            sp.nd.right = sp.nd.left;
            sp.nd.left = ch;
            sp.nd.lrank = (int16_t)rin;
        } else {
            sp.nd.right = ch;
        }
        sp.nd.rank = (int16_t)my_rank;
        sp.nd_id = nbase + (p - lane);
    } else if (lane == p) {
        sp.nd.sum = k.sum;
        sp.nd.anom = k.anom;
        sp.nd.nz = k.nz;
        sp.nd.tot = k.tot;
        sp.nd.left = p < sp.L ? stop_id : -1;
        sp.nd.right = -1;
        sp.nd.rank = 1;
        sp.nd.lrank = (int16_t)(p < sp.L ? stop_rank : 0);
        sp.nd_id = nbase;
        sp.nd_eid = k.eid;
    }
    if (lane <= p) {
        hn_store(hn + sp.nd_id, sp.nd);
        hn_eid[sp.nd_id] = sp.nd_eid;
        if (KEYS) hn_key[sp.nd_id] = keybase + (unsigned long long)(sp.nd_id - nbase);
    }
    __syncwarp();  // later inserts read these nodes from other lanes
    // new known spine: copies 0..sstar (then the old left of sstar), or copies 0..p-1 and N
    const int32_t nright = __shfl_sync(FULL, sp.nd.right, sstar >= 0 ? sstar : 0);
    const int32_t old_next = sp.next;
    sp.next = sstar >= 0 ? nright : -1;
    sp.L = sstar >= 0 ? sstar + 1 : p + 1;
    if (sp.next >= 0 && sp.next != old_next) {
        sp.nx = hn_load(hn + sp.next);
        sp.nx_eid = hn_eid[sp.next];
    }
    return p > 0 ? nbase + p : nbase;
}
__device__ __forceinline__ InsKey ins_bcast(const InsKey &kreg, int32_t src) {
    const uint32_t FULL = 0xffffffffu;
    InsKey k;
    k.sum = __shfl_sync(FULL, kreg.sum, src);
    k.anom = __shfl_sync(FULL, kreg.anom, src);
    k.nz = __shfl_sync(FULL, kreg.nz, src);
    k.tot = __shfl_sync(FULL, kreg.tot, src);
    k.eid = __shfl_sync(FULL, kreg.eid, src);
    return k;
}
// ---- the serial heap builder, flat form ------------------------------------------------------------------------------
// One warp per contig walks the contig's operation stream (f_ops_fill).  Same node ids, same nodes as f_heaps_warp; what is
// different is the length of the instruction chain of one insert (a lone warp issues a dependent instruction every 5-6
// cycles, so the chain length IS the time):
//   * no vertex loop: vertices without inserts never reach the warp, parents are resolved by the pre-pass;
//   * ranks.  With a_j = (left_j ? lrank_j : -1) + j + 1 for the copied levels j < p and a_p = p + 1 for the new node, the copy
//     of level i gets rank min_{j in [i,p]} a_j - i, and level i swaps iff a_i < min_{j in (i,p]} a_j (leftist_heap.hpp:34-38
//     unrolled).  The topmost swap m, which alone decides the next spine, is the LAST position of the minimum of a over
//     [0,p]: one redux.min over (a << 5 | 31 - lane).  Levels above m take rank amin - i; only the levels between m and p
//     (usually none or one) need a suffix scan;
//   * the whole right spine is always in registers (level i on lane i): after a swap the spine of the old left child is
//     read at once, from a direct-mapped shared-memory cache of the nodes this warp created (tag = node id; a global store
//     invalidates the line in L1, so re-reading one's own nodes from memory costs an L2 round trip per level);
//   * operations come through a shared-memory ring filled by cp.async three chunks ahead, and the next operation is read
//     while the current one is worked on.
constexpr int32_t OPRING = 128;
struct ChainSmem {
    HNode snode[NSAVE][SPMAX];
    IdEid sid[NSAVE][SPMAX];
    int32_t sord[NSAVE], sL[NSAVE];
    HOp oring[OPRING];
    // at HEAP2_FIXED_BYTES: HNode cnode[1 << bits]; IdEid ctag[1 << bits]
};
static_assert(sizeof(ChainSmem) <= HEAP2_FIXED_BYTES, "ChainSmem does not fit its shared-memory allotment");
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
#ifdef AA_HEAP_TIMERS
#define HT_DECL long long ht_t = clock64(), ht_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long ht_n[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define HT(i) do { const long long ht_now = clock64(); ht_acc[i] += ht_now - ht_t; ht_n[i]++; ht_t = ht_now; } while (0)
#define HT_COUNT(i) ht_n[i]++
#else
#define HT_DECL
#define HT(i)
#define HT_COUNT(i)
#endif
__device__ void f_heaps_chain(const Ws &w, int64_t c, void *scratch) {
    HT_DECL
    ChainSmem &sm = *reinterpret_cast<ChainSmem *>(scratch);
    const uint32_t FULL = 0xffffffffu;
    const int32_t lane = (int32_t)(threadIdx.x & 31);
    if (w.status[c] != 0 && w.status[c] != 3) return;
    if (w.hmode[c] != 0) return;  // a shallow, wide tree: built level by level (f_heaps_level)
    const int64_t v0 = w.vtx_off[c];
    const int32_t nt = w.ntree[c];
    HNode *__restrict__ hn = w.hn;
    int32_t *__restrict__ hn_eid = w.hn_eid;
    // node cache
    HNode *cnode = nullptr;
    IdEid *ctag = nullptr;
    int32_t cmask = -1;
    if (w.heap_cache_bits > 0) {
        unsigned char *base = reinterpret_cast<unsigned char *>(scratch) + HEAP2_FIXED_BYTES;
        const int32_t n = 1 << w.heap_cache_bits;
        cnode = reinterpret_cast<HNode *>(base);
        ctag = reinterpret_cast<IdEid *>(base + (size_t)n * sizeof(HNode));
        cmask = n - 1;
        for (int32_t i = lane; i < n; i += 32) ctag[i].id = -1;
    }
    if (lane < NSAVE) sm.sord[lane] = -2;
    // ---- operation stream ----
    const HOp *__restrict__ ops = w.ops + w.op_off[v0];
    const int32_t nops = (int32_t)(w.op_off[v0 + nt] - w.op_off[v0]);
    int32_t *__restrict__ chain_root = w.chain_root + w.chain_ord[v0];
    int32_t *__restrict__ resv_base = w.resv_base + w.chain_ord[v0] + c;
    auto issue_chunk = [&](int32_t ch) {  // ops [32 ch, 32 ch + 32) -> ring (one cp.async group, possibly empty)
        const int32_t i = ch * 32 + lane;
        if (i < nops) {
            const V16 *src = reinterpret_cast<const V16 *>(ops + i);
            V16 *dst = reinterpret_cast<V16 *>(&sm.oring[i & (OPRING - 1)]);
            cp_async16(dst, src);
            cp_async16(dst + 1, src + 1);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    issue_chunk(0);
    issue_chunk(1);
    issue_chunk(2);
    asm volatile("cp.async.wait_group 2;" ::: "memory");
    __syncwarp();
    // ---- the working spine: level `lane` of the right spine of heap `cur_ord` (complete: L levels) ----
    int64_t ksum = 0;
    int32_t kanom = 0, knz = 0, ktot = 0, left = -1, lrank = 0, rank = 0, id = -1, eid = 0;
    int32_t L = 0;
    int32_t cur_ord = -1;  // chain ordinal of the vertex whose heap the working spine is (-1: the empty heap)
    int32_t cv = 0;        // chain vertices finished
    int32_t cur = 0, end = 0;
    int64_t used = 0;
    bool overflow = false;
    int32_t save_at = 0;
    // a fresh arena chunk (warp-uniform result; -1: arena exhausted)
    auto new_chunk = [&](int32_t need_ids) -> int32_t {  // enough whole chunks for need_ids contiguous ids; sets `end`
        const int32_t chunks = (need_ids + HEAP_CHUNK - 1) / HEAP_CHUNK;
        unsigned long long at = 0;
        if (lane == 0) at = atomicAdd(w.heap_top, (unsigned long long)chunks * HEAP_CHUNK);
        at = __shfl_sync(FULL, at, 0);
        if ((int64_t)at + (int64_t)chunks * HEAP_CHUNK > w.Hcap) return -1;
        if (w.chunk_ctg)
            for (int32_t k = lane; k < chunks * (HEAP_CHUNK / 64); k += 32) w.chunk_ctg[(at >> 6) + k] = (int32_t)c;
        end = (int32_t)__reduce_max_sync(FULL, (uint32_t)at) + chunks * HEAP_CHUNK;
        return end - chunks * HEAP_CHUNK;
    };
    // spine of the heap rooted at `from` appended below level L0 (complete right spine, read through the cache)
    auto extend = [&](int32_t L0, int32_t from) -> int32_t {
        int32_t nx = from, l = L0;
        while (nx >= 0) {
            if (l >= SPMAX - 1) return -1;
            int32_t nright;
            bool hit = false;
            if (cmask >= 0) {
                const int32_t s = nx & cmask;
                const IdEid t = ctag[s];
                nright = cnode[s].right;
                hit = t.id == nx;
                if (hit && lane == l) {  // only the lane of this level takes the record
                    const HNode n = hn_load(&cnode[s]);
                    ksum = n.sum;
                    kanom = n.anom;
                    knz = n.nz;
                    ktot = n.tot;
                    left = n.left;
                    lrank = n.lrank;
                    rank = n.rank;
                    id = nx;
                    eid = t.eid;
                }
            }
            if (!hit) {
                const HNode n = hn_load(hn + nx);
                const int32_t ne = hn_eid[nx];
                nright = n.right;
                if (lane == l) {
                    ksum = n.sum;
                    kanom = n.anom;
                    knz = n.nz;
                    ktot = n.tot;
                    left = n.left;
                    lrank = n.lrank;
                    rank = n.rank;
                    id = nx;
                    eid = ne;
                }
            }
            l++;
            nx = nright;
            HT_COUNT(7);
        }
        return l;
    };
    const uint32_t oring_s = (uint32_t)__cvta_generic_to_shared(&sm.oring[0]);
    auto op_read = [&](int32_t i) -> HOp {
        union {
            HOp o;
            unsigned long long v[4];
        } u;
        const uint32_t a = oring_s + (uint32_t)((i & (OPRING - 1)) * (int32_t)sizeof(HOp));
        asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(u.v[0]), "=l"(u.v[1]) : "r"(a) : "memory");
        asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(u.v[2]), "=l"(u.v[3]) : "r"(a + 16) : "memory");
        return u.o;
    };
    HOp op = op_read(0);
    int32_t oi = 0;
    while (oi < nops) {
        HT(0);
        // ---- descent: first level whose key is not < k (leftist_heap.hpp:30).  The sum decides unless a level ties on
        // it; ties and the control flags of the operation take the slow path below ----
        bool known = lane < L;
        uint32_t stop = __ballot_sync(FULL, known && ksum >= op.sum);
        const uint32_t slow = __ballot_sync(FULL, (known && ksum == op.sum) || ((op.ctl & HOP_FIRST) != 0 && op.aux != cur_ord));
        if (slow) {
            if ((op.ctl & HOP_FIRST) && op.aux != cur_ord) {  // this vertex does not continue the heap just built: switch spines
                if (cur_ord >= 0) {
                    const uint32_t have = __ballot_sync(FULL, lane < NSAVE && sm.sord[lane] == cur_ord);
                    if (!have) {
                        const int32_t sl = save_at;
                        save_at = (save_at + 1) % NSAVE;
                        if (lane < L) {
                            HNode n;
                            n.sum = ksum;
                            n.anom = kanom;
                            n.nz = knz;
                            n.tot = ktot;
                            n.left = left;
                            n.right = 0;
                            n.rank = (int16_t)rank;
                            n.lrank = (int16_t)lrank;
                            hn_store(&sm.snode[sl][lane], n);
                            IdEid ie;
                            ie.id = id;
                            ie.eid = eid;
                            sm.sid[sl][lane] = ie;
                        }
                        if (lane == 0) {
                            sm.sord[sl] = cur_ord;
                            sm.sL[sl] = L;
                        }
                        __syncwarp();
                    }
                }
                const uint32_t hit = __ballot_sync(FULL, lane < NSAVE && sm.sord[lane] == op.aux);
                if (op.aux >= 0 && hit) {
                    const int32_t sl = __ffs(hit) - 1;
                    L = sm.sL[sl];
                    if (lane < L) {
                        const HNode n = hn_load(&sm.snode[sl][lane]);
                        const IdEid ie = sm.sid[sl][lane];
                        ksum = n.sum;
                        kanom = n.anom;
                        knz = n.nz;
                        ktot = n.tot;
                        left = n.left;
                        rank = n.rank;
                        lrank = n.lrank;
                        id = ie.id;
                        eid = ie.eid;
                    }
                } else {
                    L = extend(0, op.aux >= 0 ? __ldcg(chain_root + op.aux) : -1);
                    if (L < 0) {
                        overflow = true;
                        break;
                    }
                }
                cur_ord = op.aux;
                known = lane < L;
            }
            bool lt;  // a->key < k (paf_data.hpp:142-159)
            if (ksum != op.sum) lt = ksum < op.sum;
            else if (kanom != op.anom) lt = kanom < op.anom;
            else lt = (int64_t)knz * den(op.tot) > (int64_t)op.nz * den(ktot);
            stop = __ballot_sync(FULL, known && !lt);
            HT(2);
        }
        const int32_t p = stop ? __ffs(stop) - 1 : L;
        HT(3);
        // ids: the leaves in front of this vertex take theirs first (FIRST carries the count), then the p + 1 nodes of the insert
        const int32_t resv = (op.ctl >> HOP_RESV_SHIFT);
        const int32_t need = p + 1;
        if (p >= SPMAX - 1 || cur + resv + need > end) {
            if (p >= SPMAX - 1) {  // cannot happen below 2^31 nodes per heap; fail loudly rather than corrupt
                overflow = true;
                break;
            }
            cur = new_chunk(resv + need);
            if (cur < 0) {
                overflow = true;
                break;
            }
        }
        if (lane == 0 && resv > 0) resv_base[cv] = cur;
        cur += resv;
        const int32_t nbase = cur;
        cur += need;
        used += need;
        // ---- ranks: the levels that swap are the strict suffix minima of a over [0, p]; the topmost one, m, is the last
        // position of the minimum ----
        const int32_t BIG = 1 << 20;
        const int32_t a = lane < p ? (left < 0 ? lane : lrank + lane + 1) : (lane == p ? p + 1 : BIG);
        const uint32_t enc = ((uint32_t)a << 5) | (uint32_t)(31 - lane);
        const uint32_t mn = __reduce_min_sync(FULL, enc);
        const int32_t amin = (int32_t)(mn >> 5);
        const int32_t m = 31 - (int32_t)(mn & 31u);
        int32_t smin = amin;   // min of a over [lane, p] (lanes <= m; set below for the others)
        int32_t snext = amin;  // min of a over (lane, p]
        for (int32_t mi = m; mi < p;) {  // suffix minima below the topmost swap: one redux each (usually one or two)
            const uint32_t r = __reduce_min_sync(FULL, lane > mi ? enc : 0xffffffffu);
            const int32_t am = (int32_t)(r >> 5);
            const int32_t mj = 31 - (int32_t)(r & 31u);
            smin = (lane > mi && lane <= mj) ? am : smin;
            snext = (lane >= mi && lane < mj) ? am : snext;
            mi = mj;
        }
        int32_t spill = -1;  // old left child of level m: its spine follows level m
        if (m < p) spill = __shfl_sync(FULL, left, m);
        HT(4);
        // ---- records: N on lane p, the copy of level q on lane q, ids nbase + (p - q) (allocation order) ----
        const bool is_copy = lane < p, is_new = lane == p, has = p < L;
        const bool swp = is_copy && a < snext;  // swap: the new subtree goes left
        const int32_t ch = nbase + (p - 1 - lane);
        const int32_t right = is_copy ? (swp ? left : ch) : -1;
        const int32_t old_id = id, old_rank = rank;
        left = is_copy ? (swp ? ch : left) : (has ? old_id : -1);
        lrank = is_copy ? (swp ? snext - (lane + 1) : lrank) : (has ? old_rank : 0);
        rank = is_copy ? smin - lane : 1;
        id = nbase + (p - lane);
        ksum = is_new ? op.sum : ksum;
        kanom = is_new ? op.anom : kanom;
        knz = is_new ? op.nz : knz;
        ktot = is_new ? op.tot : ktot;
        eid = is_new ? op.eid : eid;
        const bool last = (op.ctl & HOP_LAST) != 0;
        if (lane <= p) {
            HNode n;
            n.sum = ksum;
            n.anom = kanom;
            n.nz = knz;
            n.tot = ktot;
            n.left = left;
            n.right = right;
            n.rank = (int16_t)rank;
            n.lrank = (int16_t)lrank;
            hn_store(hn + id, n);
            hn_eid[id] = eid;
            if (cmask >= 0) {
                const int32_t s = id & cmask;
                hn_store(&cnode[s], n);
                IdEid ie;
                ie.id = id;
                ie.eid = eid;
                ctag[s] = ie;
            }
        }
        // the next operation is read now (its chunk was issued three chunks ahead)
        oi++;
        if ((oi & 31) == 0) {
            issue_chunk((oi >> 5) + 2);
            asm volatile("cp.async.wait_group 2;" ::: "memory");
        }
        __syncwarp();
        op = op_read(oi);
        HT(5);
        if (m < p) {
            L = extend(m + 1, spill);
            if (L < 0) {
                overflow = true;
                break;
            }
        } else {
            L = p + 1;
        }
        if (last) {
            if (lane == 0) chain_root[cv] = nbase + p;
            cur_ord = cv;
            cv++;
        }
        HT(6);
    }
    if (!overflow) {  // the leaves behind the last chain vertex
        const int64_t gord = w.chain_ord[v0] + cv;
        const int32_t tail = (int32_t)(w.leaf_lp[v0 + nt] - w.leaf_lp[run_start_slot(w, v0, gord)]);
        if (tail > 0) {
            if (cur + tail > end) cur = new_chunk(tail);
            if (cur < 0) overflow = true;
            else if (lane == 0) resv_base[cv] = cur;
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#ifdef AA_HEAP_TIMERS
    if (lane == 0 && w.vbase)
        for (int i = 0; i < 8; i++) {
            w.vbase[2 * i] = ht_acc[i];
            w.vbase[2 * i + 1] = ht_n[i];
        }
#endif
    if (lane == 0) {
        w.heap_used[c] = used;
        w.status[c] = overflow ? 3 : 0;
    }
}
// ---- one warp per tree vertex: the builder of (a) shallow, wide trees (dense contigs: depth 4, tens of thousands of
// vertices per level), launched depth by depth, and (b) the leaves of every other tree, which feed no other heap
// and therefore need not wait in the serial chain of f_heaps_warp.  The parent's heap is complete (earlier launch);
// its spine is fetched on demand.  Nodes come from 64-node chunks of the arena.
constexpr int32_t LCHUNK = 64;
__device__ void f_heaps_level(const Ws &w, int64_t slot /* global BFS slot of the vertex */) {
    const uint32_t FULL = 0xffffffffu;
    const int32_t lane = (int32_t)(threadIdx.x & 31);
    const int64_t c = upper_idx(w.vtx_off, w.C, slot);
    if (other_group(w, c)) return;
    const int64_t v0 = w.vtx_off[c];
    HNode *__restrict__ hn = w.hn;
    int32_t *__restrict__ hn_eid = w.hn_eid;
    const VInfo vi = vinfo_ld(w.vinfo + slot);
    int32_t root = vi.ppos < 0 ? -1 : w.root_at[v0 + vi.ppos];
    const int32_t nins = vi.nins & (VI_KIDS - 1);
    int32_t used = 0;
    if (w.hmode[c] == 0 && w.status[c] != 0) return;  // a leaf of a contig the serial builder gave up on (arena overflow: rebuilt)
    bool overflow = *w.lvl_overflow != 0;
    if (nins > 0 && !overflow) {
        const InsKey *__restrict__ ins = w.ins + vi.ins_beg;
        Spine sp;
        sp.nd.sum = sp.nx.sum = 0;
        sp.nd.anom = sp.nd.nz = sp.nd.tot = sp.nx.anom = sp.nx.nz = sp.nx.tot = 0;
        sp.nd.left = sp.nd.right = sp.nx.left = sp.nx.right = -1;
        sp.nd.rank = sp.nd.lrank = sp.nx.rank = sp.nx.lrank = 0;
        sp.nd_id = -1;
        sp.nd_eid = sp.nx_eid = 0;
        spine_reset(sp, hn, hn_eid, root);
        int64_t cur = 0, end = 0;
        if (w.hmode[c] == 0) {  // a leaf of a streaming-mode contig: ids reserved in sequence by f_heaps_warp
            cur = w.leaf_base[slot];
            end = cur + (int64_t)nins * 32;
        }
        InsKey kreg, knext;
        kreg.sum = knext.sum = 0;
        kreg.anom = kreg.nz = kreg.tot = kreg.eid = 0;
        knext.anom = knext.nz = knext.tot = knext.eid = 0;
        if (lane < nins) kreg = ins_ld(ins + lane);
        if (32 + lane < nins) knext = ins_ld(ins + 32 + lane);
        for (int32_t t = 0; t < nins; t++) {
            if (t > 0 && (t & 31) == 0) {
                kreg = knext;
                if (t + 32 + lane < nins) knext = ins_ld(ins + t + 32 + lane);
            }
            const InsKey k = ins_bcast(kreg, t & 31);
            const int32_t p = spine_descend(sp, hn, hn_eid, k);
            if (p < 0) {
                overflow = true;
                break;
            }
            const int32_t need = p + 1;
            if (cur + need > end) {
                unsigned long long at = 0;
                if (lane == 0) at = atomicAdd(w.heap_top, (unsigned long long)LCHUNK);
                at = __shfl_sync(FULL, at, 0);
                if ((int64_t)at + LCHUNK > w.Hcap) {
                    overflow = true;
                    break;
                }
                if (lane == 0 && w.chunk_ctg) w.chunk_ctg[at >> 6] = (int32_t)c;
                cur = (int64_t)at;
                end = cur + LCHUNK;
            }
            const int32_t nbase = (int32_t)cur;
            cur += need;
            root = spine_apply<true>(sp, hn, hn_eid, w.hn_key, k, p, nbase, ((unsigned long long)(uint32_t)slot << 32) | (uint32_t)used);
            used += need;
        }
    }
    if (lane == 0) {
        if (overflow) {
            *w.lvl_overflow = 1;
        } else {
            w.root_at[slot] = root;
            w.hroot[v0 + vi.x] = root;
            w.vcnt[slot] = used;
            if (used) atomicAdd((unsigned long long *)&w.heap_used[c], (unsigned long long)used);
        }
    }
}
#endif
// ---- which vertices f_heaps_level takes from the streaming builder: tree leaves with inserts
AA_HDN void f_leaf_flag(const Ws &w, int64_t i) {
    const int64_t c = upper_idx(w.vtx_off, w.C, i);
    int32_t f = 0;
    if (w.hmode[c] == 0 && (w.status[c] == 0 || w.status[c] == 3) && i - w.vtx_off[c] < w.ntree[c]) {
        const int32_t nf = w.vinfo[i].nins;
        f = (nf & (VI_KIDS - 1)) > 0 && (nf & (VI_KIDS - 1)) <= 32 && !(nf & VI_KIDS);
    }
    w.leaf_flag[i] = f;
}
AA_HDN void f_leaf_list(const Ws &w, int64_t i) {
    if (w.leaf_flag[i]) w.leaf_list[w.leaf_off[i]] = (uint32_t)i;
}
// ---- after the build: (owner slot, number) -> rank in the sequential allocation order
AA_HDN void f_node_rank(const Ws &w, int64_t id) {
    const unsigned long long key = w.hn_key[id];
    if (key != ~0ull) w.hn_key[id] = (unsigned long long)(w.vbase[key >> 32] + (int64_t)(uint32_t)key);
}
AA_HDN void f_heaps_any(const Ws &w, int64_t c, void *scratch) {  // the serial builder of one streaming-mode contig
#if defined(__CUDA_ARCH__)
    f_heaps_chain(w, c, scratch);
#else
    (void)scratch;
    f_heaps_ops(w, c);
#endif
}
AA_HDN void f_heaps_leaf_any(const Ws &w, int64_t slot) {
#if defined(__CUDA_ARCH__)
    f_heaps_level(w, slot);
#else
    f_heaps_leaf_host(w, slot);
#endif
}
AA_HD size_t heaps_chain_smem_bytes(int bits) {  // f_heaps_chain: fixed part + node cache of 1 << bits entries (tag 8 B + node 32 B)
    return HEAP2_FIXED_BYTES + (bits > 0 ? ((size_t)1 << bits) * 40 : 0);
}

// ---- enumeration priority queue: binary min-heap under the total order (distance, node id, entry index)
AA_HD bool pq_less(const PQEnt &a, const PQEnt &b) {
    if (a.sum != b.sum) return a.sum < b.sum;
    if (a.anom != b.anom) return a.anom < b.anom;
    int64_t x = (int64_t)a.nz * den(b.tot), y = (int64_t)b.nz * den(a.tot);
    if (x != y) return x > y;
    if (a.node != b.node) return a.node < b.node;
    return a.idx < b.idx;
}
AA_HD void pq_push(PQEnt *h, int32_t &n, const PQEnt &e) {
    int32_t i = n++;
    while (i > 0) {
        int32_t p = (i - 1) >> 1;
        PQEnt pe = h[p];
        if (!pq_less(e, pe)) break;
        h[i] = pe;
        i = p;
    }
    h[i] = e;
}
AA_HD PQEnt pq_pop(PQEnt *h, int32_t &n) {
    PQEnt top = h[0];
    PQEnt last = h[--n];
    int32_t i = 0;
    for (;;) {
        int32_t l = 2 * i + 1;
        if (l >= n) break;
        int32_t m = l;
        PQEnt me = h[l];
        if (l + 1 < n) {
            PQEnt re = h[l + 1];
            if (pq_less(re, me)) {
                m = l + 1;
                me = re;
            }
        }
        if (!pq_less(me, last)) break;
        h[i] = me;
        i = m;
    }
    if (n > 0) h[i] = last;
    return top;
}

// phase: enumerate the K shortest walks (k_shortest_walks.hpp:217-251)
AA_HDN void f_enum(const Ws &w, int64_t c) {
    if (aa_lane() != 0) return;
    if (w.status[c] != 0) {
        w.n_walk[c] = 0;
        return;
    }
    Ctg g = ctg_view(w, c);
    const int64_t v0 = g.v0;
    const int64_t e0 = w.eoff[v0];
    const int64_t wo = w.walk_off[c];
    D4 *dist = w.wdist + wo;
    int32_t *last = w.wlast + wo;
    int32_t *en = w.ent_node + 3 * wo;
    int32_t *ep = w.ent_prev + 3 * wo;
    PQEnt *pq = w.pq + 3 * wo;
    const int32_t K = w.K;
    int32_t nd = 0, ne = 0, np = 0;
    D4 ds = w.d[v0 + g.src];
    dist[nd] = ds;
    last[nd] = -1;
    nd++;
    int32_t hs = w.hroot[v0 + g.src];
    if (hs >= 0) {
        HNode hr = w.hn[hs];
        PQEnt e;
        e.sum = ds.sum + hr.sum;
        e.anom = ds.anom + hr.anom;
        e.nz = ds.nz + hr.nz;
        e.tot = ds.tot + hr.tot;
        e.node = hs;
        e.idx = ne;
        e.pad = 0;
        en[ne] = hs;
        ep[ne] = -1;
        ne++;
        pq_push(pq, np, e);
        while (np > 0 && nd < K) {
            PQEnt t = pq_pop(pq, np);
            D4 cd;
            cd.sum = t.sum;
            cd.anom = t.anom;
            cd.nz = t.nz;
            cd.tot = t.tot;
            cd.aux = 0;
            dist[nd] = cd;
            last[nd] = t.idx;
            nd++;
            HNode ch = w.hn[t.node];
            int32_t hv = w.hroot[v0 + e_dst(w.edge[e0 + w.hn_eid[t.node]])];
            int32_t pre = ep[t.idx];
            if (hv >= 0) {
                HNode x = w.hn[hv];
                PQEnt a;
                a.sum = t.sum + x.sum;
                a.anom = t.anom + x.anom;
                a.nz = t.nz + x.nz;
                a.tot = t.tot + x.tot;
                a.node = hv;
                a.idx = ne;
                a.pad = 0;
                en[ne] = hv;
                ep[ne] = t.idx;
                ne++;
                pq_push(pq, np, a);
            }
            if (ch.left >= 0) {
                HNode x = w.hn[ch.left];
                PQEnt a;
                a.sum = t.sum + x.sum - ch.sum;
                a.anom = t.anom + x.anom - ch.anom;
                a.nz = t.nz + x.nz - ch.nz;
                a.tot = t.tot + x.tot - ch.tot;
                a.node = ch.left;
                a.idx = ne;
                a.pad = 0;
                en[ne] = ch.left;
                ep[ne] = pre;
                ne++;
                pq_push(pq, np, a);
            }
            if (ch.right >= 0) {
                HNode x = w.hn[ch.right];
                PQEnt a;
                a.sum = t.sum + x.sum - ch.sum;
                a.anom = t.anom + x.anom - ch.anom;
                a.nz = t.nz + x.nz - ch.nz;
                a.tot = t.tot + x.tot - ch.tot;
                a.node = ch.right;
                a.idx = ne;
                a.pad = 0;
                en[ne] = ch.right;
                ep[ne] = pre;
                ne++;
                pq_push(pq, np, a);
            }
        }
    }
    w.n_walk[c] = nd;
}

// parallel pre-pass over vertices for the warp enumeration: next-heap root + key of every out-edge
AA_HDN void f_enext(const Ws &w, int64_t gv) {
    const int64_t c = upper_idx(w.vtx_off, w.C, gv);
    if (other_group(w, c)) return;
    if (w.status[c] != 0) return;
    const int64_t v0 = w.vtx_off[c];
    const int64_t ea = w.eoff[gv], eb = w.eoff[gv + 1];
    for (int64_t k = ea; k < eb; k++) {
        const int32_t v = e_dst(w.edge[k]);
        const int32_t hv = w.hroot[v0 + v];
        ENext n;
        n.hv = hv;
        n.pad = 0;
        n.hrank = (hv >= 0 && w.hn_key) ? (int32_t)(uint32_t)w.hn_key[hv] : hv;
        n.sum = 0;
        n.anom = n.nz = n.tot = 0;
        if (hv >= 0) {
            const HNode h = hn_load(w.hn + hv);
            n.sum = h.sum;
            n.anom = h.anom;
            n.nz = h.nz;
            n.tot = h.tot;
        }
        w.enext[k] = n;
    }
}
// parallel pass over the arena: the expansion record of every heap node (hn_eid < 0: the id was never allocated)
AA_HDN void f_xrec(const Ws &w, int64_t id) {
    const int32_t eid = w.hn_eid[id];
    if (eid < 0) return;
    const int64_t c = w.chunk_ctg[id >> 6];
    if (c < 0 || other_group(w, c)) return;  // (c < 0: a chunk another group's builder has just taken; its own pass comes later)
    if (w.status[c] != 0) return;
    const bool keyed = w.hmode[c] != 0;
    const HNode ch = hn_load(w.hn + id);
    const ENext x = w.enext[w.eoff[w.vtx_off[c]] + eid];
    XRec r;
    r.left = ch.left;
    r.right = ch.right;
    r.hv = x.hv;
    r.hkey = keyed ? x.hrank : x.hv;
    r.hsum = x.sum;
    r.hanom = x.anom;
    r.hnz = x.nz;
    r.htot = x.tot;
    r.lsum = r.rsum = 0;
    r.lanom = r.ranom = r.lnz = r.ltot = r.rnz = r.rtot = 0;
    r.lkey = r.rkey = -1;
    r.pad0 = r.pad1 = r.pad2 = 0;
    if (ch.left >= 0) {
        const HNode xl = hn_load(w.hn + ch.left);
        r.lsum = xl.sum - ch.sum;
        r.lanom = xl.anom - ch.anom;
        r.lnz = xl.nz - ch.nz;
        r.ltot = xl.tot - ch.tot;
        r.lkey = keyed ? (int32_t)(uint32_t)w.hn_key[ch.left] : ch.left;
    }
    if (ch.right >= 0) {
        const HNode xr = hn_load(w.hn + ch.right);
        r.rsum = xr.sum - ch.sum;
        r.ranom = xr.anom - ch.anom;
        r.rnz = xr.nz - ch.nz;
        r.rtot = xr.tot - ch.tot;
        r.rkey = keyed ? (int32_t)(uint32_t)w.hn_key[ch.right] : ch.right;
    }
    union {
        XRec r;
        V16 v[6];
    } u;
    u.r = r;
    V16 *q = reinterpret_cast<V16 *>(w.xrec + id);
#pragma unroll
    for (int i = 0; i < 6; i++) q[i] = u.v[i];
}
#if defined(__CUDA_ARCH__)
// ---- warp-cooperative K-walk enumeration (device only; the host emulation runs f_enum) ------------------------
// Same pop sequence as the reference's std::priority_queue<tuple<Distance, heap_t*, int64_t>> (total order:
// distance, node allocation order, entry index), produced in BATCHES instead of one pop at a time:
//   * the queue is split by a threshold key T into a FRONT (all entries < T) and an unsorted BACKLOG (all
//     entries >= T; global memory, append only).  The front is a sorted run in shared memory (filled by a
//     refill, never inserted into one by one) plus up to 32 PENDING entries, sorted across the lanes in
//     registers, which take every new entry below T (one ballot + one shuffle-shift per insert);
//   * a round merges the heads of the two (a 64 -> 32 bitonic merge in registers), expands the candidates at
//     once (heap node, next-heap root, left / right child: the dependent loads of up to 32 pops overlap), and
//     commits the longest prefix that the sequential algorithm would pop in this order: candidate j stays
//     valid as long as no successor of candidates 0..j-1 precedes it (successors carry new, larger entry
//     indices, so only (distance, node) decides).  Entry indices of the committed successors are the prefix
//     sums the sequential pushes would give (next-root, left, right per pop: k_shortest_walks.hpp:245-247);
//   * full pending registers are merged into the run in one pass; when the front is empty a new threshold is
//     chosen from a small sorted sample of the backlog and the entries below it are moved over and sorted.
//     Pops are monotone, so an entry crosses from the backlog to the front once;
//   * entries that can no longer be among the K walks (at least K - popped entries are known to be smaller)
//     are dropped at the next refill and, from then on, when they are created.
// Keys are packed for cheap comparisons: the mapq ratio nz/tot (descending) becomes the exact fixed-point
// number RMAX - floor(nz * 2^S / tot), S = 2 * bits(V): two different fractions with denominators < 2^(S/2)
// differ by more than 2^-S.  Contigs too large for that (V >= 2^20) compare the ratio by cross-multiplying.
constexpr int32_t FCAP = AA_FCAP;  // run capacity (entries of 32 B)
constexpr int32_t FMASK = FCAP - 1;
constexpr int32_t FKEEP = FCAP / 2;            // entries that stay when a full run spills its upper part
constexpr int32_t REFILL_ALL = 3 * FCAP / 4;   // a backlog this small is moved as a whole
constexpr int32_t REFILL_TARGET = 3 * FCAP / 8;
#ifndef AA_NEAR
#define AA_NEAR 3072
#endif
constexpr int32_t NEAR_TARGET = AA_NEAR;  // the near backlog after a split (the only part a refill scans)
#ifndef AA_NSAMPLE
#define AA_NSAMPLE 512
#endif
constexpr int32_t NSAMPLE = AA_NSAMPLE < FCAP ? AA_NSAMPLE : FCAP;
struct __attribute__((aligned(16))) QE {
    int64_t sum;
    uint64_t k1;  // anom << (S + 1) | ratio key        (wide mode: anom << 32)
    uint64_t k2;  // node << 32 | entry index
    int32_t nz, tot;
};
static_assert(sizeof(QE) == sizeof(PQEnt), "the backlog reuses the PQEnt array");
struct EnumSmem {
    QE f[FCAP];
    QE stage[4];  // successors of a serial step, written by lanes 0..2 and read back by all lanes
};
__device__ __forceinline__ XRec xrec_ld(const XRec *p) {
    union {
        XRec r;
        V16 v[6];
    } u;
    const V16 *q = reinterpret_cast<const V16 *>(p);
#pragma unroll
    for (int i = 0; i < 6; i++) u.v[i] = q[i];
    return u.r;
}
__device__ __forceinline__ QE qe_ld(const QE *p) {
    union {
        QE e;
        V16 v[2];
    } u;
    const V16 *q = reinterpret_cast<const V16 *>(p);
    u.v[0] = q[0];
    u.v[1] = q[1];
    return u.e;
}
__device__ __forceinline__ void qe_st(QE *p, const QE &e) {
    union {
        QE e;
        V16 v[2];
    } u;
    u.e = e;
    V16 *q = reinterpret_cast<V16 *>(p);
    q[0] = u.v[0];
    q[1] = u.v[1];
}
template <class SH>
__device__ __forceinline__ QE qe_shfl(const QE &e, SH sh) {  // sh(x) = a warp shuffle of one 32/64-bit value
    QE r;
    r.sum = sh(e.sum);
    r.k1 = sh(e.k1);
    r.k2 = sh(e.k2);
    r.nz = sh(e.nz);
    r.tot = sh(e.tot);
    return r;
}
struct ShIdx {
    int32_t src;
    template <class T>
    __device__ __forceinline__ T operator()(T x) const { return __shfl_sync(0xffffffffu, x, src); }
};
struct ShUp {
    int32_t d;
    template <class T>
    __device__ __forceinline__ T operator()(T x) const { return __shfl_up_sync(0xffffffffu, x, d); }
};
struct ShDown {
    int32_t d;
    template <class T>
    __device__ __forceinline__ T operator()(T x) const { return __shfl_down_sync(0xffffffffu, x, d); }
};
struct ShXor {
    int32_t m;
    template <class T>
    __device__ __forceinline__ T operator()(T x) const { return __shfl_xor_sync(0xffffffffu, x, m); }
};
// total order (distance, node, entry index); branch-free apart from the warp-uniform `wide`
__device__ __forceinline__ bool qe_less(const QE &a, const QE &b, bool wide) {
    int32_t t = a.k2 < b.k2;
    if (wide) {
        const int64_t x = (int64_t)a.nz * den(b.tot), y = (int64_t)b.nz * den(a.tot);
        t = (x > y) | ((x == y) & t);
    }
    const int32_t k = (a.k1 < b.k1) | ((a.k1 == b.k1) & t);
    return ((a.sum < b.sum) | ((a.sum == b.sum) & k)) != 0;
}
// (distance, node) order, strict: decides whether a NEW entry precedes an existing one
__device__ __forceinline__ bool qe_dn_less(const QE &a, const QE &b, bool wide) {
    int32_t t = (a.k2 >> 32) < (b.k2 >> 32);
    if (wide) {
        const int64_t x = (int64_t)a.nz * den(b.tot), y = (int64_t)b.nz * den(a.tot);
        t = (x > y) | ((x == y) & t);
    }
    const int32_t k = (a.k1 < b.k1) | ((a.k1 == b.k1) & t);
    return ((a.sum < b.sum) | ((a.sum == b.sum) & k)) != 0;
}
#ifdef AA_ENUM_TIMERS
#define ET_DECL long long et_t = clock64(), et_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long et_n[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define ET(i) do { const long long et_now = clock64(); et_acc[i] += et_now - et_t; et_n[i]++; et_t = et_now; } while (0)
#define RT_DECL long long rt_t = 0, rt_acc[6] = {0, 0, 0, 0, 0, 0};
#define RT0 rt_t = clock64()
#define RT(i) do { const long long rt_now = clock64(); rt_acc[i] += rt_now - rt_t; rt_t = rt_now; } while (0)
#define RT_ADD(i, v) rt_acc[i] += (v)
#else
#define ET_DECL
#define ET(i)
#define RT_DECL
#define RT0
#define RT(i)
#define RT_ADD(i, v)
#endif
// WIDE: contigs of more than 2^20 vertices, whose qul ratio does not fit the packed key (the comparisons then cross-multiply).
// A template parameter because the compiler turns a run-time flag into straight-line code that does the two 64-bit products
// in EVERY comparison (7 % of the kernel's instructions on inputs that never need them).
template <bool WIDE>
__device__ void f_enum_warp_t(const Ws &w, int64_t c, void *scratch) {
    ET_DECL
#ifdef AA_ENUM_TIMERS
    int32_t plat[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long front_sz = 0, front_n = 0;
#endif
    EnumSmem &sm = *reinterpret_cast<EnumSmem *>(scratch);
    const uint32_t FULL = 0xffffffffu;
    const int32_t lane = (int32_t)(threadIdx.x & 31);
    const uint32_t lt = (1u << lane) - 1u;
    if (w.status[c] != 0) {
        if (lane == 0) w.n_walk[c] = 0;
        return;
    }
    Ctg g = ctg_view(w, c);
    const int64_t v0 = g.v0;
    const int64_t e0 = w.eoff[v0];
    const int64_t wo = w.walk_off[c];
    D4 *__restrict__ dist = w.wdist + wo;
    int32_t *__restrict__ last = w.wlast + wo;
    int32_t *__restrict__ en = w.ent_node + 3 * wo;
    int32_t *__restrict__ ep = w.ent_prev + 3 * wo;
    // The backlog is kept in two regions: `back` (near: keys in [T, T2)) and `far` (keys >= T2).  A refill scans the near
    // region only; when it runs empty the two arrays swap roles and the next refill splits the former far region again.
    QE *back = reinterpret_cast<QE *>(w.pq + 3 * wo);
    QE *far = w.pq_far ? reinterpret_cast<QE *>(w.pq_far + 3 * wo) : nullptr;
    const HNode *__restrict__ hn = w.hn;
    const ENext *__restrict__ enext = w.enext + e0;
    const XRec *__restrict__ xrec = w.xrec;  // nullptr: the arena was too large for expansion records
    const int32_t K = w.K;
    // key packing for this contig
    int32_t vb = 1;
    while ((1 << vb) <= g.V) vb++;
    constexpr bool wide = WIDE;
    // the tie-break of the queue is the node's place in the sequential allocation order (what the reference's pointer
    // comparison sees).  Streaming-mode contigs have their node ids in that order (leaves get reserved id
    // ranges); heaps built level by level do not: there the rank recorded per node is compared and ent_node[] keeps the id
    const bool keyed = w.hmode[c] != 0;
    auto okey = [&](int32_t id) -> uint64_t { return (uint64_t)(uint32_t)(keyed ? (int32_t)(uint32_t)w.hn_key[id] : id) << 32; };
    const int32_t S = wide ? 31 : 2 * vb;  // k1 = anom << (S + 1) | ratio key (<= 2^S)
    auto make_k1 = [&](int32_t anom, int32_t nz, int32_t tot) -> uint64_t {
        uint64_t rk = 0;
        if (!wide) {
            // floor(nz * 2^S / tot) without a 64-bit integer division: an fp64 quotient (error < 2^-13 after scaling, since
            // nz <= tot < 2^20 and S <= 40) corrected by one exact integer step
            const uint64_t dn = (uint64_t)(tot ? tot : 1), num = (uint64_t)(uint32_t)nz << S;
            uint64_t q = (uint64_t)(__ddiv_rn((double)nz, (double)dn) * (double)((uint64_t)1 << S));
            const uint64_t prod = q * dn;
            if (prod > num) q--;
            else if (prod + dn <= num) q++;
            rk = ((uint64_t)1 << S) - q;
        }
        return ((uint64_t)(uint32_t)anom << (S + 1)) | rk;
    };
    QE INF;
    INF.sum = I64_MAX;
    INF.k1 = INF.k2 = ~(uint64_t)0;
    INF.nz = 0;
    INF.tot = 1;

    int32_t nd = 1, ne = 0;
    int32_t head = 0, n0 = 0, np = 0, nR = 0;  // run ring [head, head + n0), pending lanes [0, np), backlog [0, nR)
    int32_t nF = 0;                            // far backlog [0, nF)
    bool hasT = false, hasB = false;           // threshold between front and backlog; bound of the useful keys
    bool hasT2 = false;                        // threshold between the near and the far backlog (T <= T2)
    QE T = INF, B = INF, P = INF, T2 = INF;    // P: this lane's pending entry
    // a warp-uniform entry that is not below T joins the backlog
    auto back_put = [&](const QE &e) {
        if (hasT2 && !qe_less(e, T2, wide)) {
            if (lane == 0) qe_st(far + nF, e);
            nF++;
        } else {
            if (lane == 0) qe_st(back + nR, e);
            nR++;
        }
    };
    QE p0u = INF;                              // warp-uniform copy of lane 0's pending entry, valid while p0_ok
    bool p0_ok = false;
    const D4 ds = w.d[v0 + g.src];
    if (lane == 0) {
        dist[0] = ds;
        last[0] = -1;
    }
    const int32_t hs = w.hroot[v0 + g.src];
    if (hs >= 0) {
        const HNode hr = hn_load(hn + hs);
        QE e;
        e.sum = ds.sum + hr.sum;
        e.nz = ds.nz + hr.nz;
        e.tot = ds.tot + hr.tot;
        e.k1 = make_k1(ds.anom + hr.anom, e.nz, e.tot);
        e.k2 = okey(hs);
        if (lane == 0) {
            en[0] = hs;
            ep[0] = -1;
            P = e;
        }
        ne = 1;
        np = 1;
    }
    auto fslot = [&](int32_t i) -> QE * { return &sm.f[(head + i) & FMASK]; };
    // number of run entries below e (per-lane binary search; the run is sorted)
    auto run_lower_bound = [&](const QE &e) -> int32_t {
        int32_t lo = 0, hi = n0;
        while (lo < hi) {
            const int32_t mid = (lo + hi) >> 1;
            if (qe_less(qe_ld(fslot(mid)), e, wide)) lo = mid + 1;
            else hi = mid;
        }
        return lo;
    };
    // merge the pending registers into the run (one pass over the part of the run that has to move)
    auto flush = [&]() {
        if (np == 0) return;
        p0_ok = false;
        if (n0 + np > FCAP) {  // the run spills its upper part to the backlog; the threshold drops to the first spilled key
            const QE t0 = qe_ld(fslot(FKEEP));
            for (int32_t i = FKEEP + lane; i < n0; i += 32) qe_st(back + nR + (i - FKEEP), qe_ld(fslot(i)));
            nR += n0 - FKEEP;
            n0 = FKEEP;
            T = t0;
            hasT = true;
            const uint32_t out = __ballot_sync(FULL, lane < np && !qe_less(P, T, wide));  // a suffix of the pending lanes
            if (out & (1u << lane)) qe_st(back + nR + __popc(out & lt), P);
            nR += __popc(out);
            np -= __popc(out);
            if (lane >= np) P = INF;  // the lanes that went to the backlog are free again
            __syncwarp();
            if (np == 0) return;
        }
        const int32_t pos = lane < np ? run_lower_bound(P) : 0x7fffffff;  // ascending over the lanes
        const int32_t pos_first = __shfl_sync(FULL, pos, 0), pos_last = __shfl_sync(FULL, pos, np - 1);
        // run element i ends up at i + #{pending l : pos_l <= i}
        auto below_cnt = [&](int32_t i) -> int32_t {
            int32_t lo = 0, hi = np;
            for (int32_t s5 = 0; s5 < 5; s5++) {  // binary search over the lanes' pos (uniform trip count)
                const int32_t mid = (lo + hi) >> 1;
                const int32_t pm = __shfl_sync(FULL, pos, mid < 31 ? mid : 31);
                if (lo < hi) {
                    if (pm <= i) lo = mid + 1;
                    else hi = mid;
                }
            }
            const int32_t pm = __shfl_sync(FULL, pos, lo < 31 ? lo : 31);
            if (lo < hi && pm <= i) lo++;
            return lo;
        };
        if (pos_last <= n0 - pos_first) {  // move the lower part down: new head = head - np
            for (int32_t i0 = 0; i0 < pos_last; i0 += 32) {
                const int32_t i = i0 + lane;
                const int32_t cnt = below_cnt(i);
                QE x = INF;
                if (i < pos_last) x = qe_ld(fslot(i));
                __syncwarp();
                if (i < pos_last) qe_st(&sm.f[(head - np + i + cnt) & FMASK], x);
                __syncwarp();
            }
            head = (head - np) & FMASK;
        } else {  // move the upper part up
            for (int32_t i1 = n0; i1 > pos_first; i1 -= 32) {
                const int32_t i = i1 - 1 - lane;
                const int32_t cnt = below_cnt(i);
                QE x = INF;
                if (i >= pos_first) x = qe_ld(fslot(i));
                __syncwarp();
                if (i >= pos_first) qe_st(fslot(i + cnt), x);
                __syncwarp();
            }
        }
        if (lane < np) qe_st(fslot(pos + lane), P);
        n0 += np;
        np = 0;
        P = INF;
        __syncwarp();
    };
    // sorted insert of a warp-uniform entry into the pending registers
    auto pend_insert = [&](const QE &e) {
        if (np == 32) {
            flush();
            if (hasT && !qe_less(e, T, wide)) {  // a spill lowered the threshold below e
                if (lane == 0) qe_st(back + nR, e);
                nR++;
                return;
            }
        }
        const int32_t pos = __popc(__ballot_sync(FULL, lane < np && qe_less(P, e, wide)));
        const QE up = qe_shfl(P, ShUp{1});
        if (lane > pos) P = up;
        if (lane == pos) P = e;
        np++;
        if (pos == 0) {  // (the serial steps keep a warp-uniform copy of the smallest pending entry)
            p0u = e;
            p0_ok = true;
        }
    };
    // bitonic sort of sm.f[0, n) padded to a power of two (run must be empty: head is reset by the caller)
    auto sort_run = [&](int32_t n) {
        int32_t Pw = 2;
        while (Pw < n) Pw <<= 1;
        for (int32_t i = n + lane; i < Pw; i += 32) qe_st(&sm.f[i], INF);
        __syncwarp();
        for (int32_t k = 2; k <= Pw; k <<= 1)
            for (int32_t j = k >> 1; j > 0; j >>= 1) {
                for (int32_t t = lane; t < (Pw >> 1); t += 32) {
                    const int32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                    const int32_t p2 = i | j;
                    const QE A = qe_ld(&sm.f[i]), Bq = qe_ld(&sm.f[p2]);
                    const bool asc = (i & k) == 0;
                    const bool sw = asc ? qe_less(Bq, A, wide) : qe_less(A, Bq, wide);
                    qe_st(&sm.f[i], sw ? Bq : A);
                    qe_st(&sm.f[p2], sw ? A : Bq);
                }
                __syncwarp();
            }
    };
    // the front is empty: pick a threshold, move the backlog entries below it into the run, sort them
    RT_DECL
    auto refill = [&]() {
        head = 0;
        RT0;
        const int32_t remaining = K - nd;
        if (nR == 0 && nF > 0) {  // the near region ran empty: the far region becomes the near one (split again below)
            QE *const t = back;
            back = far;
            far = t;
            nR = nF;
            nF = 0;
            hasT2 = false;
        }
        if (nR <= REFILL_ALL) {
            for (int32_t i = lane; i < nR; i += 32) qe_st(&sm.f[i], qe_ld(back + i));
            n0 = nR;
            nR = 0;
            hasT = nF > 0;  // what is left of the backlog starts at T2
            if (hasT) T = T2;
        } else {
            // sample stride ~ target / 4: the threshold is the 4th sample or so, the sample sort stays small
            int32_t stride = REFILL_TARGET / 4;
            if (nR / stride > NSAMPLE) stride = nR / NSAMPLE;
            const int32_t ns = nR / stride < NSAMPLE ? nR / stride : NSAMPLE;
            int32_t r = REFILL_TARGET / stride;
            if (r > ns - 1) r = ns - 1;
            // a large near region is split in the same pass: keys from the r2-th sample up move to the far region
            bool split = far != nullptr && nR > 2 * NEAR_TARGET;
            int32_t r2 = NEAR_TARGET / stride;
            if (r2 > ns - 1) r2 = ns - 1;
            for (;;) {
                if (r < 1) {
                    // last resort (a sample could not split the backlog): move the single minimum
                    QE best = INF;
                    int32_t bi = -1;
                    for (int32_t i = lane; i < nR; i += 32) {
                        const QE x = qe_ld(back + i);
                        if (bi < 0 || qe_less(x, best, wide)) {
                            best = x;
                            bi = i;
                        }
                    }
                    for (int32_t d = 16; d > 0; d >>= 1) {
                        const QE o = qe_shfl(best, ShDown{d});
                        const int32_t oi = __shfl_down_sync(FULL, bi, d);
                        if (lane + d < 32 && oi >= 0 && (bi < 0 || qe_less(o, best, wide))) {
                            best = o;
                            bi = oi;
                        }
                    }
                    best = qe_shfl(best, ShIdx{0});
                    bi = __shfl_sync(FULL, bi, 0);
                    const QE tail_e = qe_ld(back + nR - 1);
                    __syncwarp();
                    if (lane == 0) {
                        qe_st(back + bi, tail_e);
                        qe_st(&sm.f[0], best);
                    }
                    nR--;
                    n0 = 1;
                    T = best;
                    T.k2 = best.k2 + 1;  // the smallest key above `best`
                    hasT = true;
                    break;
                }
                // threshold = r-th smallest of a strided sample
                for (int32_t i = lane; i < ns; i += 32) qe_st(&sm.f[i], qe_ld(back + (int64_t)i * stride));
                __syncwarp();
                sort_run(ns);
                RT(0);
                RT_ADD(3, ns);
                RT_ADD(4, nR);
                T = qe_ld(&sm.f[r]);
                hasT = true;
                QE T2n = INF;
                if (split) {
                    T2n = qe_ld(&sm.f[r2]);
                    split = r2 > r && qe_less(T, T2n, wide);
                }
                // candidate bound of the useful keys: about 1.25 * remaining entries are below it
                bool tryB = false;
                QE Bc = INF;
                {
                    const int64_t qb = ((int64_t)remaining * 5 / 4) / stride + 2;
                    if (qb < ns) {
                        Bc = qe_ld(&sm.f[qb]);
                        tryB = !hasB || qe_less(Bc, B, wide);
                    }
                }
                __syncwarp();
                // one pass: below T -> run (unsorted for now), dead entries are dropped, the rest is compacted
                int32_t nb = 0, wpos = 0, ncb = 0;
                for (int32_t i0 = 0; i0 < nR; i0 += 128) {  // four chunks of loads in flight
                    QE xs[4];
#pragma unroll
                    for (int32_t u = 0; u < 4; u++)
                        if (i0 + 32 * u + lane < nR) xs[u] = qe_ld(back + i0 + 32 * u + lane);
#pragma unroll
                    for (int32_t u = 0; u < 4; u++) {
                        bool valid = i0 + 32 * u + lane < nR;
                        if (valid && hasB && !qe_less(xs[u], B, wide)) valid = false;  // can never be popped
                        const bool below = valid && qe_less(xs[u], T, wide);
                        const uint32_t mb = __ballot_sync(FULL, below);
                        const int32_t at = nb + __popc(mb & lt);
                        const bool stage = below && at < FCAP;
                        if (tryB) ncb += __popc(__ballot_sync(FULL, valid && qe_less(xs[u], Bc, wide)));
                        if (split) {
                            const bool hi = valid && !below && !qe_less(xs[u], T2n, wide);
                            const uint32_t mh = __ballot_sync(FULL, hi);
                            if (hi) qe_st(far + nF + __popc(mh & lt), xs[u]);
                            nF += __popc(mh);
                            if (hi) valid = false;  // it left the near region
                        }
                        const uint32_t mk = __ballot_sync(FULL, valid && !stage);
                        if (stage) qe_st(&sm.f[at], xs[u]);
                        if (valid && !stage) qe_st(back + wpos + __popc(mk & lt), xs[u]);
                        nb += __popc(mb);
                        wpos += __popc(mk);
                    }
                    __syncwarp();
                }
                RT(1);
                if (split) {  // done once: a retry below works on the near region that is left
                    T2 = T2n;
                    hasT2 = true;
                    split = false;
                }
                if (nb <= FCAP) {
                    n0 = nb;
                    nR = wpos;
                    if (tryB && ncb >= remaining) {
                        B = Bc;
                        hasB = true;
                    }
                    break;
                }
                // too many entries below this threshold: put the staged ones back and try a lower one
                for (int32_t i = lane; i < FCAP; i += 32) qe_st(back + wpos + i, qe_ld(&sm.f[i]));
                nR = wpos + FCAP;
                __syncwarp();
                r >>= 1;
            }
        }
        __syncwarp();
        RT0;
        if (n0 > 1) sort_run(n0);
        __syncwarp();
        RT(2);
        RT_ADD(5, n0);
    };

    int32_t width = 32;  // candidates per round: follows the commit length (tie-heavy queues confirm few)
    // a successor that precedes the next candidate is the next pop itself: it is recorded at once and carried into the
    // following round as candidate 0 (already popped; only its successors are still to come)
    bool carry = false;
    QE cy = INF;
    // Serial steps.  Queues full of exact distance ties pop in the order of the heap nodes, and a node's children were
    // allocated before it: the next pop is almost always a successor of the last one (a descent through equal keys), so a
    // 32-wide speculative round confirms one or two candidates.  In that regime one pop is handled at a time: lanes 0..2
    // work out the three successors, the smallest one that precedes the front is popped at once (and carried, as in
    // the rounds), the others enter the queue.  Same order, same entry indices.
    bool serial = false;
    int32_t pops16 = 16 * 16, ser_steps = 0;
    // (a lone warp per SM is latency-bound: a round of ~5 k cycles loses against ~2.2 k per serial step below ~2.3 pops per
    // round; with several contigs per SM the instruction count decides and the break-even is lower)
    const int32_t ser_thr16 = (w.heaps_variant >> 8) ? (w.heaps_variant >> 8) : (w.C > 2 * 148 ? 22 : 36);
    XRec xn;
    int32_t p12n = -1;
    bool have_xn = false;
    auto record = [&](const QE &e) {
        if (lane == 0) {
            D4 cd;
            cd.sum = e.sum;
            cd.anom = (int32_t)(e.k1 >> (S + 1));
            cd.nz = e.nz;
            cd.tot = e.tot;
            cd.aux = 0;
            dist[nd] = cd;
            last[nd] = (int32_t)(uint32_t)e.k2;
        }
        nd++;
    };
    while (nd < K) {
        if (n0 == 0 && np == 0 && !carry) {
            if (nR == 0 && nF == 0) break;
            ET(7);
            refill();
            ET(0);
            continue;
        }
        if (serial) {
            ET(7);
            QE t;
            if (carry) {
                t = cy;
                carry = false;
            } else {  // pop the front minimum
                if (!p0_ok) {
                    p0u = qe_shfl(P, ShIdx{0});
                    p0_ok = true;
                }
                const QE p0 = p0u;
                bool from_p = np > 0;
                QE rh = INF;
                if (n0 > 0) {
                    rh = qe_ld(fslot(0));
                    if (np == 0 || qe_less(rh, p0, wide)) from_p = false;
                }
                t = from_p ? p0 : rh;
                if (from_p) {
                    const QE dn = qe_shfl(P, ShDown{1});
                    P = lane + 1 < np ? dn : INF;
                    np--;
                    p0_ok = false;
                } else {
                    head = (head + 1) & FMASK;
                    n0--;
                }
                record(t);
                if (nd >= K) break;
            }
            ET(1);
            // ---- its successors: next-root on lane 0, left child on lane 1, right child on lane 2 ----
            const int32_t idx = (int32_t)(uint32_t)t.k2;
            const int32_t anom = (int32_t)(t.k1 >> (S + 1));
            QE a = INF;
            bool valid = false;
            int32_t snode = -1;
            int32_t p12;
            if (xrec) {
                XRec x;
                if (have_xn) {  // the carried entry's record was requested when it was chosen
                    x = xn;
                    p12 = p12n;
                    have_xn = false;
                } else {
                    const int32_t node = keyed ? en[idx] : (int32_t)(t.k2 >> 32);
                    x = xrec_ld(xrec + node);
                    p12 = ep[idx];
                }
                snode = lane == 0 ? x.hv : (lane == 1 ? x.left : x.right);
                valid = lane < 3 && snode >= 0;
                if (valid) {
                    const int64_t dsum = lane == 0 ? x.hsum : (lane == 1 ? x.lsum : x.rsum);
                    const int32_t danom = lane == 0 ? x.hanom : (lane == 1 ? x.lanom : x.ranom);
                    const int32_t dnz = lane == 0 ? x.hnz : (lane == 1 ? x.lnz : x.rnz);
                    const int32_t dtot = lane == 0 ? x.htot : (lane == 1 ? x.ltot : x.rtot);
                    const int32_t okv = lane == 0 ? x.hkey : (lane == 1 ? x.lkey : x.rkey);
                    a.sum = t.sum + dsum;
                    a.nz = t.nz + dnz;
                    a.tot = t.tot + dtot;
                    a.k1 = make_k1(anom + danom, a.nz, a.tot);
                    a.k2 = (uint64_t)(uint32_t)okv << 32;
                }
            } else {
                const int32_t node = keyed ? en[idx] : (int32_t)(t.k2 >> 32);
                const HNode ch = hn_load(hn + node);
                const int32_t ceid = w.hn_eid[node];
                p12 = ep[idx];
                if (lane == 0) {
                    const ENext x = enext[ceid];
                    if (x.hv >= 0) {
                        a.sum = t.sum + x.sum;
                        a.nz = t.nz + x.nz;
                        a.tot = t.tot + x.tot;
                        a.k1 = make_k1(anom + x.anom, a.nz, a.tot);
                        a.k2 = (uint64_t)(uint32_t)(keyed ? x.hrank : x.hv) << 32;
                        snode = x.hv;
                        valid = true;
                    }
                } else if (lane < 3) {
                    const int32_t cid = lane == 1 ? ch.left : ch.right;
                    if (cid >= 0) {
                        const HNode xc = hn_load(hn + cid);
                        a.sum = t.sum + xc.sum - ch.sum;
                        a.nz = t.nz + xc.nz - ch.nz;
                        a.tot = t.tot + xc.tot - ch.tot;
                        a.k1 = make_k1(anom + xc.anom - ch.anom, a.nz, a.tot);
                        a.k2 = okey(cid);
                        snode = cid;
                        valid = true;
                    }
                }
            }
            ET(2);
            const uint32_t vmask = __ballot_sync(FULL, valid);
            if (valid) {
                const int32_t id = ne + __popc(vmask & lt);
                a.k2 |= (uint32_t)id;
                en[id] = snode;
                ep[id] = lane == 0 ? idx : p12;
            }
            ne += __popc(vmask);
            if (valid && hasB && !qe_less(a, B, wide)) valid = false;  // beyond the K walks: dropped
            // ---- the smallest successor that precedes the front is the next pop ----
            if (!p0_ok) {
                p0u = qe_shfl(P, ShIdx{0});
                p0_ok = true;
            }
            QE fm = p0u;
            if (n0 > 0) {
                const QE rh = qe_ld(fslot(0));
                if (np == 0 || qe_less(rh, fm, wide)) fm = rh;
            }
            const bool front_empty = n0 == 0 && np == 0;
            if (lane < 3) qe_st(&sm.stage[lane], a);
            uint32_t live = __ballot_sync(FULL, valid) & 7u;
            __syncwarp();
            const QE b0 = qe_ld(&sm.stage[0]), b1 = qe_ld(&sm.stage[1]), b2 = qe_ld(&sm.stage[2]);
            __syncwarp();
            int32_t best = -1;
            {
                QE bm = INF;
                if ((live & 1u)) {
                    bm = b0;
                    best = 0;
                }
                if ((live & 2u) && (best < 0 || qe_less(b1, bm, wide))) {
                    bm = b1;
                    best = 1;
                }
                if ((live & 4u) && (best < 0 || qe_less(b2, bm, wide))) {
                    bm = b2;
                    best = 2;
                }
                const bool ok = best >= 0 && (front_empty ? (!hasT || qe_less(bm, T, wide)) : qe_less(bm, fm, wide));
                if (ok) {
                    record(bm);
                    cy = bm;
                    carry = true;
                    live &= ~(1u << best);
                    if (xrec && nd < K) {  // its expansion record is requested now and arrives while the others are queued
                        const int32_t nnode = __shfl_sync(FULL, snode, best);
                        xn = xrec_ld(xrec + nnode);
                        p12n = best == 0 ? idx : p12;
                        have_xn = true;
                    }
                }
            }
            ET(3);
            if (nd >= K) break;
            // ---- the other successors enter the queue ----
#pragma unroll
            for (int32_t j = 0; j < 3; j++) {
                if (!(live & (1u << j))) continue;
                const QE &e = j == 0 ? b0 : (j == 1 ? b1 : b2);
                if (hasT && !qe_less(e, T, wide)) {
                    back_put(e);
                } else {
                    pend_insert(e);
                }
            }
            ET(5);
            if (++ser_steps >= 24) {  // one speculative round as a probe: has the queue left the tie regime?
                serial = false;
                have_xn = false;  // (a carried entry is expanded by the round itself)
                width = 8;
                ser_steps = 0;
            }
            continue;
        }
        ET(7);
        const int32_t nd_round = nd;
        p0_ok = false;  // (the round moves the pending lanes)
#ifdef AA_ENUM_TIMERS
        {   // plateau statistics: entries of the front that tie with its minimum on the distance
            QE fm = qe_shfl(P, ShIdx{0});
            if (n0 > 0) {
                const QE rh = qe_ld(fslot(0));
                if (np == 0 || qe_less(rh, fm, wide)) fm = rh;
            }
            int32_t cnt = __popc(__ballot_sync(FULL, lane < np && P.sum == fm.sum && P.k1 == fm.k1));
            for (int32_t i0 = 0; i0 < n0; i0 += 32) {
                const bool in = i0 + lane < n0;
                QE x = INF;
                if (in) x = qe_ld(fslot(i0 + lane));
                const int32_t c1 = __popc(__ballot_sync(FULL, in && x.sum == fm.sum && x.k1 == fm.k1));
                cnt += c1;
                if (c1 < 32) break;
            }
            int32_t bkt = 0;
            while ((1 << bkt) < cnt && bkt < 11) bkt++;
            plat[bkt]++;
            if (!carry) { front_sz += n0 + np; front_n++; }
        }
#endif
        // ---- candidates: the smallest entries of pending + run head, sorted over the lanes ----
        QE t = INF;
        int32_t from_pend = 0;
        {
            const int32_t nb0 = n0 < 32 ? n0 : 32;
            if (np == 0) {
                if (lane < nb0) t = qe_ld(fslot(lane));
            } else if (nb0 == 0) {
                t = P;
                from_pend = lane < np;
            } else {
                QE br = INF;  // run head, reversed over the lanes
                if (31 - lane < nb0) br = qe_ld(fslot(31 - lane));
                t = P;
                from_pend = 1;
                if (qe_less(br, t, wide)) {
                    t = br;
                    from_pend = 0;
                }
#pragma unroll
                for (int32_t j = 16; j > 0; j >>= 1) {  // sort the bitonic sequence of the 32 smallest
                    const QE o = qe_shfl(t, ShXor{j});
                    const int32_t of = __shfl_xor_sync(FULL, from_pend, j);
                    const bool lower = (lane & j) == 0;
                    const bool o_less = qe_less(o, t, wide);
                    if (lower == o_less) {
                        t = o;
                        from_pend = of;
                    }
                }
            }
        }
        const int32_t cr = carry ? 1 : 0;
        if (carry) {  // the carried entry takes lane 0, the queue candidates move up
            t = qe_shfl(t, ShUp{1});
            from_pend = __shfl_up_sync(FULL, from_pend, 1);
            if (lane == 0) {
                t = cy;
                from_pend = 0;
            }
        }
        int32_t ncand = n0 + np + cr < width ? n0 + np + cr : width;
        if (ncand > K - nd + cr) ncand = K - nd + cr;
        const bool have = lane < ncand;
        ET(1);
        // ---- their successors (speculative beyond the first candidate) ----
        QE a0 = INF, a1 = INF, a2 = INF;
        int32_t p0 = -1, p12 = -1;  // prev of the successors
        bool v0s = false, v1s = false, v2s = false;
        int32_t sn0 = -1, sn1 = -1, sn2 = -1;  // node ids of the successors
        if (have && xrec) {
            const int32_t idx = (int32_t)(uint32_t)t.k2;
            const int32_t node = keyed ? en[idx] : (int32_t)(t.k2 >> 32);
            const int32_t anom = (int32_t)(t.k1 >> (S + 1));
            const XRec x = xrec_ld(xrec + node);
            p12 = ep[idx];
            p0 = idx;
            if (x.hv >= 0) {
                a0.sum = t.sum + x.hsum;
                a0.nz = t.nz + x.hnz;
                a0.tot = t.tot + x.htot;
                a0.k1 = make_k1(anom + x.hanom, a0.nz, a0.tot);
                a0.k2 = (uint64_t)(uint32_t)x.hkey << 32;
                sn0 = x.hv;
                v0s = true;
            }
            if (x.left >= 0) {
                a1.sum = t.sum + x.lsum;
                a1.nz = t.nz + x.lnz;
                a1.tot = t.tot + x.ltot;
                a1.k1 = make_k1(anom + x.lanom, a1.nz, a1.tot);
                a1.k2 = (uint64_t)(uint32_t)x.lkey << 32;
                sn1 = x.left;
                v1s = true;
            }
            if (x.right >= 0) {
                a2.sum = t.sum + x.rsum;
                a2.nz = t.nz + x.rnz;
                a2.tot = t.tot + x.rtot;
                a2.k1 = make_k1(anom + x.ranom, a2.nz, a2.tot);
                a2.k2 = (uint64_t)(uint32_t)x.rkey << 32;
                sn2 = x.right;
                v2s = true;
            }
        } else if (have) {
            const int32_t idx = (int32_t)(uint32_t)t.k2;
            const int32_t node = keyed ? en[idx] : (int32_t)(t.k2 >> 32);
            const int32_t anom = (int32_t)(t.k1 >> (S + 1));
            const HNode ch = hn_load(hn + node);
            const int32_t ceid = w.hn_eid[node];
            p12 = ep[idx];
            p0 = idx;
            HNode xl, xr;
            if (ch.left >= 0) xl = hn_load(hn + ch.left);
            if (ch.right >= 0) xr = hn_load(hn + ch.right);
            const ENext x = enext[ceid];
            if (x.hv >= 0) {
                a0.sum = t.sum + x.sum;
                a0.nz = t.nz + x.nz;
                a0.tot = t.tot + x.tot;
                a0.k1 = make_k1(anom + x.anom, a0.nz, a0.tot);
                a0.k2 = (uint64_t)(uint32_t)(keyed ? x.hrank : x.hv) << 32;
                sn0 = x.hv;
                v0s = true;
            }
            if (ch.left >= 0) {
                a1.sum = t.sum + xl.sum - ch.sum;
                a1.nz = t.nz + xl.nz - ch.nz;
                a1.tot = t.tot + xl.tot - ch.tot;
                a1.k1 = make_k1(anom + xl.anom - ch.anom, a1.nz, a1.tot);
                a1.k2 = okey(ch.left);
                sn1 = ch.left;
                v1s = true;
            }
            if (ch.right >= 0) {
                a2.sum = t.sum + xr.sum - ch.sum;
                a2.nz = t.nz + xr.nz - ch.nz;
                a2.tot = t.tot + xr.tot - ch.tot;
                a2.k1 = make_k1(anom + xr.anom - ch.anom, a2.nz, a2.tot);
                a2.k2 = okey(ch.right);
                sn2 = ch.right;
                v2s = true;
            }
        }
        ET(2);
        // ---- how many candidates does the sequential order confirm? ----
        int32_t m = ncand;
        bool violated = false;
        QE star = INF;  // the successor that precedes candidate m (valid when violated)
        if (ncand > 1) {
            QE pm = a0;  // this lane's smallest successor under (distance, node); INF when it has none
            if (qe_dn_less(a1, pm, wide)) pm = a1;
            if (qe_dn_less(a2, pm, wide)) pm = a2;
            for (int32_t d = 1; d < ncand; d <<= 1) {  // inclusive prefix minimum over the lanes
                const QE o = qe_shfl(pm, ShUp{d});
                if (lane >= d && qe_dn_less(o, pm, wide)) pm = o;
            }
            const QE ex = qe_shfl(pm, ShUp{1});
            const bool viol = have && lane > 0 && ex.sum != I64_MAX && qe_dn_less(ex, t, wide);
            const uint32_t vm = __ballot_sync(FULL, viol);
            if (vm) {
                m = __ffs(vm) - 1;
                violated = true;
                star = qe_shfl(pm, ShIdx{m - 1});  // minimum over the successors of candidates 0..m-1
            }
        }
        width = m >= ncand ? (2 * width < 32 ? 2 * width : 32) : (2 * m + 2 < 32 ? 2 * m + 2 : 32);
        ET(3);
        // ---- commit candidates 0..m-1 ----
        const bool com = lane < m;
        if (com && !(carry && lane == 0)) {
            D4 cd;
            cd.sum = t.sum;
            cd.anom = (int32_t)(t.k1 >> (S + 1));
            cd.nz = t.nz;
            cd.tot = t.tot;
            cd.aux = 0;
            dist[nd + lane - cr] = cd;
            last[nd + lane - cr] = (int32_t)(uint32_t)t.k2;
        }
        const int32_t v0c = com && v0s, v1c = com && v1s, v2c = com && v2s;
        const int32_t mycnt = v0c + v1c + v2c;
        int32_t inc = mycnt;
        for (int32_t d = 1; d < 32; d <<= 1) {
            const int32_t o = __shfl_up_sync(FULL, inc, d);
            if (lane >= d) inc += o;
        }
        const int32_t total = __shfl_sync(FULL, inc, 31);
        int32_t id = ne + inc - mycnt;
        if (v0c) {
            a0.k2 |= (uint32_t)id;
            en[id] = sn0;
            ep[id] = p0;
            id++;
        }
        if (v1c) {
            a1.k2 |= (uint32_t)id;
            en[id] = sn1;
            ep[id] = p12;
            id++;
        }
        if (v2c) {
            a2.k2 |= (uint32_t)id;
            en[id] = sn2;
            ep[id] = p12;
        }
        ne += total;
        nd += m - cr;
        // ---- the preceding successor, if any, is the next pop: record it, keep it out of the queue, carry it over ----
        bool s0 = v0c, s1 = v1c, s2 = v2c;  // successors that enter the queue
        carry = false;
        if (violated && nd < K && !wide) {  // (packed keys: equal (sum, k1, node) is equality in (distance, node))
            const uint64_t sn = star.k2 >> 32;
            const bool e0 = v0c && a0.sum == star.sum && a0.k1 == star.k1 && (a0.k2 >> 32) == sn;
            const bool e1 = v1c && a1.sum == star.sum && a1.k1 == star.k1 && (a1.k2 >> 32) == sn;
            const bool e2 = v2c && a2.sum == star.sum && a2.k1 == star.k1 && (a2.k2 >> 32) == sn;
            const uint32_t em = __ballot_sync(FULL, e0 || e1 || e2);  // equal (distance, node): the smallest entry index wins
            const int32_t L = __ffs(em) - 1;
            const QE mine = e0 ? a0 : (e1 ? a1 : a2);
            cy = qe_shfl(mine, ShIdx{L});
            if (lane == L) {
                if (e0) s0 = false;
                else if (e1) s1 = false;
                else s2 = false;
            }
            if (lane == 0) {
                D4 cd;
                cd.sum = cy.sum;
                cd.anom = (int32_t)(cy.k1 >> (S + 1));
                cd.nz = cy.nz;
                cd.tot = cy.tot;
                cd.aux = 0;
                dist[nd] = cd;
                last[nd] = (int32_t)(uint32_t)cy.k2;
            }
            nd++;
            carry = true;
        }
        // pops this round (running mean, 1/16 units): below ~2 a round costs more than serial steps
        pops16 = (pops16 + 16 * (nd - nd_round)) >> 1;
        if (!wide && pops16 < ser_thr16 && !(w.heaps_variant & 1)) serial = true;
        {  // the committed entries leave the front: a prefix of the pending lanes and a prefix of the run
            const int32_t na = __popc(__ballot_sync(FULL, com && from_pend));
            if (na > 0) {
                const QE dn = qe_shfl(P, ShDown{na});
                P = lane + na < np ? dn : INF;
                np -= na;
            }
            head = (head + (m - cr - na)) & FMASK;
            n0 -= m - cr - na;
        }
        if (nd >= K) break;
        ET(4);
        // ---- the successors enter the queue: backlog appends in parallel, pending inserts one by one ----
#pragma unroll
        for (int32_t sidx = 0; sidx < 3; sidx++) {
            const QE &a = sidx == 0 ? a0 : (sidx == 1 ? a1 : a2);
            bool valid = sidx == 0 ? s0 : (sidx == 1 ? s1 : s2);
            if (valid && hasB && !qe_less(a, B, wide)) valid = false;  // beyond the K walks: dropped
            const bool toF = valid && (!hasT || qe_less(a, T, wide));
            bool toR = valid && !toF;
            if (hasT2) {  // keys at or above T2 go to the far region
                const bool toFar = toR && !qe_less(a, T2, wide);
                const uint32_t mX = __ballot_sync(FULL, toFar);
                if (toFar) qe_st(far + nF + __popc(mX & lt), a);
                nF += __popc(mX);
                toR = toR && !toFar;
            }
            const uint32_t mR = __ballot_sync(FULL, toR);
            if (toR) qe_st(back + nR + __popc(mR & lt), a);
            nR += __popc(mR);
            uint32_t mF = __ballot_sync(FULL, toF);
            while (mF) {
                const int32_t l = __ffs(mF) - 1;
                mF &= mF - 1;
                const QE e = qe_shfl(a, ShIdx{l});
                if (hasT && !qe_less(e, T, wide)) {  // the threshold dropped since toF was evaluated (a spill)
                    if (lane == 0) qe_st(back + nR, e);
                    nR++;
                } else {
                    pend_insert(e);
                }
            }
        }
        ET(5);
    }
#ifdef AA_ENUM_TIMERS
    if (lane == 0 && (g.V > 4000 || c % 250 == 7))
        printf("enum ctg %ld V %d walks %d entries %d | refill %lld/%lld cand %lld/%lld expand %lld/%lld scan %lld/%lld commit %lld/%lld succ %lld/%lld other %lld/%lld\n",
               (long)c, g.V, nd, ne, et_acc[0], et_n[0], et_acc[1], et_n[1], et_acc[2], et_n[2], et_acc[3], et_n[3], et_acc[4], et_n[4], et_acc[5], et_n[5],
               et_acc[7], et_n[7]);
    if (lane == 0 && (g.V > 4000 || c % 250 == 7))
        printf("  refill parts: sample sort %lld pass %lld run sort %lld | sum ns %lld sum scanned %lld sum n0 %lld\n", rt_acc[0], rt_acc[1], rt_acc[2], rt_acc[3], rt_acc[4], rt_acc[5]);
    if (lane == 0 && (g.V > 4000 || c % 250 == 7))
        printf("  plateau size (log2 buckets 1,2,4,..): %d %d %d %d %d %d %d %d %d %d %d %d | avg front %lld backlog %d\n", plat[0], plat[1], plat[2], plat[3], plat[4],
               plat[5], plat[6], plat[7], plat[8], plat[9], plat[10], plat[11], front_n ? front_sz / front_n : 0, nR);
#endif
    if (lane == 0) w.n_walk[c] = nd;
}
__device__ void f_enum_warp(const Ws &w, int64_t c, void *scratch) {
    bool wide = false;
    if (w.status[c] == 0) wide = ctg_view(w, c).V >= (1 << 20);  // the packed key holds vertex numbers below 2^20 twice
    if (w.heaps_variant & 8) wide = true;  // AA_TUNE bit 3 (tests): the wide comparisons are exact for every contig
    if (wide) f_enum_warp_t<true>(w, c, scratch);
    else f_enum_warp_t<false>(w, c, scratch);
}
#endif
AA_HDN void f_enum_any(const Ws &w, int64_t c, void *scratch) {
#if defined(__CUDA_ARCH__)
    f_enum_warp(w, c, scratch);
#else
    (void)scratch;
    f_enum(w, c);
#endif
}
constexpr size_t ENUM_SMEM_BYTES = (size_t)AA_FCAP * 32 + 128;  // >= sizeof(EnumSmem) (device only)

// phase: plan the edge_path_to_paf_path calls of a contig in the reference's order (paf_data.cpp:1585-1649):
// walk 0, the walks tied with it on (score_sum, anom), then the alt candidates.
AA_HD bool same_sa(const D4 &a, const D4 &b) { return a.sum == b.sum && a.anom == b.anom; }
AA_HDN void f_plan(const Ws &w, int64_t c) {
    if (aa_lane() != 0) return;
    w.n_task[c] = 0;
    w.n_tie[c] = 0;
    w.last_group[c] = 0;
    if (w.status[c] != 0) return;
    const int64_t wo = w.walk_off[c];
    const D4 *dist = w.wdist + wo;
    Task *t = w.task + 2 * wo;
    const int32_t nw = w.n_walk[c];
    int32_t nt = 0;
    const D4 mind = dist[0];
    t[nt] = Task{(int32_t)c, 0, nt, 0};
    nt++;
    for (int32_t i = 1; i < nw && same_sa(mind, dist[i]); i++) {
        t[nt] = Task{(int32_t)c, i, nt, 0};
        nt++;
    }
    w.n_tie[c] = nt;
    int32_t group = 0;
    if (nw >= 2 && (int64_t)mind.anom != w.anom_dis[c]) {
        int64_t ans_up = 0, ans_down = 0;
        int32_t ans = -1;
        for (int32_t i = 1; i < nw; i++) {
            D4 x = dist[i];
            if (x.anom >= mind.anom) continue;
            int64_t up = x.sum - mind.sum;
            int64_t down = (int64_t)mind.anom - x.anom;
            if (ans == -1 || up * ans_down < down * ans_up) {
                ans_up = up;
                ans_down = down;
                ans = i;
                group++;
                t[nt] = Task{(int32_t)c, i, nt, group};
                nt++;
            } else if (same_sa(x, dist[ans])) {
                t[nt] = Task{(int32_t)c, i, nt, group};
                nt++;
            }
        }
    }
    w.n_task[c] = nt;
    w.last_group[c] = group;
}
#if defined(__CUDA_ARCH__)
// ---- warp-parallel plan (paf_data.cpp:1585-1649): lanes scan the distance list, the rare candidates
// (anom below the minimum walk's) are handled in order
__device__ void f_plan_warp(const Ws &w, int64_t c) {
    const uint32_t FULL = 0xffffffffu;
    const int32_t lane = (int32_t)(threadIdx.x & 31);
    if (lane == 0) {
        w.n_task[c] = 0;
        w.n_tie[c] = 0;
        w.last_group[c] = 0;
    }
    if (w.status[c] != 0) return;
    const int64_t wo = w.walk_off[c];
    const D4 *__restrict__ dist = w.wdist + wo;
    Task *__restrict__ t = w.task + 2 * wo;
    const int32_t nw = w.n_walk[c];
    const D4 mind = dist[0];
    // walks tied with walk 0: the longest prefix with equal (score_sum, anom)
    int32_t ntie = nw;
    for (int32_t base = 0; base < nw; base += 32) {
        const int32_t i = base + lane;
        bool diff = false;
        if (i < nw) {
            const D4 x = dist[i];
            diff = !same_sa(mind, x);
            if (!diff) t[i] = Task{(int32_t)c, i, i, 0};  // harmless beyond the first difference: overwritten below
        }
        const uint32_t m = __ballot_sync(FULL, diff);
        if (m) {
            ntie = base + __ffs(m) - 1;
            break;
        }
    }
    __syncwarp();
    int32_t nt = ntie, group = 0;
    if (nw >= 2 && (int64_t)mind.anom != w.anom_dis[c]) {
        int64_t ans_up = 0, ans_down = 0;
        int32_t ans = -1;
        D4 ansd = mind;
        for (int32_t base = 0; base < nw; base += 32) {
            const int32_t i = base + lane;
            D4 x = mind;
            bool cand = false;
            if (i >= 1 && i < nw) {
                x = dist[i];
                cand = x.anom < mind.anom;
            }
            uint32_t m = __ballot_sync(FULL, cand);
            while (m) {
                const int32_t l = __ffs(m) - 1;
                m &= m - 1;
                D4 y;
                y.sum = __shfl_sync(FULL, x.sum, l);
                y.anom = __shfl_sync(FULL, x.anom, l);
                y.nz = 0;
                y.tot = 0;
                y.aux = 0;
                const int32_t idx = base + l;
                const int64_t up = y.sum - mind.sum;
                const int64_t down = (int64_t)mind.anom - y.anom;
                if (ans == -1 || up * ans_down < down * ans_up) {
                    ans_up = up;
                    ans_down = down;
                    ans = idx;
                    ansd = y;
                    group++;
                    if (lane == 0) t[nt] = Task{(int32_t)c, idx, nt, group};
                    nt++;
                } else if (same_sa(y, ansd)) {
                    if (lane == 0) t[nt] = Task{(int32_t)c, idx, nt, group};
                    nt++;
                }
            }
        }
    }
    __syncwarp();
    if (lane == 0) {
        w.n_tie[c] = ntie;
        w.n_task[c] = nt;
        w.last_group[c] = group;
    }
}
#endif
AA_HDN void f_parts_any(const Ws &w, int64_t c) {  // (the device path runs three segmented scans instead: f_parts_bnd / f_parts_fin)
    f_parts(w, c);
}
AA_HDN void f_plan_any(const Ws &w, int64_t c) {
#if defined(__CUDA_ARCH__)
    f_plan_warp(w, c);
#else
    f_plan(w, c);
#endif
}

AA_HDN void f_task_compact(const Ws &w, int64_t c) {
    int64_t o = w.task_off[c];
    const Task *t = w.task + 2 * w.walk_off[c];
    const int32_t n = w.n_task[c];
    for (int32_t k = aa_lane(); k < n; k += AA_LANES) w.tasks[o + k] = t[k];
}

// ---- walk task: recover + mark + upgrade + rows ---------------------------------------------------
struct DPBuf {  // scratch of one gap-filling DP
    D5 *dp;
    int32_t *pre;
    uint8_t *seen;
    int32_t *up;   // vertices appended by the last sub_path call
    int32_t cap;   // largest topological span the buffers hold
};
struct Slot {
    int32_t *walk, *side;
    DPBuf buf;
};
AA_HD Slot slot_view(const Ws &w, int64_t s) {
    Slot x;
    int64_t o = s * w.slot_stride;
    x.walk = w.sc_walk + o;
    x.side = w.sc_side + o;
    x.buf.up = w.sc_up + o;
    x.buf.dp = w.sc_dp + o;
    x.buf.pre = w.sc_pre + o;
    x.buf.seen = w.sc_seen + o;
    x.buf.cap = (int32_t)(w.slot_stride < 0x7fffffff ? w.slot_stride : 0x7fffffff);
    return x;
}
constexpr int32_t LOCAL_SPAN = 32;
struct LocalDP {  // per-thread scratch of the parallel (speculative / row emitting) passes
    D5 dp[LOCAL_SPAN];
    int32_t pre[LOCAL_SPAN];
    int32_t up[LOCAL_SPAN];
    uint8_t seen[LOCAL_SPAN];
    AA_HD DPBuf buf() {
        DPBuf b;
        b.dp = dp;
        b.pre = pre;
        b.seen = seen;
        b.up = up;
        b.cap = LOCAL_SPAN;
        return b;
    }
};
AA_HD void vtx_xy(const Ws &w, const Ctg &g, int32_t v, int32_t &x, int32_t &y) {  // index_to_vtx
    if (v < g.n) {
        x = y = v;
    } else {
        const CandRec &p = w.pair[g.p0 + (v - g.n)];
        x = p.i;
        y = p.j;
    }
}

// kth_shortest_walk_recover (k_shortest_walks.hpp:254-290) as a vertex sequence src..dest; returns length
AA_HDN int32_t recover_walk(const Ws &w, const Ctg &g, int64_t c, int32_t k, const Slot &s) {
    const int64_t v0 = g.v0, e0 = w.eoff[v0], wo = w.walk_off[c];
    const int32_t *en = w.ent_node + 3 * wo;
    const int32_t *ep = w.ent_prev + 3 * wo;
    int32_t ns = 0;
    for (int32_t cur = w.wlast[wo + k]; cur != -1; cur = ep[cur]) s.side[ns++] = w.hn_eid[en[cur]];
    // side[] holds the sidetrack edges last-to-first
    int32_t len = 0;
    int32_t cur = g.src;
    s.walk[len++] = cur;
    int32_t idx = ns - 1;
    while (cur != g.dest || idx >= 0) {
        if (idx >= 0 && cur == w.e_src[e0 + s.side[idx]]) {
            cur = e_dst(w.edge[e0 + s.side[idx]]);
            idx--;
        } else {
            cur = w.best[v0 + cur];
        }
        s.walk[len++] = cur;
    }
    return len;
}

// internal_shortest_path_recover (paf_data.cpp:750-792): QRY_SCORE-mode DP over the forward topological
// range [order[s], order[t]).  Appends the vertices after s (optionally without t itself) to buf.up[] from
// up_len on.  Returns 0 when s == t (the reference's empty path), 1 on success, -1 when the range does not
// fit the scratch (only possible with the small per-thread buffers).
AA_HDN int sub_path(const Ws &w, const Ctg &g, const DPBuf &s, int32_t vs, int32_t vt, bool wl_flag, int32_t wl,
                    bool drop_last, int32_t &up_len) {
    if (vs == vt) return 0;
    const int64_t v0 = g.v0;
    const int32_t os = w.order[v0 + vs], ot = w.order[v0 + vt];
    const int32_t span = ot - os + 1;
    if (span > s.cap) return -1;
    for (int32_t k = 0; k < span; k++) s.seen[k] = 0;
    D5 z;
    z.qry = z.ref = 0;
    z.anom = z.nz = z.tot = z.pad = 0;
    s.dp[0] = z;
    s.pre[0] = -1;
    s.seen[0] = 1;
    for (int32_t i = os; i < ot; i++) {
        if (!s.seen[i - os]) continue;
        const int32_t u = w.topo[v0 + i];
        const D5 cur = s.dp[i - os];
        int32_t ux = -1, uy = -1;
        if (wl_flag && u != g.src && u != g.dest) vtx_xy(w, g, u, ux, uy);
        const int64_t ea = w.eoff[v0 + u], eb = w.eoff[v0 + u + 1];
        for (int64_t k = ea; k < eb; k++) {
            const Edge e = w.edge[k];
            const int32_t v = e_dst(e);
            if (wl_flag && v == vt) {
                if (u == g.src || u == g.dest) continue;
                if (uy != wl) continue;
            }
            const int32_t ov = w.order[v0 + v];
            if (ov > ot) continue;  // outside the range: never read again by the reference's loop
            D5 nx;
            nx.qry = cur.qry + e.qry;
            nx.ref = cur.ref + e.ref;
            nx.anom = cur.anom + e_anom(e);
            nx.nz = cur.nz + e_nz(e);
            nx.tot = cur.tot + e_tot(e);
            nx.pad = 0;
            const int32_t at = ov - os;
            if (!s.seen[at] || less5(nx, s.dp[at])) {
                s.dp[at] = nx;
                s.pre[at] = u;
                s.seen[at] = 1;
            }
        }
    }
    // walk back t -> s, then append in forward order
    int32_t cnt = 0;
    for (int32_t last = vt; last != vs; last = s.pre[w.order[v0 + last] - os]) cnt++;
    int32_t keep = drop_last ? cnt - 1 : cnt;  // vertices after s that are appended
    int32_t pos = up_len + cnt - 1;
    for (int32_t last = vt; last != vs; last = s.pre[w.order[v0 + last] - os]) {
        if (pos < up_len + keep) s.up[pos] = last;
        pos--;
    }
    up_len += keep;
    return 1;
}

AA_HD void mark_block(const Ws &w, int64_t gb, int32_t call) {
#if defined(__CUDA_ARCH__)
    atomicMin(&w.first_call[gb], call);
#else
    if (call < w.first_call[gb]) w.first_call[gb] = call;
#endif
}

// ---- the upgrade automaton, one step at a time -----------------------------------------------------------
// upgrade_edge_path_with_alt_path (paf_data.cpp:795-921) + the row construction of edge_path_to_paf_path
// (paf_data.cpp:1502-1557) as an automaton over the walk's vertex sequence.
// State = (position on the walk, cs = last vertex appended to the upgraded path).  A step looks at the next
// two walk vertices (v, nv), appends vertices and consumes one or two walk edges.  A row is closed, i.e.
// final, when the vertex after it is appended, because arriving at a pair vertex trims the end of the row
// before it (paf_data.cpp:1523-1531, 1546-1553).
struct Auto {
    int32_t cs;
    int64_t cov;
    int32_t rows;
};
struct Emit {      // where closed rows go
    int32_t mode;  // 0 nowhere, 1 main-chain rows (mr_*), 2 result rows (r_*[dst])
    int32_t dst;
    int64_t base;  // row index = base + number of rows closed before
    int32_t call;  // for the tp flag (paf_data.cpp:1560-1566)
};
AA_HD void auto_append(const Ws &w, const Ctg &g, Auto &A, int32_t b, const Emit &em) {
    if (A.cs != g.src) {
        int64_t qs_, rs_, qe_, re_, gb;
        if (A.cs < g.n) {
            gb = g.b0 + A.cs;
            qs_ = w.qs[gb];
            rs_ = w.rs[gb];
        } else {
            const CandRec &p = w.pair[g.p0 + (A.cs - g.n)];
            gb = g.b0 + p.j;
            qs_ = p.st_q;
            rs_ = p.st_r;
        }
        if (b >= g.n && b < g.src) {
            const CandRec &pb = w.pair[g.p0 + (b - g.n)];
            qe_ = pb.pe_q;
            re_ = pb.pe_r;
        } else {
            qe_ = w.qe[gb];
            re_ = w.re[gb];
        }
        int64_t dr = re_ - rs_;
        A.cov += (qe_ - qs_) + (dr < 0 ? -dr : dr);
        if (em.mode == 1) {
            const int64_t o = em.base + A.rows;
            w.mr_blk[o] = (int32_t)(gb - g.b0);
            w.mr_qs[o] = qs_;
            w.mr_qe[o] = qe_;
            w.mr_rs[o] = rs_;
            w.mr_re[o] = re_;
        } else if (em.mode == 2) {
            const int64_t o = em.base + A.rows;
            w.r_idx[em.dst][o] = w.orig[gb];
            w.r_qs[em.dst][o] = qs_;
            w.r_qe[em.dst][o] = qe_;
            w.r_rs[em.dst][o] = rs_;
            w.r_re[em.dst][o] = re_;
            w.r_alt[em.dst][o] = w.first_call[gb] > em.call ? 1 : 0;
        }
        A.rows++;
    }
    A.cs = b;
}
// one automaton step; returns the number of walk edges consumed (1 or 2), or -1 when the DP range does not
// fit the scratch (nothing has been appended then)
AA_HDN int32_t auto_step(const Ws &w, const Ctg &g, const DPBuf &s, Auto &A, int32_t v, int32_t nv, const Emit &em) {
    int32_t ul = 0;
    if (v == g.dest) {  // paf_data.cpp:845-858
        if (sub_path(w, g, s, A.cs, v, false, -1, false, ul) < 0) return -1;
        for (int32_t k = 0; k < ul; k++) auto_append(w, g, A, s.up[k], em);
        return 1;
    }
    int32_t x, y;
    vtx_xy(w, g, v, x, y);
    if (x != y) {  // paf_data.cpp:866-873 (after src the first vertex is always a single)
        auto_append(w, g, A, v, em);
        return 1;
    }
    int32_t nx = -1, ny = -1;
    if (nv != g.dest) vtx_xy(w, g, nv, nx, ny);
    if (nv == g.dest || nx == ny) {  // paf_data.cpp:812-833, 879-899
        const int r = sub_path(w, g, s, A.cs, nv, true, y, true, ul);
        if (r < 0) return -1;
        if (r == 0) {
            auto_append(w, g, A, v, em);
        } else {
            for (int32_t k = 0; k < ul; k++) auto_append(w, g, A, s.up[k], em);
        }
        return 1;
    }
    // nv = (y, ny): paf_data.cpp:834-843, 900-909
    const int r = sub_path(w, g, s, A.cs, nv, false, -1, false, ul);
    if (r < 0) return -1;
    if (r == 0) {
        auto_append(w, g, A, v, em);
        auto_append(w, g, A, nv, em);
    } else {
        for (int32_t k = 0; k < ul; k++) auto_append(w, g, A, s.up[k], em);
    }
    return 2;
}

// ---- main chain = the automaton's states along walk 0 (no sidetracks: the tree walk from src) -----------------
// The main-chain kernels run on the side stream beside the heaps and the enumeration.  Status 3 (heap arena overflow) is
// raised and cleared by the heaps while they run: it says nothing about walk 0.
AA_HD bool main_skip(const Ws &w, int64_t c) { return w.status[c] != 0 && w.status[c] != 3; }
// S0 (per contig): trace walk 0, mark its blocks with call 0
AA_HDN void f_main_trace(const Ws &w, int64_t c) {
    if (aa_lane() != 0) return;
    w.m_len[c] = 0;
    if (main_skip(w, c)) return;
    Ctg g = ctg_view(w, c);
    const int64_t v0 = g.v0;
    for (int32_t v = 0; v < g.V; v++) {
        w.main_pos[v0 + v] = -1;
        w.m_cs[v0 + v] = -1;
        w.m_done[v0 + v] = 0;
        w.sp_used[v0 + v] = 0;
    }
    int32_t m = 0;
    for (int32_t cur = g.src;; cur = w.best[v0 + cur]) {
        w.main_walk[v0 + m] = cur;
        w.main_pos[v0 + cur] = m;
        if (cur == g.dest) break;
        if (cur != g.src) {
            int32_t x, y;
            vtx_xy(w, g, cur, x, y);
            mark_block(w, g.b0 + x, 0);
            mark_block(w, g.b0 + y, 0);
        }
        m++;
    }
    w.m_len[c] = m;
}
// S1 (per walk-0 position, parallel): the step at position i under the assumption cs == walk vertex
AA_HDN void f_main_spec(const Ws &w, int64_t gv) {
    const int64_t c = upper_idx(w.vtx_off, w.C, gv);
    if (main_skip(w, c)) return;
    const int64_t v0 = w.vtx_off[c];
    const int32_t i = (int32_t)(gv - v0), m = w.m_len[c];
    if (i >= m) return;
    Ctg g = ctg_view(w, c);
    LocalDP loc;
    Auto A;
    A.cs = w.main_walk[v0 + i];
    A.cov = 0;
    A.rows = 0;
    Emit em;
    em.mode = 0;
    em.dst = 0;
    em.base = 0;
    em.call = 0;
    const int32_t v = w.main_walk[v0 + i + 1];
    const int32_t nv = (i + 2 <= m) ? w.main_walk[v0 + i + 2] : -1;
    const int32_t used = auto_step(w, g, loc.buf(), A, v, nv, em);
    if (used < 0) return;  // sp_used stays 0
    w.sp_cs[gv] = A.cs;
    w.sp_cov[gv] = A.cov;
    w.sp_rows[gv] = A.rows;
    w.sp_used[gv] = (uint8_t)used;
}
// S2 (per contig, sequential but cheap): resolve the true cs chain; take the speculated step whenever its
// assumption holds, otherwise run the step with the real cs (and emit its rows right away)
AA_HDN void f_main_resolve(const Ws &w, int64_t c, const Slot &s) {
    if (main_skip(w, c)) return;
    Ctg g = ctg_view(w, c);
    const int64_t v0 = g.v0;
    const int32_t m = w.m_len[c];
    Auto A;
    A.cs = g.src;
    A.cov = 0;
    A.rows = 0;
    Emit em;
    em.mode = 1;
    em.dst = 0;
    em.base = v0;
    em.call = 0;
    int32_t i = 0;
    while (i < m) {
        w.m_cs[v0 + i] = A.cs;
        w.m_cov[v0 + i] = A.cov;
        w.m_rows[v0 + i] = A.rows;
        int32_t used = w.sp_used[v0 + i];
        if (used != 0 && A.cs == w.main_walk[v0 + i]) {
            A.cov += w.sp_cov[v0 + i];
            A.rows += w.sp_rows[v0 + i];
            A.cs = w.sp_cs[v0 + i];
        } else {
            const int32_t v = w.main_walk[v0 + i + 1];
            const int32_t nv = (i + 2 <= m) ? w.main_walk[v0 + i + 2] : -1;
            used = auto_step(w, g, s.buf, A, v, nv, em);
            w.m_done[v0 + i] = 1;
        }
        i += used;
    }
    w.m_tot_cov[c] = A.cov;
    w.m_tot_rows[c] = A.rows;
}
// task 0 of a contig is its walk 0: coverage and row count come from the main chain (once the task table exists)
AA_HDN void f_main_totals(const Ws &w, int64_t c) {
    if (w.status[c] != 0) return;
    const int64_t t0 = w.task_off[c];
    w.task_cov[t0] = w.m_tot_cov[c];
    w.task_rows[t0] = w.m_tot_rows[c];
}
#if defined(__CUDA_ARCH__)
// S2 on the device: the same chain, but the per-position records are loaded 32 at a time (coalesced, one chunk
// ahead) and the chain walks through them by shuffles; only a failed speculation runs the real step (lane 0)
__device__ void f_main_resolve_warp(const Ws &w, int64_t c, const Slot &s) {
    const uint32_t FULL = 0xffffffffu;
    const int32_t lane = (int32_t)(threadIdx.x & 31);
    if (main_skip(w, c)) return;
    Ctg g = ctg_view(w, c);
    const int64_t v0 = g.v0;
    const int32_t m = w.m_len[c];
    Auto A;
    A.cs = g.src;
    A.cov = 0;
    A.rows = 0;
    Emit em;
    em.mode = 1;
    em.dst = 0;
    em.base = v0;
    em.call = 0;
    struct Rec {
        int32_t used, cs, rows, mw;
        int64_t cov;
    };
    auto load = [&](int32_t b) -> Rec {
        Rec r;
        r.used = 0;
        r.cs = r.rows = 0;
        r.mw = -1;
        r.cov = 0;
        const int32_t i = b + lane;
        if (i < m) {
            r.used = w.sp_used[v0 + i];
            r.cs = w.sp_cs[v0 + i];
            r.rows = w.sp_rows[v0 + i];
            r.cov = w.sp_cov[v0 + i];
        }
        if (i <= m) r.mw = w.main_walk[v0 + i];
        return r;
    };
    int32_t b = 0;
    Rec cur = load(0), nxt = load(32);
    // state of the automaton when it stood at this lane's position (stored when the chunk is left)
    int32_t o_cs = -1, o_rows = 0;
    int64_t o_cov = 0;
    bool o_set = false, o_done = false;
    auto store_chunk = [&]() {
        const int32_t i = b + lane;
        if (o_set && i < m) {
            w.m_cs[v0 + i] = o_cs;
            w.m_cov[v0 + i] = o_cov;
            w.m_rows[v0 + i] = o_rows;
            if (o_done) w.m_done[v0 + i] = 1;
        }
        o_set = o_done = false;
    };
    int32_t i = 0;
    while (i < m) {
        while (i >= b + 32) {
            store_chunk();
            b += 32;
            cur = nxt;
            nxt = load(b + 32);
        }
        const int32_t e = i - b;
        // which positions of this chunk can take their speculated step, and where that step leads
        const int32_t tgt = lane + cur.used;
        const int32_t mw_a = __shfl_sync(FULL, cur.mw, tgt & 31), mw_b = __shfl_sync(FULL, nxt.mw, tgt & 31);
        const int32_t mwt = tgt < 32 ? mw_a : mw_b;                           // walk vertex at the position after the step
        const uint32_t U0 = __ballot_sync(FULL, cur.used != 0);               // speculated
        const uint32_t U2 = __ballot_sync(FULL, cur.used == 2);
        const uint32_t NX = __ballot_sync(FULL, cur.used != 0 && cur.cs == mwt);  // ... and the next position's assumption holds
        const int32_t mw_e = __shfl_sync(FULL, cur.mw, e);
        if ((U0 >> e & 1u) && A.cs == mw_e) {
            // follow the chain of speculated steps through the chunk (warp-uniform bit arithmetic, no memory)
            uint32_t vis = 0;
            int32_t l = e, last;
            for (;;) {
                vis |= 1u << l;
                last = l;
                const int32_t nl = l + 1 + (int32_t)(U2 >> l & 1u);
                if (!(NX >> l & 1u) || nl >= 32 || !(U0 >> nl & 1u)) break;
                l = nl;
            }
            const bool mine = (vis >> lane & 1u) != 0;
            int64_t icov = mine ? cur.cov : 0;
            int32_t irows = mine ? cur.rows : 0;
            for (int32_t d = 1; d < 32; d <<= 1) {  // inclusive prefix sums over the visited positions
                const int64_t oc = __shfl_up_sync(FULL, icov, d);
                const int32_t orr = __shfl_up_sync(FULL, irows, d);
                if (lane >= d) {
                    icov += oc;
                    irows += orr;
                }
            }
            if (mine) {
                o_cs = cur.mw;
                o_cov = A.cov + icov - cur.cov;
                o_rows = A.rows + irows - cur.rows;
                o_set = true;
            }
            A.cov += __shfl_sync(FULL, icov, 31);
            A.rows += __shfl_sync(FULL, irows, 31);
            A.cs = __shfl_sync(FULL, cur.cs, last);
            i = b + last + 1 + (int32_t)(U2 >> last & 1u);
        } else {
            // the assumption fails here (or nothing was speculated): the real step, by lane 0
            if (lane == e) {
                o_cs = A.cs;
                o_cov = A.cov;
                o_rows = A.rows;
                o_set = true;
                o_done = true;
            }
            int32_t used = 0;
            if (lane == 0) {
                const int32_t v = w.main_walk[v0 + i + 1];
                const int32_t nv = (i + 2 <= m) ? w.main_walk[v0 + i + 2] : -1;
                used = auto_step(w, g, s.buf, A, v, nv, em);
            }
            used = __shfl_sync(FULL, used, 0);
            A.cs = __shfl_sync(FULL, A.cs, 0);
            A.cov = __shfl_sync(FULL, A.cov, 0);
            A.rows = __shfl_sync(FULL, A.rows, 0);
            i += used;
        }
    }
    store_chunk();
    if (lane == 0) {
        w.m_tot_cov[c] = A.cov;
        w.m_tot_rows[c] = A.rows;
    }
}
#endif
// S3 (per walk-0 position, parallel): emit the rows of every state the resolve pass took from speculation
AA_HDN void f_main_rows(const Ws &w, int64_t gv) {
    const int64_t c = upper_idx(w.vtx_off, w.C, gv);
    if (main_skip(w, c)) return;
    const int64_t v0 = w.vtx_off[c];
    const int32_t i = (int32_t)(gv - v0), m = w.m_len[c];
    if (i >= m || w.m_cs[gv] < 0 || w.m_done[gv]) return;
    Ctg g = ctg_view(w, c);
    LocalDP loc;
    Auto A;
    A.cs = w.m_cs[gv];
    A.cov = 0;
    A.rows = w.m_rows[gv];
    Emit em;
    em.mode = 1;
    em.dst = 0;
    em.base = v0;
    em.call = 0;
    const int32_t v = w.main_walk[v0 + i + 1];
    const int32_t nv = (i + 2 <= m) ? w.main_walk[v0 + i + 2] : -1;
    auto_step(w, g, loc.buf(), A, v, nv, em);
}

// copy rows [r0, r1) of the main chain into the task's output (lanes stride the range)
AA_HD void copy_main_rows(const Ws &w, const Ctg &g, int32_t r0, int32_t r1, const Emit &em, int32_t rows_before) {
    if (em.mode != 2) return;
    // four rows per lane and step: the dependent gathers (row -> block -> original index / first call) of the four
    // overlap instead of queueing up behind one another
    for (int32_t rb = r0 + aa_lane(); rb < r1; rb += 4 * AA_LANES) {
        int64_t gb[4], qs_[4], qe_[4], rs_[4], re_[4];
        int32_t oi[4], fc[4];
#pragma unroll
        for (int32_t u = 0; u < 4; u++) {
            const int32_t r = rb + u * AA_LANES;
            if (r < r1) {
                const int64_t src = g.v0 + r;
                gb[u] = g.b0 + w.mr_blk[src];
                qs_[u] = w.mr_qs[src];
                qe_[u] = w.mr_qe[src];
                rs_[u] = w.mr_rs[src];
                re_[u] = w.mr_re[src];
            }
        }
#pragma unroll
        for (int32_t u = 0; u < 4; u++)
            if (rb + u * AA_LANES < r1) {
                oi[u] = w.orig[gb[u]];
                fc[u] = w.first_call[gb[u]];
            }
#pragma unroll
        for (int32_t u = 0; u < 4; u++) {
            const int32_t r = rb + u * AA_LANES;
            if (r < r1) {
                const int64_t o = em.base + rows_before + (r - r0);
                w.r_idx[em.dst][o] = oi[u];
                w.r_qs[em.dst][o] = qs_[u];
                w.r_qe[em.dst][o] = qe_[u];
                w.r_rs[em.dst][o] = rs_[u];
                w.r_re[em.dst][o] = re_[u];
                w.r_alt[em.dst][o] = fc[u] > em.call ? 1 : 0;
            }
        }
    }
}

// One edge_path_to_paf_path call (paf_data.cpp:1489-1568) for an arbitrary planned walk: follow the main
// chain by prefix-sum differences (and block copies of its rows), simulate only around the walk's sidetracks.
// mark: record the blocks of the un-upgraded walk (pass A).  All lanes of the warp call this; lane 0 drives.
// SOLO: one thread per task (pass A only: nothing is copied), small private scratch; returns false when the task
// does not fit it (the caller hands it to the warp version).  !SOLO: one warp per task, lane 0 drives, all lanes copy.
template <bool SOLO>
AA_HDN bool walk_task_inc_t(const Ws &w, const Task &t, int32_t *side, int32_t side_cap, const DPBuf &buf, const Emit &em, bool mark,
                            int64_t *cov_out, int32_t *rows_out) {
    const int64_t c = t.ctg;
    Ctg g = ctg_view(w, c);
    const int64_t v0 = g.v0, e0 = w.eoff[v0], wo = w.walk_off[c];
    const int32_t *en = w.ent_node + 3 * wo;
    const int32_t *ep = w.ent_prev + 3 * wo;
    const bool lead = SOLO || aa_lane() == 0;
    int32_t ns = 0;
    if (lead)
        for (int32_t cur = w.wlast[wo + t.walk]; cur != -1; cur = ep[cur]) {
            if (SOLO && ns == side_cap) return false;
            side[ns++] = w.hn_eid[en[cur]];
        }
    int32_t si = ns - 1;  // side[] is last-to-first
    Auto A;
    A.cs = g.src;
    A.cov = 0;
    A.rows = 0;
    int32_t a = g.src, i = 0;
    bool onmain = true;
    for (;;) {
        // ---- lane 0 decides the next block copy [r0, r1) (or none) and whether the walk is finished ----
        int32_t r0 = 0, r1 = 0, rows_before = A.rows, finished = 0;
        if (lead) {
            if (onmain) {
                if (si < 0) {  // the rest is the main chain's suffix
                    r0 = w.m_rows[v0 + i];
                    r1 = w.m_tot_rows[c];
                    A.cov += w.m_tot_cov[c] - w.m_cov[v0 + i];
                    A.rows += r1 - r0;
                    finished = 1;
                } else {
                    const int32_t p = w.main_pos[v0 + w.e_src[e0 + side[si]]];
                    // main states at positions <= p-2 never look past the tail of the next sidetrack
                    int32_t j = p - 1 > i ? p - 1 : i;
                    if (w.m_cs[v0 + j] == -1) j++;
                    r0 = w.m_rows[v0 + i];
                    r1 = w.m_rows[v0 + j];
                    A.cov += w.m_cov[v0 + j] - w.m_cov[v0 + i];
                    A.rows += r1 - r0;
                    A.cs = w.m_cs[v0 + j];
                    a = w.main_walk[v0 + j];
                    i = j;
                    onmain = false;
                }
            }
        }
#if defined(__CUDA_ARCH__)
        if (!SOLO) {
            if (em.mode == 2) {
                r0 = __shfl_sync(0xffffffffu, r0, 0);
                r1 = __shfl_sync(0xffffffffu, r1, 0);
                rows_before = __shfl_sync(0xffffffffu, rows_before, 0);
            }
            finished = __shfl_sync(0xffffffffu, finished, 0);
        }
#endif
        if (!SOLO && r1 > r0) copy_main_rows(w, g, r0, r1, em, rows_before);
        if (finished) break;
        // ---- one simulated step from walk vertex a (lane 0) ----
        if (lead) {
            int32_t si1 = si, si2;
            int32_t v, nv = -1;
            if (si1 >= 0 && a == w.e_src[e0 + side[si1]]) v = e_dst(w.edge[e0 + side[si1--]]);
            else v = w.best[v0 + a];
            si2 = si1;
            if (v != g.dest) {
                if (si2 >= 0 && v == w.e_src[e0 + side[si2]]) nv = e_dst(w.edge[e0 + side[si2--]]);
                else nv = w.best[v0 + v];
            }
            const int32_t used = auto_step(w, g, buf, A, v, nv, em);
            if (SOLO && used < 0) return false;  // the gap-filling DP does not fit the private scratch
            if (mark && v != g.dest && w.main_pos[v0 + v] < 0) {
                int32_t x, y;
                vtx_xy(w, g, v, x, y);
                mark_block(w, g.b0 + x, t.call);
                mark_block(w, g.b0 + y, t.call);
            }
            if (v == g.dest) {
                finished = 1;
            } else {
                if (used == 2) {
                    if (mark && w.main_pos[v0 + nv] < 0) {
                        int32_t x, y;
                        vtx_xy(w, g, nv, x, y);
                        mark_block(w, g.b0 + x, t.call);
                        mark_block(w, g.b0 + y, t.call);
                    }
                    a = nv;
                    si = si2;
                } else {
                    a = v;
                    si = si1;
                }
                const int32_t mp = w.main_pos[v0 + a];
                if (mp >= 0 && w.m_cs[v0 + mp] == A.cs) {
                    onmain = true;
                    i = mp;
                }
            }
        }
#if defined(__CUDA_ARCH__)
        if (!SOLO) finished = __shfl_sync(0xffffffffu, finished, 0);
#endif
        if (finished) break;
    }
    if (lead) {
        if (cov_out) *cov_out = A.cov;
        if (rows_out) *rows_out = A.rows;
    }
    return true;
}
AA_HDN void walk_task_inc(const Ws &w, const Task &t, const Slot &s, const Emit &em, bool mark, int64_t *cov_out,
                          int32_t *rows_out) {
    walk_task_inc_t<false>(w, t, s.side, 0x7fffffff, s.buf, em, mark, cov_out, rows_out);
}

// worker loops (dynamic scheduling); slot = worker index.  All lanes run the loop, lane 0 pulls the work.
AA_HD int64_t next_item(const Ws &w) {
    unsigned long long k = 0;
#if defined(__CUDA_ARCH__)
    if (aa_lane() == 0) k = atomicAdd(w.task_next, 1ull);
    k = __shfl_sync(0xffffffffu, k, 0);
#else
    k = (*w.task_next)++;
#endif
    return (int64_t)k;
}
// pass A0/S2: resolve the main chain of every contig (largest contigs first through ord[])
AA_HDN void f_tasks_a0(const Ws &w, int64_t slot, const int32_t *ord) {
    Slot s = slot_view(w, slot);
    for (;;) {
        const int64_t k = next_item(w);
        if (k >= w.C) break;
#if defined(__CUDA_ARCH__)
        f_main_resolve_warp(w, ord[k], s);
#else
        f_main_resolve(w, ord[k], s);
#endif
    }
}
// pass A1: every other planned walk (coverage, row count, block marks).  The tasks are independent and each is a
// chain of dependent loads, so they run one per THREAD with a small private scratch; the few whose gap-filling DP or
// sidetrack list does not fit go to the warp version below.
constexpr int32_t SOLO_SIDE = 32;
AA_HDN void f_tasks_a1_solo(const Ws &w, int64_t t) {
    const Task tk = w.tasks[t];
    if (tk.call == 0) return;
    LocalDP loc;
    int32_t side[SOLO_SIDE];
    Emit em;
    em.mode = 0;
    em.dst = 0;
    em.base = 0;
    em.call = 0;
    if (!walk_task_inc_t<true>(w, tk, side, SOLO_SIDE, loc.buf(), em, true, &w.task_cov[t], &w.task_rows[t]))
        w.fb_list[aa_atomic_inc(w.fb_n)] = (int32_t)t;
}
AA_HDN void f_tasks_a1(const Ws &w, int64_t slot) {
    Slot s = slot_view(w, slot);
    Emit em;
    em.mode = 0;
    em.dst = 0;
    em.base = 0;
    em.call = 0;
    for (;;) {
        const int64_t k = next_item(w);
        if (k >= w.fb_count) break;
        const int64_t t = w.fb_list[k];
        const Task tk = w.tasks[t];
        walk_task_inc(w, tk, s, em, true, &w.task_cov[t], &w.task_rows[t]);
    }
}

// phase: selection (paf_data.cpp:1585-1649) from the per-task coverages
AA_HDN void f_select(const Ws &w, int64_t c, int32_t want_all) {
    if (aa_lane() != 0) return;
    w.win_out[c] = -1;
    w.win_alt[c] = -1;
    w.out_cnt[c] = 0;
    w.alt_cnt[c] = 0;
    w.all_cnt[c] = 0;
    if (w.status[c] == 1) {
        w.out_cnt[c] = 1;
        return;
    }
    if (w.status[c] != 0) return;
    const int64_t o = w.task_off[c];
    const int32_t nt = w.n_task[c], ntie = w.n_tie[c];
    int64_t best = -1;
    int32_t first = -1, equal_after = 0;
    for (int32_t k = 0; k < ntie; k++) {
        int64_t cv = w.task_cov[o + k];
        if (cv > best) {
            best = cv;
            first = k;
            equal_after = 0;
        } else if (cv == best) {
            equal_after++;
        }
    }
    w.win_out[c] = (int32_t)(o + first);
    w.out_cnt[c] = w.task_rows[o + first];
    if (want_all) w.all_cnt[c] = equal_after;
    const int32_t lg = w.last_group[c];
    if (lg > 0) {
        int64_t bc = -1;
        int32_t bk = -1;
        for (int32_t k = ntie; k < nt; k++) {
            if (w.tasks[o + k].group != lg) continue;
            int64_t cv = w.task_cov[o + k];
            if (bk < 0 || cv > bc) {
                bc = cv;
                bk = k;
            }
        }
        w.win_alt[c] = (int32_t)(o + bk);
        w.alt_cnt[c] = w.task_rows[o + bk];
    }
}
// list the .all paths of contig c (ties after the first maximum, in call order)
AA_HDN void f_all_list(const Ws &w, int64_t c) {
    if (aa_lane() != 0) return;
    if (w.all_cnt[c] == 0) return;
    const int64_t o = w.task_off[c];
    const int32_t ntie = w.n_tie[c];
    const int32_t first = w.win_out[c] - (int32_t)o;
    const int64_t best = w.task_cov[o + first];
    int64_t at = w.all_path_off[c];
    for (int32_t k = first + 1; k < ntie; k++)
        if (w.task_cov[o + k] == best) {
            w.all_task[at] = (int32_t)(o + k);
            w.all_rows[at] = w.task_rows[o + k];
            at++;
        }
}

// pass B work items: [0,C) primary chains, [C,2C) alt chains, [2C, 2C+Npaths) .all paths
AA_HDN void f_tasks_b(const Ws &w, int64_t slot, int64_t n_items, int64_t n_paths) {
    Slot s = slot_view(w, slot);
    (void)n_paths;
    for (;;) {
        const int64_t i = next_item(w);
        if (i >= n_items) break;
        Emit em;
        em.mode = 2;
        int64_t task = -1;
        if (i < w.C) {
            const int64_t c = i;
            if (w.status[c] == 1) {  // singleton contig (paf_data.cpp:235-239)
                if (aa_lane() == 0) {
                    int64_t b = w.ctg_off[c], o = w.out_off[c];
                    w.r_idx[0][o] = 0;
                    w.r_qs[0][o] = w.qs[b];
                    w.r_qe[0][o] = w.qe[b];
                    w.r_rs[0][o] = w.rs[b];
                    w.r_re[0][o] = w.re[b];
                    w.r_alt[0][o] = 0;
                }
                continue;
            }
            task = w.win_out[c];
            em.dst = 0;
            em.base = w.out_off[c];
        } else if (i < 2 * w.C) {
            const int64_t c = i - w.C;
            task = w.win_alt[c];
            em.dst = 1;
            em.base = w.alt_off[c];
        } else {
            const int64_t p = i - 2 * w.C;
            task = w.all_task[p];
            em.dst = 2;
            em.base = w.all_row_off[p];
        }
        if (task < 0) continue;
        const Task tk = w.tasks[task];
        em.call = tk.call;
        walk_task_inc(w, tk, s, em, false, nullptr, nullptr);
    }
}

}  // namespace aa
