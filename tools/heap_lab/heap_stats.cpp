// tools/heap_lab/heap_stats.cpp — CPU analysis of one contig's sidetrack insert stream (dumped by an AA_HEAP_DUMP build):
// descent depths, swap levels, and how old the nodes are that the warp builder has to fetch on demand.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <map>
#include <algorithm>
struct VInfo { int32_t x; uint32_t ins_beg; int32_t nins; int32_t ppos; };
struct InsKey { int64_t sum; int32_t anom, nz, tot, eid; };
struct Node { int64_t sum; int32_t anom, nz, tot, left, right; int16_t rank, lrank; };
static inline int64_t den(int32_t t) { return t ? t : 1; }
static bool key_lt(const Node &a, const InsKey &k) {
    if (a.sum != k.sum) return a.sum < k.sum;
    if (a.anom != k.anom) return a.anom < k.anom;
    return (int64_t)a.nz * den(k.tot) > (int64_t)k.nz * den(a.tot);
}
int main(int argc, char **argv) {
    FILE *f = fopen(argv[1], "rb");
    int64_t hdr[4];
    fread(hdr, 8, 4, f);
    int64_t nt = hdr[0], ni = hdr[1];
    std::vector<VInfo> vi(nt);
    std::vector<InsKey> ins(ni);
    fread(vi.data(), sizeof(VInfo), nt, f);
    fread(ins.data(), sizeof(InsKey), ni, f);
    fclose(f);
    printf("nt %ld ins %ld V %ld\n", nt, ni, hdr[2]);
    std::vector<Node> hn;
    hn.reserve(ni * 8);
    std::vector<int32_t> root_at(nt, -1);
    std::map<int, int64_t> hp, hs, hage, hL, hpm;
    int64_t nswap = 0, nins = 0, switches = 0, leaf_ins = 0, zero = 0, newtop = 0;
    int32_t cur_root = -2; int fm = 0, famin = 0; std::vector<int> fsm, fa; int64_t fbad = 0;
    std::vector<int64_t> fetch_age;
    int64_t deep_fetch = 0;
    for (int64_t pos = 0; pos < nt; pos++) {
        int32_t root = vi[pos].ppos < 0 ? -1 : root_at[vi[pos].ppos];
        int n = vi[pos].nins & ((1 << 30) - 1);
        bool kids = vi[pos].nins & (1 << 30);
        if (n > 0 && !kids && n <= 32) { leaf_ins += n; }
        if (n > 0 && root != cur_root) switches++;
        for (int k = 0; k < n; k++) {
            const InsKey &ik = ins[vi[pos].ins_beg + k];
            if (ik.sum == 0 && ik.anom == 0) zero++;
            // descent
            std::vector<int32_t> sp;
            int32_t a = root;
            while (a >= 0 && key_lt(hn[a], ik)) { sp.push_back(a); a = hn[a].right; }
            int p = (int)sp.size();
            if (p == 0) newtop++;
            hp[p]++;
            nins++;
            Node nn{ik.sum, ik.anom, ik.nz, ik.tot, a, -1, 1, (int16_t)(a >= 0 ? hn[a].rank : 0)};
            int32_t r = (int32_t)hn.size();
            hn.push_back(nn);
            // formula check (f_heaps_warp2): a_j, last argmin, suffix minima
            {
                std::vector<int> a(p + 1);
                for (int q = 0; q < p; q++) a[q] = (hn[sp[q]].left < 0 ? -1 : hn[sp[q]].lrank) + q + 1;
                a[p] = p + 1;
                int amin = 1 << 30, m = -1;
                for (int q = 0; q <= p; q++) if (a[q] <= amin) { amin = a[q]; m = q; }
                fm = m; famin = amin;
                fsm.assign(p + 2, 1 << 30);
                for (int q = p; q >= 0; q--) fsm[q] = std::min(a[q], fsm[q + 1]);
                fa = a;
            }
            int rr = 1, topswap = -1;
            for (int q = p - 1; q >= 0; q--) {
                Node o = hn[sp[q]];
                int32_t l = o.left, lr = o.lrank, rc = r, rk = rr;
                if (l < 0 || lr < rk) { std::swap(l, rc); std::swap(lr, rk); topswap = q; }
                o.left = l; o.right = rc; o.lrank = lr; o.rank = rc >= 0 ? rk + 1 : 0;
                r = (int32_t)hn.size(); hn.push_back(o); rr = o.rank;
                const bool fswap = fa[q] < fsm[q + 1];
                const int frank = fsm[q] - q, flr = fswap ? fsm[q + 1] - (q + 1) : hn[sp[q]].lrank;
                if (fswap != (topswap == q && true) && false) {}
                if (frank != o.rank || flr != o.lrank || fswap != (o.left == r - 1 && (q == p - 1 ? true : true) && (hn[sp[q]].left != o.left))) { if (fbad++ < 10) printf("formula mismatch ins %ld q %d: rank %d/%d lrank %d/%d swap %d\n", nins, q, frank, o.rank, flr, o.lrank, (int)fswap); }
            }
            if ((topswap < 0 ? p : topswap) != fm) { if (fbad++ < 10) printf("m mismatch ins %ld: %d vs %d\n", nins, topswap, fm); }
            hpm[p - (topswap < 0 ? p : topswap)]++;
            if (topswap >= 0) { nswap++; hs[topswap]++;
                int32_t nr = hn[r].right; // walk to find
                // the node that becomes spine level topswap+1:
                int32_t x = r; for (int q = 0; q < topswap; q++) x = hn[x].right;
                int32_t L = hn[x].right;
                if (L >= 0) { int64_t age = (int64_t)hn.size() - L; int b = 0; while ((1ll << b) < age) b++; hage[b]++; }
                (void)nr;
            }
            root = r;
            if (!(n > 0 && !kids)) cur_root = root;
        }
        root_at[pos] = root;
    }
    printf("inserts %ld (leaf %ld) nodes %zu  per insert %.2f  switches %ld  zero-keys %ld newtop %ld swaps %ld\n", nins, leaf_ins, hn.size(), (double)hn.size() / nins, switches, zero, newtop, nswap);
    printf("formula mismatches: %ld\n", fbad);
    printf("p hist:"); for (auto &e : hp) printf(" %d:%ld", e.first, e.second); printf("\n");
    printf("p - m hist:"); for (auto &e : hpm) printf(" %d:%ld", e.first, e.second); printf("\n");
    printf("topswap level hist:"); for (auto &e : hs) printf(" %d:%ld", e.first, e.second); printf("\n");
    printf("age(log2 nodes) of node entering spine after swap:"); for (auto &e : hage) printf(" %d:%ld", e.first, e.second); printf("\n");
    // spine length of final versions
    return 0;
}
