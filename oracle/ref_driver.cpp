// oracle/ref_driver.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Driver around the UNMODIFIED reference hot path: it is compiled together with
// /root/reference/src/paf_data.cpp (sources stay where they lie; see oracle/Makefile) and restates the
// parts of /root/reference/src/alignasm.cpp that need the absent third-party libraries
// (csv-parser, argparse, TBB, indicators):
//   * the PAF reader + contig bucketing            alignasm.cpp:110-181
//   * --alt ingestion                             alignasm.cpp:186-332
//   * the solve loop over contigs                  alignasm.cpp:346-379   (std::thread pool stands in for TBB)
//   * the three writers                            alignasm.cpp:398-490
//
// Build variants (oracle/Makefile):
//   alignasm_ref        glibc malloc, -O3 -DNDEBUG: the CPU baseline that bench.py times.
//   alignasm_ref_canon  -DREF_BUMP_ALLOC: every solve_ctg_read call runs on a fresh ascending bump
//                       arena, which pins the reference's pointer-valued PQ tie-break
//                       (k_shortest_walks.hpp:231-247) to allocation order.
//                       This is the golden generator.
//   alignasm_ref_dump   canonical + the hook TU (ref_dump_tu.cpp) that dumps graph/d/best/walks.
#include "paf_data.hpp"

#include <chrono>
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>
#include <sys/mman.h>
#include <atomic>
#include <thread>

bool NON_SKIP_LINKABLE;  // alignasm.cpp:26

#ifdef REF_USE_B200
// the binding of INTEGRATION.md §2 (oracle/solve_batch_b200.cpp): replaces the solve loop below
void solve_all_contigs_b200(std::vector<std::vector<PafReadData>> &paf_data, std::vector<std::vector<PafOutputData>> &out,
                            std::vector<std::vector<PafOutputData>> &alt_out,
                            std::vector<std::vector<std::vector<PafOutputData>>> &max_out, int device);
#endif

#ifdef REF_DUMP
FILE *g_ref_dump_file = nullptr;  // consumed by ref_dump_tu.cpp
#endif

// ---------------------------------------------------------------------------------------------
// canonical allocator: ascending bump arena, reset before every contig
#ifdef REF_BUMP_ALLOC
namespace {
struct Arena {
    char *base = nullptr;
    size_t cap = 0, top = 0;
};
thread_local Arena t_arena;
void arena_init(Arena &a) {
    const char *gb = std::getenv("REF_ARENA_GB");
    a.cap = (size_t)(gb ? std::atof(gb) : 40.0) * (size_t(1) << 30);
    void *p = mmap(nullptr, a.cap, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (p == MAP_FAILED) {
        std::fputs("ref arena mmap failed\n", stderr);
        std::abort();
    }
    a.base = (char *)p;
    a.top = 0;
}
inline void *arena_alloc(size_t n, size_t align) {
    Arena &a = t_arena;
    if (!a.base) arena_init(a);
    size_t at = (a.top + align - 1) & ~(align - 1);
    if (at + n > a.cap) {
        std::fputs("ref arena exhausted (raise REF_ARENA_GB)\n", stderr);
        std::abort();
    }
    a.top = at + n;
    return a.base + at;
}
}  // namespace
void *operator new(size_t n) { return arena_alloc(n ? n : 1, 16); }
void *operator new[](size_t n) { return arena_alloc(n ? n : 1, 16); }
void *operator new(size_t n, std::align_val_t al) { return arena_alloc(n ? n : 1, (size_t)al); }
void *operator new[](size_t n, std::align_val_t al) { return arena_alloc(n ? n : 1, (size_t)al); }
void operator delete(void *) noexcept {}
void operator delete[](void *) noexcept {}
void operator delete(void *, size_t) noexcept {}
void operator delete[](void *, size_t) noexcept {}
void operator delete(void *, std::align_val_t) noexcept {}
void operator delete[](void *, std::align_val_t) noexcept {}
void operator delete(void *, size_t, std::align_val_t) noexcept {}
void operator delete[](void *, size_t, std::align_val_t) noexcept {}
#endif

namespace {

// rows kept outside operator new so the per-contig arena can be rewound
struct Row {
    int32_t ctg_index;
    int64_t qs, qe, rs, re;
    bool is_alt;
};
struct RowList {
    Row *p = nullptr;
    size_t n = 0;
    void assign(const std::vector<PafOutputData> &v) {
        n = v.size();
        p = (Row *)std::malloc(sizeof(Row) * (n ? n : 1));
        for (size_t i = 0; i < n; i++)
            p[i] = {v[i].ctg_index, v[i].edited_qry_str, v[i].edited_qry_end, v[i].edited_ref_str,
                    v[i].edited_ref_end, v[i].is_alt_path};
    }
};
struct RowListList {
    RowList *p = nullptr;
    size_t n = 0;
};

void split_tabs(std::string_view line, std::vector<std::string_view> &f) {
    f.clear();
    size_t s = 0;
    for (;;) {
        size_t e = line.find('\t', s);
        if (e == std::string_view::npos) {
            f.push_back(line.substr(s));
            return;
        }
        f.push_back(line.substr(s, e - s));
        s = e + 1;
    }
}
int64_t to_i64(std::string_view s) { return std::strtoll(std::string(s).c_str(), nullptr, 10); }

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

int main(int argc, char **argv) {
    std::string in_path, out_prefix, alt_path;
    double alt_baseline = 0.5;
    int threads = 1;
    bool no_write = false;
    int64_t limit_contigs = -1;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (a == "--non_skip_linkable") NON_SKIP_LINKABLE = true;
        else if ((a == "-t" || a == "--thread") && i + 1 < argc) threads = std::atoi(argv[++i]);
        else if ((a == "-a" || a == "--alt") && i + 1 < argc) alt_path = argv[++i];
        else if ((a == "-b" || a == "--alt_baseline") && i + 1 < argc) alt_baseline = std::atof(argv[++i]);
        else if (a == "--no-write") no_write = true;
        else if (a == "--out-prefix" && i + 1 < argc) out_prefix = argv[++i];
        else if (a == "--limit-contigs" && i + 1 < argc) limit_contigs = std::atoll(argv[++i]);
#ifdef REF_DUMP
        else if (a == "--dump" && i + 1 < argc) g_ref_dump_file = std::fopen(argv[++i], "w");
#endif
        else if (in_path.empty()) in_path = a;
        else {
            std::fprintf(stderr, "unknown argument %s\n", a.c_str());
            return 1;
        }
    }
    if (in_path.size() < 4 || in_path.substr(in_path.size() - 4) != ".paf") {
        std::fprintf(stderr, "Wrong PAF file : %s\n", in_path.c_str());
        return 1;
    }
    if (out_prefix.empty()) out_prefix = in_path.substr(0, in_path.size() - 4);

    double t0 = now_s();
    // ---- reader: alignasm.cpp:110-181 ----
    std::unordered_map<std::string, int32_t> chr_map;
    std::vector<std::string> chr_rev;
    std::vector<std::vector<PafReadData>> paf_data;
    std::vector<PafReadData> cur;
    std::vector<std::string> ctg_names;
    std::string ctg_chr;
    std::unordered_map<std::string, int32_t> paf_map;
    // one row -> PafReadData (alignasm.cpp:138-176 and 259-300 share this shape); q_off shifts the query interval
    auto fill_row = [&](const std::vector<std::string_view> &f, int64_t q_off, bool alt_row, PafReadData &d) -> bool {
        std::string ref_chr(f[PAF_REF_CHR]);
        if (!chr_map.count(ref_chr)) {
            chr_map[ref_chr] = (int32_t)chr_rev.size();
            chr_rev.push_back(ref_chr);
        }
        d.qry_str = to_i64(f[PAF_QRY_STR]) + q_off;
        d.qry_end = to_i64(f[PAF_QRY_END]) + q_off - 1;
        d.ref_total_length = to_i64(f[PAF_REF_TOT]);
        d.ref_str = to_i64(f[PAF_REF_STR]);
        d.ref_end = to_i64(f[PAF_REF_END]) - 1;
        d.ref_chr = chr_map[ref_chr];
        d.aln_fwd = f[PAF_ALN_FWD][0] == '+';
        if (!d.aln_fwd) std::swap(d.ref_str, d.ref_end);
        d.map_qul = (uint8_t)to_i64(f[PAF_MAT_QUL]);
        std::string_view cs;
        for (size_t k = PAF_MAT_QUL + 1; k < f.size(); k++)
            if (f[k].size() >= 5 && f[k].substr(0, 5) == "cs:Z:") {
                cs = f[k];
                break;
            }
        if (cs.empty()) {
            std::cerr << "Missing cs:Z tag in " << (alt_row ? "alternative " : "") << "PAF record for query '" << f[PAF_QRY_CHR] << "'\n";
            return false;
        }
        d.cs_string = cs;
        d.mat_num = (int32_t)to_i64(f[PAF_MAT_NUM]);
        d.aln_len = (int32_t)to_i64(f[PAF_ALN_LEN]);
        get_overlap_range(d, cs);
        return true;
    };
    {
        std::ifstream in(in_path);
        if (!in) {
            std::fprintf(stderr, "cannot open %s\n", in_path.c_str());
            return 1;
        }
        std::string line;
        std::vector<std::string_view> f;
        int32_t ctg_index = 0, paf_index = 0, row_global = 0;
        while (std::getline(in, line)) {
            if (!line.empty() && line.back() == '\r') line.pop_back();
            if (line.empty()) continue;
            split_tabs(line, f);
            if (f.size() < 12) {
                std::fprintf(stderr, "short PAF row\n");
                return 1;
            }
            std::string qry_chr(f[PAF_QRY_CHR]), ref_chr(f[PAF_REF_CHR]);
            if (ctg_chr.empty()) ctg_chr = qry_chr;
            if (ctg_chr != qry_chr) {
                paf_data.push_back(cur);
                ctg_names.push_back(ctg_chr);
                ctg_chr = qry_chr;
                ctg_index = 0;
                cur.clear();
                paf_index++;
            }
            PafReadData d{};
            paf_map[qry_chr] = paf_index;
            d.paf_index = paf_index;
            d.ctg_index = ctg_index++;
            d.qry_total_length = to_i64(f[PAF_QRY_TOT]);
            d.original_cord = {TYPE_MAIN, row_global++};
            if (!fill_row(f, 0, false, d)) return 1;
            cur.push_back(d);
        }
        ctg_names.push_back(ctg_chr);
        paf_data.push_back(cur);
    }
    // ---- --alt: alignasm.cpp:186-332 ----
    if (!alt_path.empty()) {
        if (alt_path.size() < 4 || alt_path.substr(alt_path.size() - 4) != ".paf") {
            std::fprintf(stderr, "Wrong PAF file : %s\n", alt_path.c_str());
            return 1;
        }
        std::ifstream in(alt_path);
        if (!in) {
            std::fprintf(stderr, "cannot open %s\n", alt_path.c_str());
            return 1;
        }
        std::string line, grp_ctg;
        std::vector<std::string_view> f;
        int64_t grp_off = -1;
        bool grp_open = false, grp_hit = false;
        double grp_ratio = 0;
        PafReadData grp_best{};
        auto close_group = [&]() {  // :247-255
            if (!grp_open || grp_hit) return;
            auto &dst = paf_data[(size_t)paf_map[grp_ctg]];
            grp_best.ctg_index = (int32_t)dst.size();
            dst.push_back(grp_best);
        };
        int32_t alt_row = 0;
        while (std::getline(in, line)) {
            if (!line.empty() && line.back() == '\r') line.pop_back();
            if (line.empty()) continue;
            split_tabs(line, f);
            if (f.size() < 12) {
                std::fprintf(stderr, "short PAF row\n");
                return 1;
            }
            // "<contig>:<START>-<END>" (:211-234)
            std::string qname(f[PAF_QRY_CHR]);
            size_t colon = qname.find(':');
            if (colon == std::string::npos) throw std::invalid_argument("Invalid input string format");
            size_t dash = qname.find('-', colon + 1);
            if (dash == std::string::npos) dash = qname.size();
            int64_t q_off = to_i64(std::string_view(qname).substr(colon + 1, dash - colon - 1)) - 1;
            std::string real = qname.substr(0, colon);
            auto &tail = paf_data[(size_t)paf_map[real]].back();
            PafReadData d{};
            d.paf_index = tail.paf_index;
            d.qry_total_length = tail.qry_total_length;
            d.original_cord = {TYPE_ALT, alt_row};
            if (!fill_row(f, q_off, true, d)) return 1;
            if (!grp_open || grp_off != q_off || grp_ctg != real) {
                close_group();
                grp_open = true;
                grp_hit = false;
                grp_ratio = 0;
                grp_off = q_off;
                grp_ctg = real;
                grp_best = {};
            }
            double ratio = std::atof(std::string(f[PAF_ALN_LEN]).c_str()) / std::atof(std::string(f[PAF_QRY_TOT]).c_str());
            if (ratio > grp_ratio) {
                grp_ratio = ratio;
                grp_best = d;
            }
            if (ratio > alt_baseline) {
                auto &dst = paf_data[(size_t)paf_map[real]];
                d.ctg_index = (int32_t)dst.size();
                dst.push_back(d);
                grp_hit = true;
            }
            alt_row++;
        }
        close_group();
    }
    if (limit_contigs >= 0 && (size_t)limit_contigs < paf_data.size()) {
        paf_data.resize((size_t)limit_contigs);
        ctg_names.resize((size_t)limit_contigs);
    }
    double t1 = now_s();
    int64_t n_ctg = (int64_t)paf_data.size(), n_blk = 0;
    for (auto &c : paf_data) n_blk += (int64_t)c.size();

    // ---- solve loop: alignasm.cpp:346-379 ----
    std::vector<RowList> out((size_t)n_ctg), alt((size_t)n_ctg);
    std::vector<RowListList> all((size_t)n_ctg);
    if (threads < 1) threads = 1;
    double t2 = now_s();
    // dynamic one-contig-at-a-time scheduling, standing in for tbb::parallel_for (alignasm.cpp:351-359)
    std::atomic<int64_t> next_ctg{0};
    auto worker = [&]() {
      for (;;) {
        int64_t i = next_ctg.fetch_add(1);
        if (i >= n_ctg) break;
#ifdef REF_BUMP_ALLOC
        if (!t_arena.base) arena_init(t_arena);
        size_t mark = t_arena.top;
#endif
        {
            std::vector<PafOutputData> o, a;
            std::vector<std::vector<PafOutputData>> m;
#ifdef REF_DUMP
            if (g_ref_dump_file) std::fprintf(g_ref_dump_file, "C %" PRId64 " %zu\n", i, paf_data[(size_t)i].size());
#endif
            solve_ctg_read(paf_data[(size_t)i], o, a, m);
            out[(size_t)i].assign(o);
            alt[(size_t)i].assign(a);
            all[(size_t)i].n = m.size();
            all[(size_t)i].p = (RowList *)std::calloc(m.size() ? m.size() : 1, sizeof(RowList));
            for (size_t k = 0; k < m.size(); k++) all[(size_t)i].p[k].assign(m[k]);
        }
#ifdef REF_BUMP_ALLOC
        t_arena.top = mark;
#endif
      }
    };
#ifdef REF_USE_B200
    {   // alignasm.cpp:346-397 replaced by ONE call into libalignasm_b200.so; containers pre-sized as alignasm.cpp:343-344 does
        (void)worker;
        std::vector<std::vector<PafOutputData>> o((size_t)n_ctg), a((size_t)n_ctg);
        std::vector<std::vector<std::vector<PafOutputData>>> m((size_t)n_ctg);
        solve_all_contigs_b200(paf_data, o, a, m, 0);
        for (int64_t i = 0; i < n_ctg; i++) {
            out[(size_t)i].assign(o[(size_t)i]);
            alt[(size_t)i].assign(a[(size_t)i]);
            all[(size_t)i].n = m[(size_t)i].size();
            all[(size_t)i].p = (RowList *)std::calloc(m[(size_t)i].size() ? m[(size_t)i].size() : 1, sizeof(RowList));
            for (size_t k = 0; k < m[(size_t)i].size(); k++) all[(size_t)i].p[k].assign(m[(size_t)i][k]);
        }
    }
#else
    if (threads == 1) {
        worker();
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; t++) pool.emplace_back(worker);
        for (auto &t : pool) t.join();
    }
#endif
    double t3 = now_s();

    // ---- writers: alignasm.cpp:398-490 ----
    if (!no_write) {
        auto write_row = [&](FILE *fp, size_t i, const std::string &qname, const Row &r) {
            PafReadData &src = paf_data[i][(size_t)r.ctg_index];
            PafOutputData po;
            po.ctg_index = r.ctg_index;
            po.edited_qry_str = r.qs;
            po.edited_qry_end = r.qe;
            po.edited_ref_str = r.rs;
            po.edited_ref_end = r.re;
            po.is_alt_path = r.is_alt;
            PafEditData e = get_edited_paf_data(po, src);
            std::fprintf(fp,
                         "%s\t%" PRId64 "\t%" PRId64 "\t%" PRId64 "\t%s\t%s\t%" PRId64 "\t%" PRId64 "\t%" PRId64
                         "\t%d\t%d\t%d\t%s\txi:Z:%s%d\t%s\n",
                         qname.c_str(), src.qry_total_length, r.qs, r.qe + 1, src.aln_fwd ? "+" : "-",
                         chr_rev[(size_t)src.ref_chr].c_str(), src.ref_total_length, src.aln_fwd ? r.rs : r.re,
                         (src.aln_fwd ? r.re : r.rs) + 1, e.mat_num, e.aln_len, (int)src.map_qul,
                         r.is_alt ? "tp:A:S" : "tp:A:P", src.original_cord.first == TYPE_MAIN ? "P_" : "A_",
                         src.original_cord.second, e.edit_cs_string.c_str());
        };
        FILE *f1 = std::fopen((out_prefix + ".aln.paf").c_str(), "w");
        FILE *f2 = std::fopen((out_prefix + ".aln.alt.paf").c_str(), "w");
        FILE *f3 = std::fopen((out_prefix + ".aln.all.paf").c_str(), "w");
        if (!f1 || !f2 || !f3) {
            std::fprintf(stderr, "cannot open outputs\n");
            return 1;
        }
        for (size_t i = 0; i < (size_t)n_ctg; i++) {
            for (size_t k = 0; k < out[i].n; k++) write_row(f1, i, ctg_names[i], out[i].p[k]);
            for (size_t k = 0; k < alt[i].n; k++) write_row(f2, i, ctg_names[i], alt[i].p[k]);
            for (size_t m = 0; m < all[i].n; m++) {
                std::string qn = ctg_names[i] + "." + std::to_string(m + 1);
                for (size_t k = 0; k < all[i].p[m].n; k++) write_row(f3, i, qn, all[i].p[m].p[k]);
            }
        }
        std::fclose(f1);
        std::fclose(f2);
        std::fclose(f3);
    }
    double t4 = now_s();
#ifdef REF_DUMP
    if (g_ref_dump_file) std::fclose(g_ref_dump_file);
#endif
    std::printf("{\"contigs\": %" PRId64 ", \"blocks\": %" PRId64 ", \"threads\": %d, \"read_s\": %.6f, "
                "\"solve_s\": %.6f, \"write_s\": %.6f}\n",
                n_ctg, n_blk, threads, t1 - t0, t3 - t2, t4 - t3);
    return 0;
}
