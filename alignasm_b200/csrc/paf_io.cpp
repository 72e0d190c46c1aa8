// paf_io.cpp — host side of the drop-in: PAF reader, contig bucketing, cs:Z: codec, writers.
//
// Restates, without csv-parser/argparse, the host code either side of the hot path:
//   reader + bucketing            reference src/alignasm.cpp:110-181
//   cs:Z: -> exact-match runs     reference src/paf_data.cpp:29-123   (parse_short_cs, get_overlap_range)
//   cs:Z: re-cut for output       reference src/paf_data.cpp:125-220  (get_edited_paf_data)
//   the three writers             reference src/alignasm.cpp:398-490
// The parsed batch is structure-of-arrays from the start (the layout the kernels consume), not an
// array of row objects.
#include "../../include/alignasm_b200.h"

#include <cctype>
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <charconv>
#include <chrono>
#include <memory>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

// std::vector whose resize() leaves trivially-constructible elements uninitialised: the joined tables are sized once and
// filled by all host threads, so the pages are first touched in parallel instead of being zeroed by one thread
template <class T>
struct NoInit : std::allocator<T> {
    template <class U>
    struct rebind {
        using other = NoInit<U>;
    };
    template <class U>
    void construct(U *p) noexcept {
        ::new ((void *)p) U;
    }
    template <class U, class... A>
    void construct(U *p, A &&...a) {
        ::new ((void *)p) U(std::forward<A>(a)...);
    }
};
template <class T>
using Vec = std::vector<T, NoInit<T>>;

// a PAF file as read: mapped when possible, else read into memory
struct Image {
    const char *data = nullptr;
    size_t size = 0;
    bool mapped = false;
    std::string own;
    Image() = default;
    Image(const Image &) = delete;
    Image &operator=(const Image &) = delete;
    ~Image() {
        if (mapped) munmap(const_cast<char *>(data), size);
    }
    bool open(const char *path) {
        int fd = ::open(path, O_RDONLY);
        if (fd < 0) return false;
        struct stat sb;
        if (fstat(fd, &sb) == 0 && S_ISREG(sb.st_mode) && sb.st_size > 0) {
            void *m = mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (m != MAP_FAILED) {
                data = (const char *)m;
                size = (size_t)sb.st_size;
                mapped = true;
                ::close(fd);
                return true;
            }
        }
        char buf[1 << 16];
        ssize_t got;
        while ((got = ::read(fd, buf, sizeof buf)) > 0) own.append(buf, (size_t)got);
        ::close(fd);
        data = own.data();
        size = own.size();
        return true;
    }
};

// one table of parsed rows, structure-of-arrays (the main PAF, and the rows of the alternative PAF before the merge)
struct Rows {
    Vec<int64_t> qs, qe, rs, re, qtot, rtot, run_off{0}, run_ql, run_qr, run_rl;
    Vec<int32_t> chr, mat_num, aln_len, orig_idx;
    Vec<uint8_t> fwd, mapq, orig_alt;  // orig_alt: TYPE_ALT row (xi:Z:A_<row>), else TYPE_MAIN (xi:Z:P_<row>)
    Vec<std::string_view> cs;          // original cs:Z: field per row (a view into the file image; re-cut for output)
    size_t size() const { return qs.size(); }
};

// target names in order of first appearance (chr_map / chr_rev_map, alignasm.cpp:119-123)
struct ChrTable {
    std::unordered_map<std::string, int32_t> map;
    std::vector<std::string> names;
    int32_t id(std::string_view name) {
        std::string key(name);
        auto it = map.find(key);
        if (it != map.end()) return it->second;
        const int32_t k = (int32_t)names.size();
        map.emplace(key, k);
        names.push_back(std::move(key));
        return k;
    }
};

struct aa_paf {
    aa_batch batch{};
    Rows r;
    std::vector<int64_t> ctg_off;
    std::vector<std::string> ctg_name;
    ChrTable chrs;
    std::unordered_map<std::string, int32_t> paf_map;  // query name -> last bucket of that name (alignasm.cpp:136)
    std::vector<std::unique_ptr<Image>> images;  // the files as read: Rows::cs points into them
    void bind();
};

namespace {

void set_err(char *err, int64_t cap, const std::string &msg) {
    if (err && cap > 0) {
        std::snprintf(err, (size_t)cap, "%s", msg.c_str());
    }
}

inline bool is_alpha(char c) { return std::isalpha((unsigned char)c) != 0; }

struct CsOp {
    char type;
    int64_t len;
    uint32_t at, n;  // slice of the cs string holding the op text
};

// parse_short_cs (paf_data.cpp:29-72); returns false with a message on the reference's throw sites
bool parse_cs(std::string_view cs, std::vector<CsOp> &ops, std::string &why) {
    ops.clear();
    if (cs.size() < 5 || cs.substr(0, 5) != "cs:Z:") {
        why = "PAF record does not contain a short-form cs:Z tag";
        return false;
    }
    size_t pos = 5;
    while (pos < cs.size()) {
        size_t start = pos;
        char t = cs[pos++];
        int64_t len = 0;
        if (t == ':') {
            size_t p0 = pos;
            bool neg = false;
            if (pos < cs.size() && cs[pos] == '-') {  // std::from_chars accepts a sign; value must be > 0 anyway
                neg = true;
                pos++;
            }
            size_t d0 = pos;
            while (pos < cs.size() && cs[pos] >= '0' && cs[pos] <= '9') {
                len = len * 10 + (cs[pos] - '0');
                pos++;
            }
            if (pos == d0 || neg || len <= 0) {
                (void)p0;
                why = "Invalid :length operation in cs tag";
                return false;
            }
        } else if (t == '*') {
            if (pos + 2 > cs.size() || !is_alpha(cs[pos]) || !is_alpha(cs[pos + 1])) {
                why = "Invalid substitution operation in cs tag";
                return false;
            }
            pos += 2;
            len = 1;
        } else if (t == '+' || t == '-') {
            size_t s0 = pos;
            while (pos < cs.size() && is_alpha(cs[pos])) pos++;
            len = (int64_t)(pos - s0);
            if (len == 0) {
                why = "Empty indel operation in cs tag";
                return false;
            }
        } else {
            why = "Unsupported operation in short-form cs tag";
            return false;
        }
        ops.push_back({t, len, (uint32_t)start, (uint32_t)(pos - start)});
    }
    return true;
}

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

bool to_i64(std::string_view s, int64_t &v) {
    if (s.empty()) return false;
    size_t i = 0;
    bool neg = false;
    if (s[0] == '-' || s[0] == '+') {
        neg = s[0] == '-';
        i = 1;
    }
    if (i >= s.size()) return false;
    int64_t x = 0;
    for (; i < s.size(); i++) {
        if (s[i] < '0' || s[i] > '9') return false;
        x = x * 10 + (s[i] - '0');
    }
    v = neg ? -x : x;
    return true;
}

}  // namespace

namespace {

// One PAF row -> one entry of `t` (alignasm.cpp:138-176 / 270-300): closed intervals, ref_str > ref_end on the minus
// strand, exact-match runs from the cs tag (get_overlap_range, paf_data.cpp:90-123).  `q_off` shifts the query
// coordinates (rows of the alternative PAF are relative to their `ctg:START-END` segment, alignasm.cpp:266-268).
aa_status parse_row(const std::vector<std::string_view> &f, int64_t row, int64_t q_off, const char *what, ChrTable &chrs, Rows &t,
                    std::vector<CsOp> &ops, std::string &why, bool defer_cs = false) {
    if (f.size() < 12) {
        why = std::string(what) + " row " + std::to_string(row) + " has fewer than 12 columns";
        return AA_ERR_FORMAT;
    }
    const int32_t chr_id = chrs.id(f[5]);
    int64_t qtot, qs, qe, rtot, rs, re, mat, aln, mq;
    if (!to_i64(f[1], qtot) || !to_i64(f[2], qs) || !to_i64(f[3], qe) || !to_i64(f[6], rtot) || !to_i64(f[7], rs) ||
        !to_i64(f[8], re) || !to_i64(f[9], mat) || !to_i64(f[10], aln) || !to_i64(f[11], mq) || f[4].empty()) {
        why = std::string(what) + " row " + std::to_string(row) + ": non-numeric coordinate field";
        return AA_ERR_FORMAT;
    }
    qs += q_off;
    qe += q_off - 1;  // closed intervals (alignasm.cpp:143-150)
    re -= 1;
    bool fwd = f[4][0] == '+';
    if (!fwd) std::swap(rs, re);  // alignasm.cpp:155-159
    std::string_view cs;
    for (size_t k = 12; k < f.size(); k++)
        if (f[k].size() >= 5 && f[k].substr(0, 5) == "cs:Z:") {
            cs = f[k];
            break;
        }
    if (cs.empty()) {  // alignasm.cpp:165-168, 288-291
        why = std::string("Missing cs:Z tag in ") + (what[0] == 'a' ? "alternative " : "") + "PAF record for query '" +
              std::string(f[0]) + "'";
        return AA_ERR_FORMAT;
    }
    // get_overlap_range (paf_data.cpp:90-123): walk the ops in query orientation
    // (defer_cs: aa_paf_read_device leaves this to the device codec, csrc/cs_codec.cu, once all rows are known)
    if (!defer_cs && !parse_cs(cs, ops, why)) return AA_ERR_FORMAT;
    int64_t step = fwd ? 1 : -1, ri = rs, qi = qs;
    size_t nop = defer_cs ? 0 : ops.size();
    for (size_t k = 0; k < nop; k++) {
        const CsOp &o = fwd ? ops[k] : ops[nop - 1 - k];
        if (o.type == ':') {
            t.run_ql.push_back(qi);
            t.run_qr.push_back(qi + o.len - 1);
            t.run_rl.push_back(ri);
            ri += o.len * step;
            qi += o.len;
        } else if (o.type == '+') {
            qi += o.len;
        } else if (o.type == '-') {
            ri += o.len * step;
        } else {
            ri += step;
            qi += 1;
        }
    }
    if (!defer_cs && (qi != qe + 1 || ri != re + step)) {
        t.run_ql.resize((size_t)t.run_off.back());
        t.run_qr.resize((size_t)t.run_off.back());
        t.run_rl.resize((size_t)t.run_off.back());
        why = "cs tag consumption does not match PAF coordinates (" + std::string(what) + " row " + std::to_string(row) + ")";
        return AA_ERR_FORMAT;
    }
    t.run_off.push_back((int64_t)t.run_ql.size());
    t.qs.push_back(qs);
    t.qe.push_back(qe);
    t.rs.push_back(rs);
    t.re.push_back(re);
    t.qtot.push_back(qtot);
    t.rtot.push_back(rtot);
    t.chr.push_back(chr_id);
    t.fwd.push_back(fwd ? 1 : 0);
    t.mapq.push_back((uint8_t)mq);
    t.mat_num.push_back((int32_t)mat);
    t.aln_len.push_back((int32_t)aln);
    t.cs.push_back(cs);
    return AA_OK;
}

// next non-empty line of [pos, end): trailing '\r' / '\n' stripped; false at the end
inline bool next_line(const char *base, size_t &pos, size_t end, std::string_view &line) {
    while (pos < end) {
        const char *nl = (const char *)std::memchr(base + pos, '\n', end - pos);
        size_t stop = nl ? (size_t)(nl - base) : end, len = stop - pos;
        while (len > 0 && (base[pos + len - 1] == '\r' || base[pos + len - 1] == '\n')) len--;
        line = std::string_view(base + pos, len);
        pos = nl ? stop + 1 : end;
        if (len > 0) return true;
    }
    return false;
}

int host_threads(size_t bytes) {
    if (bytes < (1u << 20)) return 1;
    int n = 0;
    if (const char *e = std::getenv("AA_HOST_THREADS")) n = std::atoi(e);
    if (n <= 0) n = (int)std::thread::hardware_concurrency();
    if (n <= 0) n = 1;
    return n > 64 ? 64 : n;
}

// AA_IO_TRACE=1: stage times of the reader / writer on stderr
struct IoTrace {
    bool on = std::getenv("AA_IO_TRACE") != nullptr;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    void mark(const char *what) {
        if (!on) return;
        auto t1 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[aa_io] %-18s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

template <class F>
void run_parallel(int n, F f) {
    if (n <= 1) {
        f(0);
        return;
    }
    std::vector<std::thread> pool;
    for (int t = 1; t < n; t++) pool.emplace_back(f, t);
    f(0);
    for (auto &t : pool) t.join();
}

// what one reader thread produces from its slice of the file
struct Chunk {
    size_t beg = 0, end = 0;
    int64_t row_base = 0, n_rows = 0;
    Rows rows;
    ChrTable chrs;                                             // slice-local target ids
    std::vector<std::pair<int64_t, std::string_view>> names;  // (local row, query name) wherever the name changes
    aa_status st = AA_OK;
    std::string why;
};

// the reference's `aln_len / qry_total` as doubles (csv-parser get<double>, alignasm.cpp:310)
bool to_f64(std::string_view s, double &v) {
    std::string tmp(s);
    char *e = nullptr;
    v = std::strtod(tmp.c_str(), &e);
    return e && e != tmp.c_str() && *e == 0;
}

}  // namespace

void aa_paf::bind() {
    aa_batch &b = batch;
    b.n_ctg = (int64_t)ctg_off.size() - 1;
    b.n_blk = (int64_t)r.size();
    b.n_run = (int64_t)r.run_ql.size();
    b.ctg_off = ctg_off.data();
    b.qry_str = r.qs.data();
    b.qry_end = r.qe.data();
    b.ref_str = r.rs.data();
    b.ref_end = r.re.data();
    b.qry_total = r.qtot.data();
    b.ref_chr = r.chr.data();
    b.aln_fwd = r.fwd.data();
    b.map_qul = r.mapq.data();
    b.run_off = r.run_off.data();
    b.run_ql = r.run_ql.data();
    b.run_qr = r.run_qr.data();
    b.run_rl = r.run_rl.data();
}

extern "C" {

static aa_status paf_read_impl(const char *path, aa_ctx *ctx, aa_paf **out, char *err, int64_t err_cap) {
    if (!path || !out) return AA_ERR_INVALID;
    *out = nullptr;
    const bool defer_cs = ctx != nullptr;
    IoTrace tr;
    auto img = std::make_unique<Image>();
    if (!img->open(path)) {
        set_err(err, err_cap, std::string("cannot open ") + path);
        return AA_ERR_IO;
    }
    tr.mark("read file");
    // The file is cut into one slice per host thread at line boundaries.  Every slice is parsed on its own (rows, runs,
    // slice-local target ids, places where the query name changes); the slices are then joined in file order, which
    // gives the row numbers, target ids (first appearance) and contig buckets (change of name, alignasm.cpp:115-133)
    // of a sequential read, and the first error in file order.
    const char *base = img->data;
    const size_t size = img->size;
    const int T = host_threads(size);
    std::vector<Chunk> ch((size_t)T);
    for (int t = 0; t < T; t++) {
        size_t b = size * (size_t)t / (size_t)T;
        if (t > 0 && b < size) {
            const char *nl = (const char *)std::memchr(base + b - 1, '\n', size - b + 1);
            b = nl ? (size_t)(nl - base) + 1 : size;
        }
        ch[(size_t)t].beg = b;
        if (t > 0) ch[(size_t)t - 1].end = b;
    }
    ch[(size_t)T - 1].end = size;
    run_parallel(T, [&](int t) {  // pass 1: rows per slice
        Chunk &c = ch[(size_t)t];
        size_t pos = c.beg;
        std::string_view line;
        while (next_line(base, pos, c.end, line)) c.n_rows++;
    });
    for (int t = 1; t < T; t++) ch[(size_t)t].row_base = ch[(size_t)t - 1].row_base + ch[(size_t)t - 1].n_rows;
    tr.mark("count rows");
    run_parallel(T, [&](int t) {  // pass 2: parse
        Chunk &c = ch[(size_t)t];
        std::vector<std::string_view> f;
        std::vector<CsOp> ops;
        std::string_view line, cur;
        size_t pos = c.beg;
        int64_t row = 0;
        c.rows.qs.reserve((size_t)c.n_rows);
        while (next_line(base, pos, c.end, line)) {
            split_tabs(line, f);
            c.st = parse_row(f, c.row_base + row, 0, "PAF", c.chrs, c.rows, ops, c.why, defer_cs);
            if (c.st != AA_OK) return;
            if (row == 0 || cur != f[0]) {
                cur = f[0];
                c.names.emplace_back(row, cur);
            }
            row++;
        }
    });
    tr.mark("parse");
    const int64_t n_rows = ch[(size_t)T - 1].row_base + ch[(size_t)T - 1].n_rows;
    for (int t = 0; t < T; t++)
        if (ch[(size_t)t].st != AA_OK) {
            set_err(err, err_cap, ch[(size_t)t].why);
            return ch[(size_t)t].st;
        }
    if (n_rows == 0) {
        set_err(err, err_cap, "PAF file holds no rows");
        return AA_ERR_FORMAT;
    }
    aa_paf *p = new aa_paf();
    std::vector<std::vector<int32_t>> remap((size_t)T);
    std::vector<int64_t> run_base((size_t)T + 1, 0);
    std::string_view cur;
    for (int t = 0; t < T; t++) {
        Chunk &c = ch[(size_t)t];
        for (const std::string &name : c.chrs.names) remap[(size_t)t].push_back(p->chrs.id(name));
        // bucket by *change* of the query name: a name that re-appears later opens a new contig, as in the reference
        for (auto &nm : c.names)
            if (p->ctg_off.empty() || cur != nm.second) {
                cur = nm.second;
                p->ctg_off.push_back(c.row_base + nm.first);
                p->ctg_name.emplace_back(cur);
                p->paf_map[p->ctg_name.back()] = (int32_t)p->ctg_off.size() - 1;
            }
        run_base[(size_t)t + 1] = run_base[(size_t)t] + (int64_t)c.rows.run_ql.size();
    }
    p->ctg_off.push_back(n_rows);
    Rows &r = p->r;
    const size_t N = (size_t)n_rows, R = (size_t)run_base[(size_t)T];
    for (auto *v : {&r.qs, &r.qe, &r.rs, &r.re, &r.qtot, &r.rtot}) v->resize(N);
    for (auto *v : {&r.run_ql, &r.run_qr, &r.run_rl}) v->resize(R);
    for (auto *v : {&r.chr, &r.mat_num, &r.aln_len, &r.orig_idx}) v->resize(N);
    for (auto *v : {&r.fwd, &r.mapq, &r.orig_alt}) v->resize(N);
    r.run_off.resize(N + 1);
    r.cs.resize(N);
    run_parallel(T, [&](int t) {
        const Chunk &c = ch[(size_t)t];
        const Rows &s = c.rows;
        const size_t at = (size_t)c.row_base, n = s.size(), rat = (size_t)run_base[(size_t)t], rn = s.run_ql.size();
        auto put = [&](auto &dst, const auto &src, size_t where, size_t cnt) {
            if (cnt) std::memcpy(dst.data() + where, src.data(), cnt * sizeof(src[0]));
        };
        put(r.qs, s.qs, at, n);
        put(r.qe, s.qe, at, n);
        put(r.rs, s.rs, at, n);
        put(r.re, s.re, at, n);
        put(r.qtot, s.qtot, at, n);
        put(r.rtot, s.rtot, at, n);
        put(r.mat_num, s.mat_num, at, n);
        put(r.aln_len, s.aln_len, at, n);
        put(r.fwd, s.fwd, at, n);
        put(r.mapq, s.mapq, at, n);
        put(r.cs, s.cs, at, n);
        put(r.run_ql, s.run_ql, rat, rn);
        put(r.run_qr, s.run_qr, rat, rn);
        put(r.run_rl, s.run_rl, rat, rn);
        for (size_t i = 0; i < n; i++) {
            r.chr[at + i] = remap[(size_t)t][(size_t)s.chr[i]];
            r.orig_idx[at + i] = (int32_t)(at + i);  // original_cord = {TYPE_MAIN, row_global_index} (alignasm.cpp:172)
            r.orig_alt[at + i] = 0;
            r.run_off[at + i] = s.run_off[i] + (int64_t)rat;
        }
    });
    r.run_off[N] = (int64_t)R;
    tr.mark("join");
    if (defer_cs) {  // parse_short_cs + get_overlap_range on the device, over the file image as it lies
        std::vector<int64_t> cs_off(N);
        std::vector<int32_t> cs_len(N);
        for (size_t i = 0; i < N; i++) {
            cs_off[i] = (int64_t)(r.cs[i].data() - base);
            cs_len[i] = (int32_t)r.cs[i].size();
        }
        aa_cs_rows rows{(int64_t)N, cs_off.data(), cs_len.data(), r.qs.data(), r.qe.data(), r.rs.data(), r.re.data(), r.fwd.data()};
        aa_cs_runs runs{};
        aa_status st = aa_cs_runs_device(ctx, base, (int64_t)size, &rows, &runs);
        if (st != AA_OK) {
            set_err(err, err_cap, std::string("cs codec on the device: ") + aa_cs_last_error());
            delete p;
            return st;
        }
        for (size_t i = 0; i < N; i++)
            if (runs.err[i]) {  // the first row in error, in file order, with the reference's text
                std::string why = aa_cs_error_text(runs.err[i]);
                if (runs.err[i] == AA_CS_ERR_CONSUME) why += " (PAF row " + std::to_string(i) + ")";
                set_err(err, err_cap, why);
                aa_cs_runs_free(&runs);
                delete p;
                return AA_ERR_FORMAT;
            }
        r.run_ql.assign(runs.run_ql, runs.run_ql + runs.n_run);
        r.run_qr.assign(runs.run_qr, runs.run_qr + runs.n_run);
        r.run_rl.assign(runs.run_rl, runs.run_rl + runs.n_run);
        std::memcpy(r.run_off.data(), runs.run_off, (N + 1) * sizeof(int64_t));
        aa_cs_runs_free(&runs);
        tr.mark("cs runs (device)");
    }
    p->images.push_back(std::move(img));
    p->bind();
    *out = p;
    return AA_OK;
}
aa_status aa_paf_read(const char *path, aa_paf **out, char *err, int64_t err_cap) { return paf_read_impl(path, nullptr, out, err, err_cap); }
aa_status aa_paf_read_device(const char *path, aa_ctx *ctx, aa_paf **out, char *err, int64_t err_cap) {
    if (!ctx) return AA_ERR_INVALID;
    return paf_read_impl(path, ctx, out, err, err_cap);
}

// --alt ingestion (alignasm.cpp:186-332).  Rows of the alternative PAF are alignments of contig segments named
// `<contig>:<START>-<END>`; their query coordinates are shifted by START-1 and they are appended to that contig's
// blocks after its own rows.  Consecutive rows of one segment form a group: every row whose aln_len / segment length
// exceeds `alt_baseline` is taken, and a group without such a row contributes its best-ratio row (alignasm.cpp:247-255,
// 302-327).  A segment of an unknown contig lands in bucket 0, as `paf_map[name]` default-constructs in the reference.
aa_status aa_paf_read_alt(aa_paf *p, const char *alt_path, double alt_baseline, char *err, int64_t err_cap) {
    if (!p || !alt_path) return AA_ERR_INVALID;
    std::string ap(alt_path);
    if (ap.size() < 4 || ap.compare(ap.size() - 4, 4, ".paf") != 0) {  // alignasm.cpp:191-195
        set_err(err, err_cap, "Wrong PAF file : \"" + ap + "\"");
        return AA_ERR_INVALID;
    }
    auto img = std::make_unique<Image>();
    if (!img->open(alt_path)) {
        set_err(err, err_cap, std::string("cannot open ") + alt_path);
        return AA_ERR_IO;
    }
    Rows alt;
    std::vector<std::pair<int32_t, int32_t>> taken;  // (alt row, contig), in the reference's push_back order
    std::vector<std::string_view> f;
    std::vector<CsOp> ops;
    std::string why, seg_ctg;
    int64_t seg_off = -1, best_row = -1;
    bool grouped = false, group_took = false;
    double best_ratio = 0;
    auto flush = [&]() {  // flush_alt_group (alignasm.cpp:247-255)
        if (!grouped || group_took) return;
        if (best_row >= 0) taken.emplace_back((int32_t)best_row, p->paf_map[seg_ctg]);
    };
    int64_t row = 0;
    aa_status st = AA_OK;
    std::string_view lv;
    size_t pos = 0;
    while (next_line(img->data, pos, img->size, lv)) {
        split_tabs(lv, f);
        // parseString (alignasm.cpp:211-234): "<contig>:<START>[-...]" -> (contig, START - 1)
        std::string_view qn = f[0];
        size_t colon = qn.find(':');
        int64_t start1 = 0;
        bool ok = colon != std::string_view::npos;
        if (ok) {
            size_t dash = qn.find('-', colon + 1);
            std::string_view num = qn.substr(colon + 1, (dash == std::string_view::npos ? qn.size() : dash) - colon - 1);
            ok = !num.empty() && num[0] != '+' && to_i64(num, start1);
        }
        if (!ok) {
            set_err(err, err_cap, "alternative PAF row " + std::to_string(row) + ": query name is not <contig>:<start>-<end>");
            st = AA_ERR_FORMAT;
            break;
        }
        std::string real(qn.substr(0, colon));
        int64_t q_off = start1 - 1;
        int32_t c = p->paf_map[real];
        st = parse_row(f, row, q_off, "alternative PAF", p->chrs, alt, ops, why);
        double aln_len, seg_len;
        if (st == AA_OK && (!to_f64(f[10], aln_len) || !to_f64(f[1], seg_len))) {
            why = "alternative PAF row " + std::to_string(row) + ": non-numeric length field";
            st = AA_ERR_FORMAT;
        }
        if (st != AA_OK) {
            set_err(err, err_cap, why);
            break;
        }
        // the block inherits the contig's total length (alignasm.cpp:262-265); the segment length only enters the ratio
        alt.qtot.back() = p->r.qtot[(size_t)p->ctg_off[(size_t)c + 1] - 1];
        alt.orig_idx.push_back((int32_t)row);  // original_cord = {TYPE_ALT, row_global_index} (alignasm.cpp:297)
        alt.orig_alt.push_back(1);
        if (!grouped || seg_off != q_off || seg_ctg != real) {
            flush();
            grouped = true;
            group_took = false;
            best_ratio = 0;
            best_row = -1;
            seg_off = q_off;
            seg_ctg = real;
        }
        double ratio = aln_len / seg_len;
        if (ratio > best_ratio) {
            best_ratio = ratio;
            best_row = row;
        }
        if (ratio > alt_baseline) {
            taken.emplace_back((int32_t)row, c);
            group_took = true;
        }
        row++;
    }
    if (st != AA_OK) return st;
    flush();
    if (taken.empty()) return AA_OK;
    p->images.push_back(std::move(img));
    // merge: every contig keeps its own rows, then the taken rows in the order the reference appended them
    int64_t n_ctg = (int64_t)p->ctg_off.size() - 1;
    std::vector<std::vector<int32_t>> add((size_t)n_ctg);
    for (auto &t : taken) add[(size_t)t.second].push_back(t.first);
    Rows m;
    std::vector<int64_t> off{0};
    auto copy_row = [&](const Rows &s, size_t i) {
        m.qs.push_back(s.qs[i]);
        m.qe.push_back(s.qe[i]);
        m.rs.push_back(s.rs[i]);
        m.re.push_back(s.re[i]);
        m.qtot.push_back(s.qtot[i]);
        m.rtot.push_back(s.rtot[i]);
        m.chr.push_back(s.chr[i]);
        m.mat_num.push_back(s.mat_num[i]);
        m.aln_len.push_back(s.aln_len[i]);
        m.orig_idx.push_back(s.orig_idx[i]);
        m.fwd.push_back(s.fwd[i]);
        m.mapq.push_back(s.mapq[i]);
        m.orig_alt.push_back(s.orig_alt[i]);
        m.cs.push_back(s.cs[i]);
        for (int64_t k = s.run_off[i]; k < s.run_off[i + 1]; k++) {
            m.run_ql.push_back(s.run_ql[(size_t)k]);
            m.run_qr.push_back(s.run_qr[(size_t)k]);
            m.run_rl.push_back(s.run_rl[(size_t)k]);
        }
        m.run_off.push_back((int64_t)m.run_ql.size());
    };
    for (int64_t c = 0; c < n_ctg; c++) {
        for (int64_t g = p->ctg_off[(size_t)c]; g < p->ctg_off[(size_t)c + 1]; g++) copy_row(p->r, (size_t)g);
        for (int32_t j : add[(size_t)c]) copy_row(alt, (size_t)j);
        off.push_back((int64_t)m.size());
    }
    p->r = std::move(m);
    p->ctg_off = std::move(off);
    p->bind();
    return AA_OK;
}

const aa_batch *aa_paf_batch(const aa_paf *paf) { return paf ? &paf->batch : nullptr; }

void aa_paf_free(aa_paf *paf) { delete paf; }

}  // extern "C"

namespace {

// get_edited_paf_data (paf_data.cpp:125-220)
bool edit_row(const aa_paf &p, int64_t g, int64_t eqs, int64_t eqe, int64_t ers, int64_t ere, std::string &cs_out,
              int32_t &mat, int32_t &aln, std::vector<CsOp> &ops, std::vector<CsOp> &kept, std::string &why) {
    if (eqs == p.r.qs[(size_t)g] && eqe == p.r.qe[(size_t)g]) {
        cs_out.assign(p.r.cs[(size_t)g]);
        mat = p.r.mat_num[(size_t)g];
        aln = p.r.aln_len[(size_t)g];
        return true;
    }
    const std::string_view cs = p.r.cs[(size_t)g];
    if (!parse_cs(cs, ops, why)) return false;
    bool fwd = p.r.fwd[(size_t)g] != 0;
    kept.clear();
    int64_t qi = p.r.qs[(size_t)g];
    size_t nop = ops.size();
    for (size_t k = 0; k < nop; k++) {
        const CsOp &o = fwd ? ops[k] : ops[nop - 1 - k];
        if (o.type == ':') {
            int64_t oe = qi + o.len - 1;
            int64_t a = qi > eqs ? qi : eqs, b = oe < eqe ? oe : eqe;
            if (a <= b) kept.push_back({':', b - a + 1, 0, 0});
            qi += o.len;
        } else if (o.type == '+') {
            int64_t oe = qi + o.len - 1;
            if (qi <= eqe && eqs <= oe) {
                if (qi < eqs || eqe < oe) {
                    why = "Alignment was clipped inside a cs insertion";
                    return false;
                }
                kept.push_back(o);
            }
            qi += o.len;
        } else if (o.type == '*') {
            if (eqs <= qi && qi <= eqe) kept.push_back(o);
            qi += 1;
        } else {
            if (eqs < qi && qi <= eqe) kept.push_back(o);
        }
    }
    cs_out = "cs:Z:";
    mat = 0;
    aln = 0;
    int64_t qb = 0, rb = 0;
    size_t nk = kept.size();
    for (size_t k = 0; k < nk; k++) {
        const CsOp &o = fwd ? kept[k] : kept[nk - 1 - k];
        if (o.type == ':') {
            cs_out += ':';
            cs_out += std::to_string(o.len);
            mat += (int32_t)o.len;
            aln += (int32_t)o.len;
            qb += o.len;
            rb += o.len;
        } else {
            cs_out.append(cs.substr(o.at, o.n));
            aln += (int32_t)o.len;
            if (o.type == '+') qb += o.len;
            else if (o.type == '-') rb += o.len;
            else {
                qb += 1;
                rb += 1;
            }
        }
    }
    int64_t want_r = ere > ers ? ere - ers : ers - ere;
    if (qb != eqe - eqs + 1 || rb != want_r + 1) {
        why = "Edited cs tag does not match edited PAF coordinates";
        return false;
    }
    return true;
}

}  // namespace

static aa_status paf_write_impl(const aa_paf *paf, aa_ctx *ctx, const aa_result *res, const char *out_prefix, char *err,
                                int64_t err_cap) {
    if (!paf || !res || !out_prefix) return AA_ERR_INVALID;
    const aa_paf &p = *paf;
    if (res->n_ctg != p.batch.n_ctg) {
        set_err(err, err_cap, "result does not belong to this PAF");
        return AA_ERR_INVALID;
    }
    // get_edited_paf_data of the primary and alternative rows on the device (the .aln.all.paf list, which can be gigabytes,
    // stays with the host codec below)
    aa_cs_edits ded{};
    struct EditsGuard {
        aa_cs_edits *e;
        ~EditsGuard() { aa_cs_edits_free(e); }
    } ded_guard{&ded};
    const int64_t n_out_rows = res->out.n, n_alt_rows = res->alt.n;
    if (ctx) {
        const size_t N = p.r.size();
        std::vector<int64_t> cs_off(N);
        std::vector<int32_t> cs_len(N);
        const char *base = nullptr;
        int64_t text_len = 0;
        std::string joined;  // more than one file image (--alt): the fields are laid out back to back
        if (p.images.size() == 1) {
            base = p.images[0]->data;
            text_len = (int64_t)p.images[0]->size;
            for (size_t i = 0; i < N; i++) cs_off[i] = (int64_t)(p.r.cs[i].data() - base);
        } else {
            size_t total = 0;
            for (size_t i = 0; i < N; i++) total += p.r.cs[i].size();
            joined.reserve(total);
            for (size_t i = 0; i < N; i++) {
                cs_off[i] = (int64_t)joined.size();
                joined.append(p.r.cs[i]);
            }
            base = joined.data();
            text_len = (int64_t)joined.size();
        }
        for (size_t i = 0; i < N; i++) cs_len[i] = (int32_t)p.r.cs[i].size();
        const int64_t M = n_out_rows + n_alt_rows;
        std::vector<int64_t> orow((size_t)M), eqs((size_t)M), eqe((size_t)M), ers((size_t)M), ere((size_t)M);
        for (int64_t c = 0; c < res->n_ctg; c++) {
            for (int64_t k = res->out_off[c]; k < res->out_off[c + 1]; k++) orow[(size_t)k] = p.ctg_off[(size_t)c] + res->out.ctg_index[k];
            for (int64_t k = res->alt_off[c]; k < res->alt_off[c + 1]; k++)
                orow[(size_t)(n_out_rows + k)] = p.ctg_off[(size_t)c] + res->alt.ctg_index[k];
        }
        for (int64_t k = 0; k < n_out_rows; k++) {
            eqs[(size_t)k] = res->out.qry_str[k];
            eqe[(size_t)k] = res->out.qry_end[k];
            ers[(size_t)k] = res->out.ref_str[k];
            ere[(size_t)k] = res->out.ref_end[k];
        }
        for (int64_t k = 0; k < n_alt_rows; k++) {
            eqs[(size_t)(n_out_rows + k)] = res->alt.qry_str[k];
            eqe[(size_t)(n_out_rows + k)] = res->alt.qry_end[k];
            ers[(size_t)(n_out_rows + k)] = res->alt.ref_str[k];
            ere[(size_t)(n_out_rows + k)] = res->alt.ref_end[k];
        }
        aa_cs_rows rows{(int64_t)N, cs_off.data(), cs_len.data(), p.r.qs.data(), p.r.qe.data(), p.r.rs.data(), p.r.re.data(), p.r.fwd.data()};
        aa_status dst = aa_cs_edit_device(ctx, base, text_len, &rows, p.r.mat_num.data(), p.r.aln_len.data(), M, orow.data(), eqs.data(),
                                          eqe.data(), ers.data(), ere.data(), &ded);
        if (dst != AA_OK) {
            set_err(err, err_cap, std::string("cs codec on the device: ") + aa_cs_last_error());
            return dst;
        }
    }
    std::string pre(out_prefix);
    FILE *fo[3] = {std::fopen((pre + ".aln.paf").c_str(), "wb"), std::fopen((pre + ".aln.alt.paf").c_str(), "wb"),
                   std::fopen((pre + ".aln.all.paf").c_str(), "wb")};
    if (!fo[0] || !fo[1] || !fo[2]) {
        for (FILE *f : fo)
            if (f) std::fclose(f);
        set_err(err, err_cap, "cannot open output files for " + pre);
        return AA_ERR_IO;
    }
    // Contigs are formatted by all host threads (dynamic, one contig at a time) and written in input-contig order
    // (alignasm.cpp:417-441, 456-482) AS SOON AS their turn comes: a worker that finishes contig c hands its text over, and
    // whoever completes the contig the files are waiting for writes it and everything behind it that is ready, then frees
    // the text.  Workers run at most WINDOW contigs ahead of the files, so the text in memory is bounded; a contig whose
    // .aln.all.paf list is large (every tied max-coverage walk: gigabytes for a big contig) is not buffered at all: its
    // worker waits for the contig's turn and streams the list to the file path by path, like the reference does.
    const int64_t C = res->n_ctg;
    const int T = host_threads((size_t)(res->out_off[C] + res->alt_off[C]) * 64);
    const int64_t WINDOW = std::max<int64_t>(4 * (int64_t)T, 16);
    int64_t ALL_STREAM_ROWS = 1 << 16;  // ~8 MB of text
    if (const char *e = std::getenv("AA_WRITE_STREAM_ROWS")) ALL_STREAM_ROWS = std::atoll(e);  // (tests force the streamed form)
    struct Slot {
        std::string text[3];
        bool done = false, bad = false;
        std::string why;
    };
    std::vector<Slot> slot((size_t)WINDOW);
    std::mutex mu;
    std::condition_variable cv;
    int64_t written = 0;  // next contig the files are waiting for
    aa_status st = AA_OK;
    bool stop = false;     // a bad row or a short write: rows up to there are on disk, like the reference's throw mid-write
    std::atomic<int64_t> next{0};
    auto write_text = [&](int k, const std::string &t) {  // under mu
        if (!t.empty() && !stop && std::fwrite(t.data(), 1, t.size(), fo[k]) != t.size()) {
            st = AA_ERR_IO;
            stop = true;
            set_err(err, err_cap, "short write to the output files of " + pre);
        }
    };
    run_parallel(T, [&](int) {
        std::vector<CsOp> ops, kept;
        std::string cs_out, why;
        char num[24];
        auto put_i = [&](std::string &dst, int64_t v) {
            auto r = std::to_chars(num, num + sizeof num, v);
            dst.append(num, (size_t)(r.ptr - num));
            dst.push_back('\t');
        };
        // one output row (alignasm.cpp:426-440 / 467-481)
        auto put = [&](std::string &dst, int64_t c, const std::string &qname, const aa_rows &rows, int64_t k) -> bool {
            const size_t g = (size_t)(p.ctg_off[(size_t)c] + rows.ctg_index[k]);
            int32_t mat, aln;
            const int64_t qs = rows.qry_str[k], qe = rows.qry_end[k], rs = rows.ref_str[k], re = rows.ref_end[k];
            if (ctx && &rows != &res->all) {  // edited on the device: row k of the primary list, n_out_rows + k of the alternative one
                const int64_t e = (&rows == &res->out) ? k : n_out_rows + k;
                if (ded.err[e]) {
                    why = aa_cs_error_text(ded.err[e]);
                    return false;
                }
                cs_out.assign(ded.text + ded.off[e], (size_t)(ded.off[e + 1] - ded.off[e]));
                mat = ded.mat_num[e];
                aln = ded.aln_len[e];
            } else if (!edit_row(p, (int64_t)g, qs, qe, rs, re, cs_out, mat, aln, ops, kept, why)) {
                return false;
            }
            const bool fwd = p.r.fwd[g] != 0;
            dst.append(qname);
            dst.push_back('\t');
            put_i(dst, p.r.qtot[g]);
            put_i(dst, qs);
            put_i(dst, qe + 1);
            dst.append(fwd ? "+\t" : "-\t");
            dst.append(p.chrs.names[(size_t)p.r.chr[g]]);
            dst.push_back('\t');
            put_i(dst, p.r.rtot[g]);
            put_i(dst, fwd ? rs : re);
            put_i(dst, (fwd ? re : rs) + 1);
            put_i(dst, mat);
            put_i(dst, aln);
            put_i(dst, (int64_t)p.r.mapq[g]);
            dst.append(rows.is_alt[k] ? "tp:A:S\txi:Z:" : "tp:A:P\txi:Z:");
            dst.append(p.r.orig_alt[g] ? "A_" : "P_");  // cord_to_index_string (alignasm.cpp:398-405)
            put_i(dst, (int64_t)p.r.orig_idx[g]);
            dst.append(cs_out);
            dst.push_back('\n');
            return true;
        };
        std::string t0, t1, t2;
        for (;;) {
            const int64_t c = next.fetch_add(1);
            if (c >= C) break;
            {
                std::unique_lock<std::mutex> lock(mu);
                cv.wait(lock, [&] { return stop || c < written + WINDOW; });
                if (stop) break;
            }
            const std::string &name = p.ctg_name[(size_t)c];
            t0.clear();
            t1.clear();
            t2.clear();
            bool ok = true;
            for (int64_t k = res->out_off[c]; k < res->out_off[c + 1] && ok; k++) ok = put(t0, c, name, res->out, k);
            for (int64_t k = res->alt_off[c]; k < res->alt_off[c + 1] && ok; k++) ok = put(t1, c, name, res->alt, k);
            const bool has_all = res->all_path_off && res->all_row_off;
            int64_t all_rows = 0;
            if (has_all) all_rows = res->all_row_off[res->all_path_off[c + 1]] - res->all_row_off[res->all_path_off[c]];
            const bool stream_all = has_all && all_rows > ALL_STREAM_ROWS;
            if (has_all && !stream_all) {
                int32_t cnt = 0;
                for (int64_t m = res->all_path_off[c]; m < res->all_path_off[c + 1] && ok; m++) {
                    const std::string qn = name + "." + std::to_string(++cnt);
                    for (int64_t k = res->all_row_off[m]; k < res->all_row_off[m + 1] && ok; k++) ok = put(t2, c, qn, res->all, k);
                }
            }
            std::unique_lock<std::mutex> lock(mu);
            if (stream_all) {  // wait for this contig's turn, then write straight to the files, one path at a time
                cv.wait(lock, [&] { return stop || written == c; });
                if (stop) break;
                write_text(0, t0);
                write_text(1, t1);
                int32_t cnt = 0;
                for (int64_t m = res->all_path_off[c]; m < res->all_path_off[c + 1] && ok && !stop; m++) {
                    const std::string qn = name + "." + std::to_string(++cnt);
                    t2.clear();
                    for (int64_t k = res->all_row_off[m]; k < res->all_row_off[m + 1] && ok; k++) ok = put(t2, c, qn, res->all, k);
                    write_text(2, t2);
                }
                if (!ok && st == AA_OK) {
                    st = AA_ERR_FORMAT;
                    stop = true;
                    set_err(err, err_cap, why);
                }
                written++;
            } else {
                Slot &s = slot[(size_t)(c % WINDOW)];
                s.text[0].swap(t0);
                s.text[1].swap(t1);
                s.text[2].swap(t2);
                s.bad = !ok;
                s.why = ok ? std::string() : why;
                s.done = true;
            }
            // write whatever is ready, in order
            while (!stop && written < C && slot[(size_t)(written % WINDOW)].done) {
                Slot &s = slot[(size_t)(written % WINDOW)];
                for (int k = 0; k < 3; k++) {
                    write_text(k, s.text[k]);
                    std::string().swap(s.text[k]);
                }
                if (s.bad && st == AA_OK) {
                    st = AA_ERR_FORMAT;
                    stop = true;
                    set_err(err, err_cap, s.why);
                }
                s.done = false;
                written++;
            }
            lock.unlock();
            cv.notify_all();
        }
        cv.notify_all();
    });
    for (FILE *f : fo)
        if (std::fclose(f) != 0 && st == AA_OK) {
            st = AA_ERR_IO;
            set_err(err, err_cap, "cannot finish the output files of " + pre);
        }
    return st;
}
extern "C" aa_status aa_paf_write(const aa_paf *paf, const aa_result *res, const char *out_prefix, char *err, int64_t err_cap) {
    return paf_write_impl(paf, nullptr, res, out_prefix, err, err_cap);
}
extern "C" aa_status aa_paf_write_device(const aa_paf *paf, aa_ctx *ctx, const aa_result *res, const char *out_prefix, char *err,
                                         int64_t err_cap) {
    if (!ctx) return AA_ERR_INVALID;
    return paf_write_impl(paf, ctx, res, out_prefix, err, err_cap);
}
