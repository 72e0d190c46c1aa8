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
#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

// one table of parsed rows, structure-of-arrays (the main PAF, and the rows of the alternative PAF before the merge)
struct Rows {
    std::vector<int64_t> qs, qe, rs, re, qtot, rtot, run_off{0}, run_ql, run_qr, run_rl;
    std::vector<int32_t> chr, mat_num, aln_len, orig_idx;
    std::vector<uint8_t> fwd, mapq, orig_alt;  // orig_alt: TYPE_ALT row (xi:Z:A_<row>), else TYPE_MAIN (xi:Z:P_<row>)
    std::vector<std::string> cs;               // original cs:Z: field per row (needed to re-cut for output)
    size_t size() const { return qs.size(); }
};

struct aa_paf {
    aa_batch batch{};
    Rows r;
    std::vector<int64_t> ctg_off;
    std::vector<std::string> ctg_name, chr_name;
    std::unordered_map<std::string, int32_t> chr_map;
    std::unordered_map<std::string, int32_t> paf_map;  // query name -> last bucket of that name (alignasm.cpp:136)
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
aa_status parse_row(const std::vector<std::string_view> &f, int64_t row, int64_t q_off, const char *what, aa_paf &p, Rows &t,
                    std::vector<CsOp> &ops, std::string &why) {
    if (f.size() < 12) {
        why = std::string(what) + " row " + std::to_string(row) + " has fewer than 12 columns";
        return AA_ERR_FORMAT;
    }
    std::string ref_chr(f[5]);
    auto it = p.chr_map.find(ref_chr);
    int32_t chr_id;
    if (it == p.chr_map.end()) {
        chr_id = (int32_t)p.chr_name.size();
        p.chr_map.emplace(ref_chr, chr_id);
        p.chr_name.push_back(ref_chr);
    } else {
        chr_id = it->second;
    }
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
    if (!parse_cs(cs, ops, why)) return AA_ERR_FORMAT;
    int64_t step = fwd ? 1 : -1, ri = rs, qi = qs;
    size_t nop = ops.size();
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
    if (qi != qe + 1 || ri != re + step) {
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
    t.cs.emplace_back(cs);
    return AA_OK;
}

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

aa_status aa_paf_read(const char *path, aa_paf **out, char *err, int64_t err_cap) {
    if (!path || !out) return AA_ERR_INVALID;
    *out = nullptr;
    FILE *fp = std::fopen(path, "rb");
    if (!fp) {
        set_err(err, err_cap, std::string("cannot open ") + path);
        return AA_ERR_IO;
    }
    aa_paf *p = new aa_paf();
    std::string ctg_chr;
    std::vector<std::string_view> f;
    std::vector<CsOp> ops;
    std::string why;
    char *line = nullptr;
    size_t cap = 0;
    ssize_t got;
    int64_t row = 0;
    aa_status st = AA_OK;
    while ((got = getline(&line, &cap, fp)) >= 0) {
        while (got > 0 && (line[got - 1] == '\n' || line[got - 1] == '\r')) got--;
        if (got == 0) continue;
        std::string_view lv(line, (size_t)got);
        split_tabs(lv, f);
        st = parse_row(f, row, 0, "PAF", *p, p->r, ops, why);
        if (st != AA_OK) {
            set_err(err, err_cap, why);
            break;
        }
        // bucket by *change* of the query name (alignasm.cpp:115-133): a name that re-appears later
        // opens a new contig, exactly as the reference does
        if (p->ctg_off.empty() || ctg_chr != f[0]) {
            p->ctg_off.push_back(row);
            p->ctg_name.emplace_back(f[0]);
            ctg_chr.assign(f[0]);
        }
        p->paf_map[ctg_chr] = (int32_t)p->ctg_off.size() - 1;
        p->r.orig_idx.push_back((int32_t)row);  // original_cord = {TYPE_MAIN, row_global_index} (alignasm.cpp:172)
        p->r.orig_alt.push_back(0);
        row++;
    }
    std::free(line);
    std::fclose(fp);
    if (st == AA_OK && row == 0) {
        set_err(err, err_cap, "PAF file holds no rows");
        st = AA_ERR_FORMAT;
    }
    if (st != AA_OK) {
        delete p;
        return st;
    }
    p->ctg_off.push_back(row);
    p->bind();
    *out = p;
    return AA_OK;
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
    FILE *fp = std::fopen(alt_path, "rb");
    if (!fp) {
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
    char *line = nullptr;
    size_t cap = 0;
    ssize_t got;
    int64_t row = 0;
    aa_status st = AA_OK;
    while ((got = getline(&line, &cap, fp)) >= 0) {
        while (got > 0 && (line[got - 1] == '\n' || line[got - 1] == '\r')) got--;
        if (got == 0) continue;
        std::string_view lv(line, (size_t)got);
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
        st = parse_row(f, row, q_off, "alternative PAF", *p, alt, ops, why);
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
    std::free(line);
    std::fclose(fp);
    if (st != AA_OK) return st;
    flush();
    if (taken.empty()) return AA_OK;
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
        cs_out = p.r.cs[(size_t)g];
        mat = p.r.mat_num[(size_t)g];
        aln = p.r.aln_len[(size_t)g];
        return true;
    }
    const std::string &cs = p.r.cs[(size_t)g];
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
            cs_out.append(cs, o.at, o.n);
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

extern "C" aa_status aa_paf_write(const aa_paf *paf, const aa_result *res, const char *out_prefix, char *err,
                                  int64_t err_cap) {
    if (!paf || !res || !out_prefix) return AA_ERR_INVALID;
    const aa_paf &p = *paf;
    if (res->n_ctg != p.batch.n_ctg) {
        set_err(err, err_cap, "result does not belong to this PAF");
        return AA_ERR_INVALID;
    }
    std::string pre(out_prefix);
    FILE *f1 = std::fopen((pre + ".aln.paf").c_str(), "wb");
    FILE *f2 = std::fopen((pre + ".aln.alt.paf").c_str(), "wb");
    FILE *f3 = std::fopen((pre + ".aln.all.paf").c_str(), "wb");
    if (!f1 || !f2 || !f3) {
        if (f1) std::fclose(f1);
        if (f2) std::fclose(f2);
        if (f3) std::fclose(f3);
        set_err(err, err_cap, "cannot open output files for " + pre);
        return AA_ERR_IO;
    }
    std::vector<CsOp> ops, kept;
    std::string cs_out, why;
    aa_status st = AA_OK;
    // one output row (alignasm.cpp:426-440 / 467-481)
    auto put = [&](FILE *fp, int64_t c, const std::string &qname, const aa_rows &rows, int64_t k) -> bool {
        int64_t g = p.ctg_off[(size_t)c] + rows.ctg_index[k];
        int32_t mat, aln;
        int64_t qs = rows.qry_str[k], qe = rows.qry_end[k], rs = rows.ref_str[k], re = rows.ref_end[k];
        if (!edit_row(p, g, qs, qe, rs, re, cs_out, mat, aln, ops, kept, why)) return false;
        bool fwd = p.r.fwd[(size_t)g] != 0;
        std::fprintf(fp,
                     "%s\t%" PRId64 "\t%" PRId64 "\t%" PRId64 "\t%s\t%s\t%" PRId64 "\t%" PRId64 "\t%" PRId64
                     "\t%d\t%d\t%d\t%s\txi:Z:%s%d\t%s\n",
                     qname.c_str(), p.r.qtot[(size_t)g], qs, qe + 1, fwd ? "+" : "-", p.chr_name[(size_t)p.r.chr[(size_t)g]].c_str(),
                     p.r.rtot[(size_t)g], fwd ? rs : re, (fwd ? re : rs) + 1, mat, aln, (int)p.r.mapq[(size_t)g],
                     rows.is_alt[k] ? "tp:A:S" : "tp:A:P", p.r.orig_alt[(size_t)g] ? "A_" : "P_", (int)p.r.orig_idx[(size_t)g],
                     cs_out.c_str());
        return true;
    };
    for (int64_t c = 0; c < res->n_ctg && st == AA_OK; c++) {
        const std::string &name = p.ctg_name[(size_t)c];
        for (int64_t k = res->out_off[c]; k < res->out_off[c + 1]; k++)
            if (!put(f1, c, name, res->out, k)) st = AA_ERR_FORMAT;
        for (int64_t k = res->alt_off[c]; k < res->alt_off[c + 1]; k++)
            if (!put(f2, c, name, res->alt, k)) st = AA_ERR_FORMAT;
        if (res->all_path_off && res->all_row_off) {
            int32_t cnt = 0;
            for (int64_t m = res->all_path_off[c]; m < res->all_path_off[c + 1]; m++) {
                std::string qn = name + "." + std::to_string(++cnt);
                for (int64_t k = res->all_row_off[m]; k < res->all_row_off[m + 1]; k++)
                    if (!put(f3, c, qn, res->all, k)) st = AA_ERR_FORMAT;
            }
        }
    }
    std::fclose(f1);
    std::fclose(f2);
    std::fclose(f3);
    if (st != AA_OK) set_err(err, err_cap, why);
    return st;
}
