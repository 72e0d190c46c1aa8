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

struct aa_paf {
    aa_batch batch{};
    std::vector<int64_t> ctg_off, qs, qe, rs, re, qtot, rtot, run_off, run_ql, run_qr, run_rl;
    std::vector<int32_t> chr, mat_num, aln_len;
    std::vector<uint8_t> fwd, mapq;
    std::vector<std::string> ctg_name, chr_name;
    std::vector<std::string> cs;  // original cs:Z: field per row (needed to re-cut for output)
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
    std::unordered_map<std::string, int32_t> chr_map;
    std::string ctg_chr;
    std::vector<std::string_view> f;
    std::vector<CsOp> ops;
    std::string why;
    char *line = nullptr;
    size_t cap = 0;
    ssize_t got;
    int64_t row = 0;
    aa_status st = AA_OK;
    p->run_off.push_back(0);
    while ((got = getline(&line, &cap, fp)) >= 0) {
        while (got > 0 && (line[got - 1] == '\n' || line[got - 1] == '\r')) got--;
        if (got == 0) continue;
        std::string_view lv(line, (size_t)got);
        split_tabs(lv, f);
        if (f.size() < 12) {
            set_err(err, err_cap, "PAF row " + std::to_string(row) + " has fewer than 12 columns");
            st = AA_ERR_FORMAT;
            break;
        }
        // bucket by *change* of the query name (alignasm.cpp:115-133): a name that re-appears later
        // opens a new contig, exactly as the reference does
        if (p->ctg_off.empty() || ctg_chr != f[0]) {
            p->ctg_off.push_back(row);
            p->ctg_name.emplace_back(f[0]);
            ctg_chr.assign(f[0]);
        }
        std::string ref_chr(f[5]);
        auto it = chr_map.find(ref_chr);
        int32_t chr_id;
        if (it == chr_map.end()) {
            chr_id = (int32_t)p->chr_name.size();
            chr_map.emplace(ref_chr, chr_id);
            p->chr_name.push_back(ref_chr);
        } else {
            chr_id = it->second;
        }
        int64_t qtot, qs, qe, rtot, rs, re, mat, aln, mq;
        if (!to_i64(f[1], qtot) || !to_i64(f[2], qs) || !to_i64(f[3], qe) || !to_i64(f[6], rtot) ||
            !to_i64(f[7], rs) || !to_i64(f[8], re) || !to_i64(f[9], mat) || !to_i64(f[10], aln) ||
            !to_i64(f[11], mq) || f[4].empty()) {
            set_err(err, err_cap, "PAF row " + std::to_string(row) + ": non-numeric coordinate field");
            st = AA_ERR_FORMAT;
            break;
        }
        qe -= 1;  // closed intervals (alignasm.cpp:143-150)
        re -= 1;
        bool fwd = f[4][0] == '+';
        if (!fwd) std::swap(rs, re);  // alignasm.cpp:155-159
        std::string_view cs;
        for (size_t k = 12; k < f.size(); k++)
            if (f[k].size() >= 5 && f[k].substr(0, 5) == "cs:Z:") {
                cs = f[k];
                break;
            }
        if (cs.empty()) {  // alignasm.cpp:165-168
            set_err(err, err_cap, "Missing cs:Z tag in PAF record for query '" + std::string(f[0]) + "'");
            st = AA_ERR_FORMAT;
            break;
        }
        // get_overlap_range (paf_data.cpp:90-123): walk the ops in query orientation
        if (!parse_cs(cs, ops, why)) {
            set_err(err, err_cap, why);
            st = AA_ERR_FORMAT;
            break;
        }
        int64_t step = fwd ? 1 : -1, ri = rs, qi = qs;
        size_t nop = ops.size();
        for (size_t k = 0; k < nop; k++) {
            const CsOp &o = fwd ? ops[k] : ops[nop - 1 - k];
            if (o.type == ':') {
                p->run_ql.push_back(qi);
                p->run_qr.push_back(qi + o.len - 1);
                p->run_rl.push_back(ri);
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
            set_err(err, err_cap, "cs tag consumption does not match PAF coordinates (row " + std::to_string(row) + ")");
            st = AA_ERR_FORMAT;
            break;
        }
        p->run_off.push_back((int64_t)p->run_ql.size());
        p->qs.push_back(qs);
        p->qe.push_back(qe);
        p->rs.push_back(rs);
        p->re.push_back(re);
        p->qtot.push_back(qtot);
        p->rtot.push_back(rtot);
        p->chr.push_back(chr_id);
        p->fwd.push_back(fwd ? 1 : 0);
        p->mapq.push_back((uint8_t)mq);
        p->mat_num.push_back((int32_t)mat);
        p->aln_len.push_back((int32_t)aln);
        p->cs.emplace_back(cs);
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
    aa_batch &b = p->batch;
    b.n_ctg = (int64_t)p->ctg_off.size() - 1;
    b.n_blk = row;
    b.n_run = (int64_t)p->run_ql.size();
    b.ctg_off = p->ctg_off.data();
    b.qry_str = p->qs.data();
    b.qry_end = p->qe.data();
    b.ref_str = p->rs.data();
    b.ref_end = p->re.data();
    b.qry_total = p->qtot.data();
    b.ref_chr = p->chr.data();
    b.aln_fwd = p->fwd.data();
    b.map_qul = p->mapq.data();
    b.run_off = p->run_off.data();
    b.run_ql = p->run_ql.data();
    b.run_qr = p->run_qr.data();
    b.run_rl = p->run_rl.data();
    *out = p;
    return AA_OK;
}

const aa_batch *aa_paf_batch(const aa_paf *paf) { return paf ? &paf->batch : nullptr; }

void aa_paf_free(aa_paf *paf) { delete paf; }

}  // extern "C"

namespace {

// get_edited_paf_data (paf_data.cpp:125-220)
bool edit_row(const aa_paf &p, int64_t g, int64_t eqs, int64_t eqe, int64_t ers, int64_t ere, std::string &cs_out,
              int32_t &mat, int32_t &aln, std::vector<CsOp> &ops, std::vector<CsOp> &kept, std::string &why) {
    if (eqs == p.qs[(size_t)g] && eqe == p.qe[(size_t)g]) {
        cs_out = p.cs[(size_t)g];
        mat = p.mat_num[(size_t)g];
        aln = p.aln_len[(size_t)g];
        return true;
    }
    const std::string &cs = p.cs[(size_t)g];
    if (!parse_cs(cs, ops, why)) return false;
    bool fwd = p.fwd[(size_t)g] != 0;
    kept.clear();
    int64_t qi = p.qs[(size_t)g];
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
        bool fwd = p.fwd[(size_t)g] != 0;
        std::fprintf(fp,
                     "%s\t%" PRId64 "\t%" PRId64 "\t%" PRId64 "\t%s\t%s\t%" PRId64 "\t%" PRId64 "\t%" PRId64
                     "\t%d\t%d\t%d\t%s\txi:Z:P_%" PRId64 "\t%s\n",
                     qname.c_str(), p.qtot[(size_t)g], qs, qe + 1, fwd ? "+" : "-", p.chr_name[(size_t)p.chr[(size_t)g]].c_str(),
                     p.rtot[(size_t)g], fwd ? rs : re, (fwd ? re : rs) + 1, mat, aln, (int)p.mapq[(size_t)g],
                     rows.is_alt[k] ? "tp:A:S" : "tp:A:P", g, cs_out.c_str());
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
