// synth_paf — deterministic synthetic PAF generator for the alignasm hot-path benchmarks.
//
// Produces PAF rows grouped by contig, each with a short-form cs:Z: tag that is consistent
// with its coordinates (the reference rejects anything else: /root/reference/src/paf_data.cpp:119-122).
// Shapes follow SURVEY.md §8(d): a contig is a left-to-right walk over the query; blocks may
// partially overlap, be contained, jump chromosome (translocation), flip strand (inversion) or be
// duplicated at another locus (exact ties).  Targets are the 24 GRCh38 primary sequences.
//
// usage: synth_paf --preset c1|c2|c3|c4|c5 [--seed S] [--scale F] [--n N] [-o out.paf]
//        synth_paf --contigs N --blocks MEAN [--sd SD] [--lmin A --lmax B | --lmed M] ...
#include <cinttypes>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>

namespace {

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 0x1234567ull) {}
    uint64_t next() {  // splitmix64
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double uni() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
    // inclusive integer range
    int64_t range(int64_t lo, int64_t hi) {
        if (hi <= lo) return lo;
        return lo + (int64_t)(next() % (uint64_t)(hi - lo + 1));
    }
    bool chance(double p) { return uni() < p; }
    double normal() {
        double u1 = uni(), u2 = uni();
        if (u1 < 1e-300) u1 = 1e-300;
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
    }
};

const char *CHR_NAME[24] = {"chr1", "chr2", "chr3", "chr4", "chr5", "chr6", "chr7", "chr8",
                            "chr9", "chr10", "chr11", "chr12", "chr13", "chr14", "chr15", "chr16",
                            "chr17", "chr18", "chr19", "chr20", "chr21", "chr22", "chrX", "chrY"};
const int64_t CHR_LEN[24] = {248956422, 242193529, 198295559, 190214555, 181538259, 170805979,
                             159345973, 145138636, 138394717, 133797422, 135086622, 133275309,
                             114364328, 107043718, 101991189, 90338345,  83257441,  80373285,
                             58617616,  64444167,  46709983,  50818468,  156040895, 57227415};

struct Params {
    int64_t contigs = 1000;
    double blocks_mean = 50, blocks_sd = 15;
    int64_t blocks_min = 2;
    // block length: uniform [lmin,lmax] unless lmed>0 (log-normal around lmed)
    int64_t lmin = 20000, lmax = 200000;
    double lmed = 0, lsigma = 0.8;
    double p_trans = 0.05, p_inv = 0.05, p_ovl = 0.4, p_cont = 0.05, p_dup = 0.0;
    int64_t gap_max = 30000;
    // genome-shaped contig sizes (c2/c3/c5): contig lengths log-normal, blocks proportional
    bool genome = false;
    int64_t total_blocks = 500000;
    double genome_bp = 3.0e9;
    double ctg_sigma = 1.3;
    // dense single-contig ladder (c4)
    bool dense = false;
    int64_t dense_n = 845;
    std::string name_suffix;
    uint64_t seed = 1;
};

struct Op {
    char type;  // ':', '*', '+', '-'
    int64_t len;
};

// Build a cs op list (query orientation) consuming exactly qlen query bases.
void make_ops(Rng &rng, int64_t qlen, std::vector<Op> &ops, int64_t &ref_len) {
    ops.clear();
    int64_t q = 0, r = 0;
    for (;;) {
        int64_t m = rng.range(50, 5000);
        if (q + m >= qlen || qlen - (q + m) < 8) m = qlen - q;
        ops.push_back({':', m});
        q += m;
        r += m;
        if (q >= qlen) break;
        double u = rng.uni();
        if (u < 0.5) {
            ops.push_back({'*', 1});
            q += 1;
            r += 1;
        } else if (u < 0.75) {
            int64_t k = rng.range(1, 5);
            ops.push_back({'+', k});
            q += k;
        } else {
            int64_t k = rng.range(1, 5);
            ops.push_back({'-', k});
            r += k;
        }
        // the trailing piece is always a match run of >=1 base: q < qlen is guaranteed by the
        // ">= 8 left" rule above (a separator consumes at most 5 query bases)
    }
    ref_len = r;
}

struct Block {
    int64_t qs, qe;  // [qs, qe)
    int chr;
    int64_t ts, te;  // [ts, te)
    bool fwd;
    int mapq;
    std::vector<Op> ops;  // query orientation
};

void emit(FILE *out, const std::string &qname, int64_t qtotal, const Block &b) {
    int64_t nmatch = 0, alnlen = 0;
    for (const auto &o : b.ops) {
        if (o.type == ':') nmatch += o.len;
        alnlen += o.len;
    }
    std::fprintf(out, "%s\t%" PRId64 "\t%" PRId64 "\t%" PRId64 "\t%c\t%s\t%" PRId64 "\t%" PRId64 "\t%" PRId64
                      "\t%" PRId64 "\t%" PRId64 "\t%d\ttp:A:P\tcs:Z:",
                 qname.c_str(), qtotal, b.qs, b.qe, b.fwd ? '+' : '-', CHR_NAME[b.chr], CHR_LEN[b.chr], b.ts,
                 b.te, nmatch, alnlen, b.mapq);
    auto put = [&](const Op &o) {
        if (o.type == ':') {
            std::fprintf(out, ":%" PRId64, o.len);
        } else if (o.type == '*') {
            std::fputs("*ag", out);
        } else {
            std::fputc(o.type, out);
            for (int64_t k = 0; k < o.len; k++) std::fputc(o.type == '+' ? 'a' : 't', out);
        }
    };
    if (b.fwd) {
        for (const auto &o : b.ops) put(o);
    } else {
        for (auto it = b.ops.rbegin(); it != b.ops.rend(); ++it) put(*it);
    }
    std::fputc('\n', out);
}

const int MAPQ_CHOICES[7] = {0, 0, 1, 30, 60, 60, 60};

int64_t draw_len(Rng &rng, const Params &p) {
    if (p.lmed > 0) {
        double v = p.lmed * std::exp(p.lsigma * rng.normal());
        if (v < 300) v = 300;
        if (v > 2.0e6) v = 2.0e6;
        return (int64_t)v;
    }
    return rng.range(p.lmin, p.lmax);
}

// place [ref_len] on (chr, strand) near `anchor` (query-facing start), clamped into the chromosome
void place(Rng &rng, Block &b, int64_t ref_len, int chr, bool fwd, int64_t anchor) {
    b.chr = chr;
    b.fwd = fwd;
    int64_t tl = CHR_LEN[chr];
    if (ref_len >= tl) ref_len = tl - 1;  // never happens with our block sizes
    int64_t ts;
    if (fwd) ts = anchor;            // query-facing start is ts
    else ts = anchor - ref_len;      // query-facing start is te
    if (ts < 0) ts = 0;
    if (ts + ref_len > tl) ts = tl - ref_len;
    (void)rng;
    b.ts = ts;
    b.te = ts + ref_len;
}

void gen_contig(Rng &rng, const Params &p, FILE *out, const std::string &name, int64_t nblocks,
                int64_t *rows) {
    std::vector<Block> blocks;
    blocks.reserve((size_t)nblocks + 8);
    int chr = (int)rng.range(0, 23);
    bool fwd = rng.chance(0.5);
    int64_t tpos = rng.range(0, CHR_LEN[chr] / 2);  // query-facing position on the target
    int64_t qpos = rng.range(0, 5000);
    int64_t prev_qs = -1, prev_qe = -1;
    for (int64_t k = 0; k < nblocks; k++) {
        Block b;
        int64_t L = draw_len(rng, p);
        int64_t qs, qe;
        if (p.dense) {
            int64_t L0 = (p.lmin + p.lmax) / 2;
            qs = (prev_qs < 0) ? qpos : prev_qs + rng.range(L0 / 50, L0 / 10);
            qe = qs + L;
        } else if (prev_qs >= 0 && rng.chance(p.p_ovl)) {
            int64_t plen = prev_qe - prev_qs;
            int64_t ovl = rng.range(100, std::max<int64_t>(100, plen / 3));
            if (ovl >= plen) ovl = plen / 2;
            qs = prev_qe - ovl;
            if (qs <= prev_qs) qs = prev_qs + 1;
            if (L <= prev_qe - qs + 50) L = prev_qe - qs + 50 + rng.range(0, 1000);
            qe = qs + L;
        } else if (prev_qs >= 0 && prev_qe - prev_qs > 1200 && rng.chance(p.p_cont)) {
            int64_t plen = prev_qe - prev_qs;
            int64_t cl = rng.range(300, plen - 200);
            qs = prev_qs + rng.range(1, plen - cl - 1);
            qe = qs + cl;
        } else {
            qs = (prev_qs < 0) ? qpos : prev_qe + rng.range(0, p.gap_max);
            qe = qs + L;
        }
        b.qs = qs;
        b.qe = qe;
        int64_t ref_len;
        make_ops(rng, qe - qs, b.ops, ref_len);
        // target placement
        if (k > 0 && rng.chance(p.p_trans)) {
            chr = (int)rng.range(0, 23);
            tpos = rng.range(0, CHR_LEN[chr] - 1);
            fwd = rng.chance(0.5);
        } else if (k > 0 && rng.chance(p.p_inv)) {
            fwd = !fwd;
        } else if (k > 0) {
            // colinear continuation: advance the target by the query delta plus a small indel
            int64_t dq = qs - prev_qs + rng.range(-50, 50);
            tpos += fwd ? dq : -dq;
        }
        place(rng, b, ref_len, chr, fwd, tpos);
        tpos = fwd ? b.ts : b.te;  // query-facing start of this block
        b.mapq = MAPQ_CHOICES[rng.range(0, 6)];
        blocks.push_back(b);
        // contained blocks do not move the walk forward
        if (qe > prev_qe) {
            prev_qs = qs;
            prev_qe = qe;
        }
        if (p.p_dup > 0 && rng.chance(p.p_dup)) {
            Block d = blocks.back();
            int dchr = (int)rng.range(0, 23);
            bool dfwd = rng.chance(0.5);
            int64_t rl = d.te - d.ts;
            place(rng, d, rl, dchr, dfwd, rng.range(0, CHR_LEN[dchr] - 1));
            d.mapq = MAPQ_CHOICES[rng.range(0, 6)];
            blocks.push_back(d);
        }
    }
    int64_t qtotal = 0;
    for (const auto &b : blocks) qtotal = std::max(qtotal, b.qe);
    qtotal += rng.range(0, 5000);
    for (const auto &b : blocks) emit(out, name, qtotal, b);
    *rows += (int64_t)blocks.size();
}

void preset(Params &p, const std::string &name) {
    if (name == "c1") {
        p.contigs = 1000; p.blocks_mean = 50; p.blocks_sd = 15; p.lmin = 20000; p.lmax = 200000;
        p.p_trans = p.p_inv = 0.05; p.p_ovl = 0.4; p.p_cont = 0.05; p.p_dup = 0; p.seed = 1;
    } else if (name == "c2") {
        p.genome = true; p.total_blocks = 500000; p.lmed = 4000; p.p_trans = p.p_inv = 0.01;
        p.p_ovl = 0.4; p.p_cont = 0.05; p.p_dup = 0; p.seed = 2; p.gap_max = 3000;
    } else if (name == "c3") {
        p.genome = true; p.total_blocks = 500000; p.lmed = 4000; p.p_trans = p.p_inv = 0.15;
        p.p_ovl = 0.4; p.p_cont = 0.05; p.p_dup = 0.10; p.seed = 3; p.gap_max = 3000;
    } else if (name == "c4") {
        p.dense = true; p.contigs = 1; p.lmin = 20000; p.lmax = 200000; p.p_trans = p.p_inv = 0.05;
        p.p_dup = 0; p.seed = 4;
    } else if (name == "c5") {
        // one replica of the box-scaling input; the caller concatenates 8 of them (seeds 30..37)
        p.genome = true; p.total_blocks = 500000; p.lmed = 4000; p.p_trans = p.p_inv = 0.15;
        p.p_ovl = 0.4; p.p_cont = 0.05; p.p_dup = 0.10; p.seed = 30; p.gap_max = 3000;
    } else {
        std::fprintf(stderr, "unknown preset %s\n", name.c_str());
        std::exit(2);
    }
}

}  // namespace

int main(int argc, char **argv) {
    Params p;
    std::string out_path;
    double scale = 1.0;
    int replicas = 1;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto need = [&](const char *what) -> const char * {
            if (i + 1 >= argc) {
                std::fprintf(stderr, "missing value for %s\n", what);
                std::exit(2);
            }
            return argv[++i];
        };
        if (a == "--preset") preset(p, need("--preset"));
        else if (a == "--seed") p.seed = std::strtoull(need("--seed"), nullptr, 10);
        else if (a == "--scale") scale = std::atof(need("--scale"));
        else if (a == "--contigs") p.contigs = std::atoll(need("--contigs"));
        else if (a == "--blocks") p.blocks_mean = std::atof(need("--blocks"));
        else if (a == "--sd") p.blocks_sd = std::atof(need("--sd"));
        else if (a == "--lmin") p.lmin = std::atoll(need("--lmin"));
        else if (a == "--lmax") p.lmax = std::atoll(need("--lmax"));
        else if (a == "--lmed") p.lmed = std::atof(need("--lmed"));
        else if (a == "--p_trans") p.p_trans = std::atof(need("--p_trans"));
        else if (a == "--p_inv") p.p_inv = std::atof(need("--p_inv"));
        else if (a == "--p_ovl") p.p_ovl = std::atof(need("--p_ovl"));
        else if (a == "--p_cont") p.p_cont = std::atof(need("--p_cont"));
        else if (a == "--p_dup") p.p_dup = std::atof(need("--p_dup"));
        else if (a == "--gap_max") p.gap_max = std::atoll(need("--gap_max"));
        else if (a == "--genome") p.genome = true;
        else if (a == "--total_blocks") p.total_blocks = std::atoll(need("--total_blocks"));
        else if (a == "--dense") p.dense = true;
        else if (a == "--n") p.dense_n = std::atoll(need("--n"));
        else if (a == "--replicas") replicas = std::atoi(need("--replicas"));
        else if (a == "-o") out_path = need("-o");
        else {
            std::fprintf(stderr, "unknown argument %s\n", a.c_str());
            return 2;
        }
    }
    FILE *out = out_path.empty() ? stdout : std::fopen(out_path.c_str(), "w");
    if (!out) {
        std::perror("open output");
        return 1;
    }
    static char buf[1 << 20];
    std::setvbuf(out, buf, _IOFBF, sizeof buf);

    int64_t rows = 0, contigs = 0;
    for (int rep = 0; rep < replicas; rep++) {
        Params q = p;
        q.seed = p.seed + (uint64_t)rep;
        Rng rng(q.seed);
        std::string suffix = replicas > 1 ? "_r" + std::to_string(rep) : "";
        if (q.dense) {
            gen_contig(rng, q, out, "ctg_dense" + suffix, q.dense_n, &rows);
            contigs++;
        } else if (q.genome) {
            // contig lengths: log-normal, rescaled to genome_bp; blocks proportional to length
            int64_t nctg = std::max<int64_t>(2, (int64_t)(260 * std::min(1.0, scale * 4)));
            std::vector<double> len((size_t)nctg);
            double tot = 0;
            for (auto &l : len) {
                l = std::exp(q.ctg_sigma * rng.normal());
                tot += l;
            }
            int64_t want = (int64_t)(q.total_blocks * scale);
            for (int64_t c = 0; c < nctg; c++) {
                int64_t nb = std::max<int64_t>(2, (int64_t)std::llround(want * (len[(size_t)c] / tot)));
                gen_contig(rng, q, out, "ctg" + std::to_string(c) + suffix, nb, &rows);
                contigs++;
            }
        } else {
            int64_t nctg = std::max<int64_t>(1, (int64_t)std::llround(q.contigs * scale));
            for (int64_t c = 0; c < nctg; c++) {
                int64_t nb = (int64_t)std::llround(q.blocks_mean + q.blocks_sd * rng.normal());
                if (nb < q.blocks_min) nb = q.blocks_min;
                gen_contig(rng, q, out, "ctg" + std::to_string(c) + suffix, nb, &rows);
                contigs++;
            }
        }
    }
    if (out != stdout) std::fclose(out);
    std::fprintf(stderr, "synth_paf: %" PRId64 " contigs, %" PRId64 " rows\n", contigs, rows);
    return 0;
}
