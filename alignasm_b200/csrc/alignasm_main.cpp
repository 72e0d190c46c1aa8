// alignasm — drop-in CLI: `alignasm <input.paf>` writes <input>.aln.paf, <input>.aln.alt.paf, <input>.aln.all.paf.
// Restates the orchestration of reference src/alignasm.cpp:28-72, 340-349, 487-490 on top of the C ABI;
// the per-contig solve loop (alignasm.cpp:346-397) is one aa_solve() call on the GPU.
#include "../../include/alignasm_b200.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static void usage(FILE *f) {
    std::fputs(
        "Usage: alignasm [--help] [--version] [--thread THREAD] [--alt PAF_ALT_LOC] [--alt_baseline ALT_BASELINE]\n"
        "                [--non_skip_linkable] [--device N] [--devices A,B,...] [--no_all] PAF_LOC\n\n"
        "Positional arguments:\n  PAF_LOC                        Location of PAF file [required]\n\n"
        "Optional arguments:\n"
        "  -h, --help                     shows help message and exits\n"
        "  -v, --version                  prints version information and exits\n"
        "  -t, --thread THREAD            Number of threads (accepted for compatibility; contigs run on the GPU) [default: 1]\n"
        "  -a, --alt PAF_ALT_LOC          Location of alternative PAF file\n"
        "  -b, --alt_baseline ALT_BASELINE  Baseline for coverage of alternative PAF file [default: 0.5]\n"
        "  --non_skip_linkable            no edge a -> b when a -> c -> b exists\n"
        "  --device N                     CUDA device ordinal [default: 0]\n"
        "  --devices A,B,...              shard the contigs over several CUDA devices (cost-balanced, merged in input order)\n"
        "  --no_all                       do not materialise <input>.aln.all.paf (written empty)\n"
        "  --cs_device                    parse and re-cut the cs:Z: tags on the GPU as well (same bytes; single device)\n",
        f);
}

int main(int argc, char **argv) {
    std::string paf_loc, alt_loc;
    int threads = 1, device = 0;
    double alt_baseline = 0.5;
    std::vector<int32_t> devices;
    bool nsl = false, want_all = true, cs_device = false;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto val = [&](const char *) -> const char * { return i + 1 < argc ? argv[++i] : nullptr; };
        if (a == "-h" || a == "--help") {
            usage(stdout);
            return 0;
        } else if (a == "-v" || a == "--version") {
            std::puts("0.1.0");
            return 0;
        } else if (a == "-t" || a == "--thread") {
            const char *v = val("-t");
            if (!v) { usage(stderr); return 1; }
            threads = std::atoi(v);
        } else if (a == "-a" || a == "--alt") {
            const char *v = val("-a");
            if (!v) { usage(stderr); return 1; }
            alt_loc = v;
        } else if (a == "-b" || a == "--alt_baseline") {
            const char *v = val("-b");
            char *e = nullptr;
            if (v) alt_baseline = std::strtod(v, &e);
            if (!v || e == v || *e) { usage(stderr); return 1; }
        } else if (a == "--non_skip_linkable") {
            nsl = true;
        } else if (a == "--device") {
            const char *v = val("--device");
            if (!v) { usage(stderr); return 1; }
            device = std::atoi(v);
        } else if (a == "--devices") {
            const char *v = val("--devices");
            if (!v) { usage(stderr); return 1; }
            for (const char *p = v; *p;) {
                devices.push_back((int32_t)std::strtol(p, const_cast<char **>(&p), 10));
                if (*p == ',') p++;
                else if (*p) { usage(stderr); return 1; }
            }
        } else if (a == "--no_all") {
            want_all = false;
        } else if (a == "--cs_device") {
            cs_device = true;
        } else if (!a.empty() && a[0] == '-') {
            usage(stderr);
            return 1;
        } else if (paf_loc.empty()) {
            paf_loc = a;
        } else {
            usage(stderr);
            return 1;
        }
    }
    if (paf_loc.empty()) {
        usage(stderr);
        return 1;
    }
    if (paf_loc.size() < 4 || paf_loc.compare(paf_loc.size() - 4, 4, ".paf") != 0) {  // alignasm.cpp:68-72
        std::fprintf(stderr, "Wrong PAF file : \"%s\"", paf_loc.c_str());
        usage(stderr);
        return 1;
    }
    if (!alt_loc.empty() && (alt_loc.size() < 4 || alt_loc.compare(alt_loc.size() - 4, 4, ".paf") != 0)) {  // alignasm.cpp:191-195
        std::fprintf(stderr, "Wrong PAF file : \"%s\"", alt_loc.c_str());
        usage(stderr);
        return 1;
    }
    char err[512] = {0};
    aa_paf *paf = nullptr;
    aa_ctx *cs_ctx = nullptr;  // --cs_device: the context that parses / re-cuts the cs:Z: tags (and solves, on one device)
    aa_status st;
    if (cs_device) {
        if (devices.size() == 1) device = devices[0];
        st = aa_create(&cs_ctx, device);
        if (st != AA_OK) {
            std::fprintf(stderr, "alignasm: %s (this build has no CPU path)\n", aa_last_error(nullptr));
            return 1;
        }
        st = aa_paf_read_device(paf_loc.c_str(), cs_ctx, &paf, err, sizeof err);
    } else {
        st = aa_paf_read(paf_loc.c_str(), &paf, err, sizeof err);
    }
    if (st != AA_OK) {
        std::fprintf(stderr, "%s\n", err);
        if (cs_ctx) aa_destroy(cs_ctx);
        return 1;
    }
    if (!alt_loc.empty()) {  // alignasm.cpp:186-332
        st = aa_paf_read_alt(paf, alt_loc.c_str(), alt_baseline, err, sizeof err);
        if (st != AA_OK) {
            std::fprintf(stderr, "%s\n", err);
            aa_paf_free(paf);
            return 1;
        }
    }
    std::puts("File read complete");
    const aa_batch *b = aa_paf_batch(paf);
    if (threads > 1) std::printf("Analyze PAF %lld data in parallel\n", (long long)b->n_ctg);
    std::fflush(stdout);
    aa_opts opt{};
    opt.non_skip_linkable = nsl;
    opt.want_all = want_all;
    aa_result res{};
    if (devices.size() > 1) {
        st = aa_solve_multi(devices.data(), (int32_t)devices.size(), b, &opt, &res);
        if (st != AA_OK) {
            std::fprintf(stderr, "alignasm: %s (this build has no CPU path)\n", aa_multi_last_error());
            aa_paf_free(paf);
            return 1;
        }
    } else {
        if (devices.size() == 1) device = devices[0];
        aa_ctx *ctx = cs_ctx;
        st = ctx ? AA_OK : aa_create(&ctx, device);
        if (st != AA_OK) {
            std::fprintf(stderr, "alignasm: %s (this build has no CPU path)\n", aa_last_error(nullptr));
            aa_paf_free(paf);
            return 1;
        }
        st = aa_solve(ctx, b, &opt, &res);
        if (st != AA_OK) {
            std::fprintf(stderr, "alignasm: %s\n", aa_last_error(ctx));
            aa_destroy(ctx);
            aa_paf_free(paf);
            return 1;
        }
        if (!cs_ctx) aa_destroy(ctx);
    }
    std::puts("Write output PAF file");
    std::string prefix = paf_loc.substr(0, paf_loc.size() - 4);
    st = cs_ctx ? aa_paf_write_device(paf, cs_ctx, &res, prefix.c_str(), err, sizeof err) : aa_paf_write(paf, &res, prefix.c_str(), err, sizeof err);
    if (st != AA_OK) std::fprintf(stderr, "alignasm: %s\n", err);
    if (cs_ctx) aa_destroy(cs_ctx);
    aa_result_free(&res);
    aa_paf_free(paf);
    return st == AA_OK ? 0 : 1;
}
