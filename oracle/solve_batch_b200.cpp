// oracle/solve_batch_b200.cpp — TEST INFRASTRUCTURE: the reference-side binding of INTEGRATION.md §2, verbatim, compiled.
// This is the file a reference maintainer would add as src/solve_batch_b200.cpp: it packs the reference's own
// std::vector<std::vector<PafReadData>> into one aa_batch, calls aa_solve, and unpacks into the three containers the
// reference's writers consume.  oracle/Makefile links it with the reference's paf_data.cpp (unmodified, for
// get_edited_paf_data) and libalignasm_b200.so into oracle/_ref/alignasm_ref_b200; tests/test_gpu_parity.py runs that
// binary on the GPU box and compares its three output files with the goldens byte for byte.  Keep in sync with the
// snippet in INTEGRATION.md (tests/test_abi.py checks that the two texts are identical).
// src/solve_batch_b200.cpp  — link with -lalignasm_b200
#include "paf_data.hpp"
#include "alignasm_b200.h"
#include <stdexcept>

void solve_all_contigs_b200(std::vector<std::vector<PafReadData>>& paf_data,
                            std::vector<std::vector<PafOutputData>>& out,
                            std::vector<std::vector<PafOutputData>>& alt_out,
                            std::vector<std::vector<std::vector<PafOutputData>>>& max_out,
                            int device) {
    // ---- pack: structure-of-arrays in FILE order (ctg_index == position, alignasm.cpp:138-139) ----
    std::vector<int64_t> ctg_off{0}, qs, qe, rs, re, qt, run_off{0}, ql, qr, rl;
    std::vector<int32_t> chr; std::vector<uint8_t> fwd, mq;
    for (auto& ctg : paf_data) {
        for (auto& b : ctg) {
            qs.push_back(b.qry_str); qe.push_back(b.qry_end);          // closed intervals (alignasm.cpp:141-151)
            rs.push_back(b.ref_str); re.push_back(b.ref_end);          // ref_str > ref_end on '-' (:155-159)
            qt.push_back(b.qry_total_length); chr.push_back((int32_t)b.ref_chr);
            fwd.push_back(b.aln_fwd ? 1 : 0); mq.push_back((uint8_t)b.map_qul);
            for (size_t k = 0; k < b.qry_overlap_range.size(); k++) {  // get_overlap_range output (paf_data.cpp:90-123)
                ql.push_back(b.qry_overlap_range[k].first); qr.push_back(b.qry_overlap_range[k].second);
                rl.push_back(b.ref_overlap_range[k].first);            // r_r = r_l ± (q_r - q_l), paf_data.cpp:102-105
            }
            run_off.push_back((int64_t)ql.size());
        }
        ctg_off.push_back((int64_t)qs.size());
    }
    aa_batch in{(int64_t)paf_data.size(), (int64_t)qs.size(), (int64_t)ql.size(), ctg_off.data(), qs.data(), qe.data(),
                rs.data(), re.data(), qt.data(), chr.data(), fwd.data(), mq.data(), run_off.data(), ql.data(),
                qr.data(), rl.data()};
    aa_opts opt{NON_SKIP_LINKABLE ? 1 : 0, /*want_all=*/1, /*max_walks=*/0, /*keep_debug=*/0};

    aa_ctx* ctx = nullptr;
    if (aa_create(&ctx, device) != AA_OK) throw std::runtime_error(aa_last_error(nullptr));  // no CPU fallback
    aa_result r{};
    if (aa_solve(ctx, &in, &opt, &r) != AA_OK) { std::string m = aa_last_error(ctx); aa_destroy(ctx); throw std::logic_error(m); }

    // ---- unpack into the containers process_output / process_max_output read (alignasm.cpp:407-490) ----
    auto row = [](const aa_rows& x, int64_t k) {
        PafOutputData o; o.ctg_index = x.ctg_index[k];
        o.edited_qry_str = x.qry_str[k]; o.edited_qry_end = x.qry_end[k];
        o.edited_ref_str = x.ref_str[k]; o.edited_ref_end = x.ref_end[k];
        o.is_alt_path = x.is_alt[k] != 0; return o; };
    int64_t blk = 0;
    for (int64_t c = 0; c < r.n_ctg; c++) {
        for (int64_t k = r.out_off[c]; k < r.out_off[c + 1]; k++) out[c].push_back(row(r.out, k));
        for (int64_t k = r.alt_off[c]; k < r.alt_off[c + 1]; k++) alt_out[c].push_back(row(r.alt, k));
        for (int64_t p = r.all_path_off[c]; p < r.all_path_off[c + 1]; p++) {
            max_out[c].emplace_back();
            for (int64_t k = r.all_row_off[p]; k < r.all_row_off[p + 1]; k++) max_out[c].back().push_back(row(r.all, k));
        }
        for (auto& b : paf_data[c]) b.ctg_sorted_index = r.sorted_index[blk++];   // paf_data.cpp:236,244
    }
    aa_result_free(&r);
    aa_destroy(ctx);
}
