// tests/emul/emul_backend.cpp — TEST INFRASTRUCTURE ONLY.
// Runs the product's phase functions (alignasm_b200/csrc/aa_core.cuh + aa_pipeline.cuh) through a
// host-loop Backend so that their logic can be checked on a machine without a GPU.  This is NOT a
// product path: it is compiled only by tests/ into tests/emul/_build/libaa_emul.so and exports
// emul_solve(), which nothing in alignasm_b200/ links or loads.
#include "../../alignasm_b200/csrc/aa_pipeline.cuh"

#include <cstdio>
#include <numeric>

namespace {
struct HostBackend {
    std::vector<void *> blocks;
    std::string none;
    bool ok() const { return true; }
    const std::string &error() const { return none; }
    void *alloc_bytes(size_t n) {
        void *p = std::malloc(n ? n : 1);
        blocks.push_back(p);
        return p;
    }
    size_t alloc_mark() const { return blocks.size(); }
    void release_to(size_t mark) { release_last((int)(blocks.size() - mark)); }
    void release_last(int k) {
        while (k-- > 0 && !blocks.empty()) {
            std::free(blocks.back());
            blocks.pop_back();
        }
    }
    void *alloc_persistent(size_t n) { return std::malloc(n ? n : 1); }
    void free_persistent(void *p) { std::free(p); }
    void h2d(void *d, const void *h, size_t n) { std::memcpy(d, h, n); }
    void stage(void *d, const void *h, size_t n) { std::memcpy(d, h, n); }
    void flush_staged() {}
    void d2h(void *h, const void *d, size_t n) { std::memcpy(h, d, n); }
    void stage_d2h(void *h, const void *d, size_t n) { std::memcpy(h, d, n); }
    void flush_d2h(bool = false) {}
    void *result_slab(size_t) { return nullptr; }  // pageable arrays
    void zero(void *p, size_t n) { std::memset(p, 0, n); }
    void sync() {}
    template <class F>
    void for_each(const char *, int64_t n, F f) {
        for (int64_t i = 0; i < n; i++) f(i, nullptr);
    }
    template <class F>
    void for_each_contig(const char *, int64_t n, F f, size_t smem = 0) {
        std::vector<unsigned char> scratch(smem + 16);
        for (int64_t i = 0; i < n; i++) f(i, scratch.data());
    }
    template <class F>
    void workers(const char *, int64_t n, F f, size_t smem = 0) {
        std::vector<unsigned char> scratch(smem + 16);
        for (int64_t i = 0; i < n; i++) f(i, scratch.data());
    }
    void scan_i32(const int32_t *in, int64_t *out, int64_t n) {
        int64_t s = 0;
        for (int64_t i = 0; i < n; i++) {
            out[i] = s;
            s += in[i];
        }
    }
    void sort_pairs_u32(const uint32_t *kin, uint32_t *kout, const uint32_t *vin, uint32_t *vout, int64_t n, int) {
        std::vector<int64_t> idx((size_t)n);
        std::iota(idx.begin(), idx.end(), 0);
        std::stable_sort(idx.begin(), idx.end(), [&](int64_t a, int64_t b) { return kin[a] < kin[b]; });
        for (int64_t i = 0; i < n; i++) {
            kout[i] = kin[idx[(size_t)i]];
            vout[i] = vin[idx[(size_t)i]];
        }
    }
    void sort_pairs_u64(const uint64_t *kin, uint64_t *kout, const uint32_t *vin, uint32_t *vout, int64_t n, int) {
        std::vector<int64_t> idx((size_t)n);
        std::iota(idx.begin(), idx.end(), 0);
        std::stable_sort(idx.begin(), idx.end(), [&](int64_t a, int64_t b) { return kin[a] < kin[b]; });
        for (int64_t i = 0; i < n; i++) {
            kout[i] = kin[idx[(size_t)i]];
            vout[i] = vin[idx[(size_t)i]];
        }
    }
    void fill_ff(void *p, size_t n) { std::memset(p, 0xff, n); }
    int64_t read_i64(const int64_t *p) { return *p; }
    bool device_kahn() const { return false; }
    void aux_begin() {}
    void aux_enter() {}
    void aux_end() {}
    void aux_join() {}
    void side_begin() {}
    void side_end() {}
    void side_join() {}
    int host_threads() { return 1; }
    int64_t max_workers() { return 3; }  // >1 so that slot indexing is exercised
    int64_t scratch_budget() { return (int64_t)1 << 30; }
    void begin_solve(bool = false) {}
    void end_solve(aa_stats &) {
        for (void *p : blocks) std::free(p);
        blocks.clear();
    }
    void phase_begin(int) {}
    void phase_end(int) {}
};
}  // namespace

extern "C" int emul_solve(const aa_batch *b, const aa_opts *opt, aa_result *res) {
    HostBackend bk;
    aa::Pipeline<HostBackend> p(bk);
    aa::DevBatch *d = nullptr;
    aa_status st = p.upload(b, d);
    if (st != AA_OK) return st;
    aa_opts o{};
    if (opt) o = *opt;
    st = p.solve(*d, o, res);
    if (st != AA_OK && st != AA_ERR_UNSOLVABLE) bk.end_solve(p.last_stats);
    p.free_batch(d);
    if (st != AA_OK) std::fprintf(stderr, "emul_solve: %s\n", p.err.c_str());
    return st;
}
extern "C" void emul_result_free(aa_result *res) { aa::result_free_host(res); }
