// tools/heap_lab/heap_lab.cu — stand-alone harness for the serial heap builder (NOT product code).
// Loads one contig's insert stream (dumped by an -DAA_HEAP_DUMP build: tools/heap_lab/README), runs builder variants from
// aa_core.cuh on one CTA, checks every vertex's heap against a sequential CPU build (structural hash) and times them.
//   nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -o heap_lab heap_lab.cu
//   ./heap_lab stream.bin [variants...]
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#include "../../alignasm_b200/csrc/aa_core.cuh"
using namespace aa;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

template <int VAR>
__global__ void __launch_bounds__(32) k_build(Ws w) {
    extern __shared__ __align__(16) unsigned char smem[];
    f_heaps_chain(w, 0, smem);
}
__global__ void k_root_fill(Ws w, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) f_root_fill(w, i);
}
__global__ void __launch_bounds__(32) k_leaf(Ws w, int64_t n) {
    if ((int64_t)blockIdx.x < n) f_heaps_level(w, (int64_t)w.leaf_list[blockIdx.x]);
}

// ---- micro latencies (one warp, dependent chains of 256 ops) ----
__global__ void k_micro(long long *out) {
    __shared__ int sm[1024];
    const int lane = threadIdx.x;
    for (int i = lane; i < 1024; i += 32) sm[i] = (i * 7 + 1) & 1023;
    __syncwarp();
    unsigned x = lane * 2654435761u + 12345u;
    long long t0, t1;
    const int N = 256;
    // shfl
    t0 = clock64();
    for (int i = 0; i < N; i++) x = __shfl_sync(0xffffffffu, x, (x >> 3) & 31) + 1;
    t1 = clock64(); if (lane == 0) out[0] = (t1 - t0) / N;
    t0 = clock64();
    for (int i = 0; i < N; i++) x = __ballot_sync(0xffffffffu, (x >> (lane & 7)) & 1) + lane;
    t1 = clock64(); if (lane == 0) out[1] = (t1 - t0) / N;
    t0 = clock64();
    for (int i = 0; i < N; i++) x = __reduce_min_sync(0xffffffffu, x ^ (lane << 3)) + 3;
    t1 = clock64(); if (lane == 0) out[2] = (t1 - t0) / N;
    t0 = clock64();
    for (int i = 0; i < N; i++) x = sm[x & 1023];
    t1 = clock64(); if (lane == 0) out[3] = (t1 - t0) / N;
    t0 = clock64();
    for (int i = 0; i < N; i++) x = __match_any_sync(0xffffffffu, x & 3) + i;
    t1 = clock64(); if (lane == 0) out[4] = (t1 - t0) / N;
    t0 = clock64();
    for (int i = 0; i < N; i++) x = x * 3 + 1;
    t1 = clock64(); if (lane == 0) out[5] = (t1 - t0) / N;
    t0 = clock64();
    for (int i = 0; i < N; i++) x = __reduce_or_sync(0xffffffffu, x & (1u << (lane & 31))) + 1;
    t1 = clock64(); if (lane == 0) out[6] = (t1 - t0) / N;
    unsigned long long y = x;
    t0 = clock64();
    for (int i = 0; i < N; i++) y = (y < 77777ull + i ? y * 3 : y + 1);
    t1 = clock64(); if (lane == 0) out[7] = (t1 - t0) / N;
    t0 = clock64();
    for (int i = 0; i < N; i++) { sm[(x + lane) & 1023] = x; __syncwarp(); x = sm[(x + 5) & 1023] + 1; }
    t1 = clock64(); if (lane == 0) out[8] = (t1 - t0) / N;
    if (x == 0xdeadbeef || y == 42) out[15] = x;
}

struct CNode { int64_t sum; int32_t anom, nz, tot, left, right; int16_t rank, lrank; int32_t eid; };
static inline int64_t den_h(int32_t t) { return t ? t : 1; }
static bool key_lt_h(const CNode &a, const InsKey &k) {
    if (a.sum != k.sum) return a.sum < k.sum;
    if (a.anom != k.anom) return a.anom < k.anom;
    return (int64_t)a.nz * den_h(k.tot) > (int64_t)k.nz * den_h(a.tot);
}
static inline uint64_t mix(uint64_t h, uint64_t v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); return h * 0xff51afd7ed558ccdull; }

int main(int argc, char **argv) {
    if (argc < 2) { printf("usage: heap_lab stream.bin [variant ...]\n"); return 1; }
    FILE *f = fopen(argv[1], "rb");
    if (!f) { printf("cannot open %s\n", argv[1]); return 1; }
    int64_t hdr[4];
    if (fread(hdr, 8, 4, f) != 4) return 1;
    const int64_t nt = hdr[0], ni = hdr[1];
    std::vector<VInfo> vi((size_t)nt);
    std::vector<InsKey> ins((size_t)ni);
    if (fread(vi.data(), sizeof(VInfo), (size_t)nt, f) != (size_t)nt) return 1;
    if (fread(ins.data(), sizeof(InsKey), (size_t)ni, f) != (size_t)ni) return 1;
    fclose(f);
    printf("stream: %ld tree vertices, %ld inserts\n", (long)nt, (long)ni);

    // ---- micro latencies ----
    {
        long long *d; CK(cudaMalloc(&d, 16 * 8)); CK(cudaMemset(d, 0, 128));
        k_micro<<<1, 32>>>(d); CK(cudaDeviceSynchronize());
        k_micro<<<1, 32>>>(d); CK(cudaDeviceSynchronize());
        long long h[16]; CK(cudaMemcpy(h, d, 128, cudaMemcpyDeviceToHost));
        printf("micro (cycles per dependent op): shfl+add %lld  ballot+add %lld  redux.min+xor+add %lld  lds %lld  match_any+add %lld  imad %lld  redux.or %lld  u64 cmp-select %lld  sts+syncwarp+lds %lld\n",
               h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8]);
        cudaFree(d);
    }

    // ---- CPU sequential build ----
    std::vector<CNode> cn; cn.reserve((size_t)ni * 9);
    std::vector<int32_t> croot((size_t)nt, -1);
    std::vector<int32_t> leaf_list;
    for (int64_t pos = 0; pos < nt; pos++) {
        int32_t root = vi[pos].ppos < 0 ? -1 : croot[vi[pos].ppos];
        const int n = vi[pos].nins & (VI_KIDS - 1);
        if (n > 0 && n <= 32 && !(vi[pos].nins & VI_KIDS)) leaf_list.push_back((int32_t)pos);
        for (int k = 0; k < n; k++) {
            const InsKey &ik = ins[vi[pos].ins_beg + k];
            std::vector<int32_t> sp; int32_t a = root;
            while (a >= 0 && key_lt_h(cn[a], ik)) { sp.push_back(a); a = cn[a].right; }
            CNode nn{ik.sum, ik.anom, ik.nz, ik.tot, a, -1, 1, (int16_t)(a >= 0 ? cn[a].rank : 0), ik.eid};
            int32_t r = (int32_t)cn.size(); cn.push_back(nn); int rr = 1;
            for (int q = (int)sp.size() - 1; q >= 0; q--) {
                CNode o = cn[sp[q]];
                int32_t l = o.left, lr = o.lrank, rc = r, rk = rr;
                if (l < 0 || lr < rk) { std::swap(l, rc); std::swap(lr, rk); }
                o.left = l; o.right = rc; o.lrank = (int16_t)lr; o.rank = (int16_t)(rc >= 0 ? rk + 1 : 0);
                r = (int32_t)cn.size(); cn.push_back(o); rr = o.rank;
            }
            root = r;
        }
        croot[pos] = root;
    }
    std::vector<uint64_t> chash(cn.size());
    for (size_t i = 0; i < cn.size(); i++) {
        const CNode &n = cn[i];
        uint64_t h = mix(1, (uint64_t)n.sum); h = mix(h, (uint32_t)n.anom); h = mix(h, (uint32_t)n.nz); h = mix(h, (uint32_t)n.tot);
        h = mix(h, (uint32_t)n.eid); h = mix(h, (uint16_t)n.rank); h = mix(h, (uint16_t)n.lrank);
        h = mix(h, n.left >= 0 ? chash[n.left] : 7); h = mix(h, n.right >= 0 ? chash[n.right] : 11);
        chash[i] = h;
    }
    printf("cpu: %zu nodes, %zu leaves\n", cn.size(), leaf_list.size());

    // ---- device workspace ----
    Ws w; memset(&w, 0, sizeof(w));
    const int64_t Hcap = (int64_t)cn.size() + 64 * (int64_t)ni + (1 << 20);
    auto D = [&](size_t bytes) { void *p; CK(cudaMalloc(&p, bytes ? bytes : 16)); CK(cudaMemset(p, 0, bytes ? bytes : 16)); return p; };
    w.C = 1;
    int64_t h_voff[2] = {0, hdr[2]};
    w.vtx_off = (int64_t *)D(16); CK(cudaMemcpy(w.vtx_off, h_voff, 16, cudaMemcpyHostToDevice));
    w.status = (int32_t *)D(4); w.hmode = (int32_t *)D(4);
    w.ntree = (int32_t *)D(4); { int32_t x = (int32_t)nt; CK(cudaMemcpy(w.ntree, &x, 4, cudaMemcpyHostToDevice)); }
    w.vinfo = (VInfo *)D(sizeof(VInfo) * (size_t)hdr[2]); CK(cudaMemcpy(w.vinfo, vi.data(), sizeof(VInfo) * (size_t)nt, cudaMemcpyHostToDevice));
    w.ins = (InsKey *)D(sizeof(InsKey) * (size_t)ni); CK(cudaMemcpy(w.ins, ins.data(), sizeof(InsKey) * (size_t)ni, cudaMemcpyHostToDevice));
    std::vector<int64_t> ioff((size_t)hdr[2] + 2, ni);
    for (int64_t i = 0; i < nt; i++) ioff[(size_t)i] = vi[(size_t)i].ins_beg;
    w.ins_off = (int64_t *)D(8 * ioff.size()); CK(cudaMemcpy(w.ins_off, ioff.data(), 8 * ioff.size(), cudaMemcpyHostToDevice));
    w.hn = (HNode *)D(sizeof(HNode) * (size_t)Hcap); w.hn_eid = (int32_t *)D(4 * (size_t)Hcap);
    w.hn_key = (unsigned long long *)D(8 * (size_t)Hcap); w.Hcap = Hcap;
    w.heap_top = (unsigned long long *)D(8);
    w.hroot = (int32_t *)D(4 * (size_t)hdr[2]); w.root_at = (int32_t *)D(4 * (size_t)hdr[2]);
    w.vcnt = (int32_t *)D(4 * ((size_t)hdr[2] + 1)); w.leaf_base = (int32_t *)D(4 * (size_t)hdr[2]);
    w.vbase = (int64_t *)D(256);
    w.heap_used = (int64_t *)D(8); w.lvl_overflow = (int32_t *)D(4);
    w.leaf_list = (uint32_t *)D(4 * leaf_list.size());
    CK(cudaMemcpy(w.leaf_list, leaf_list.data(), 4 * leaf_list.size(), cudaMemcpyHostToDevice));

    // ---- operation stream of f_heaps_chain: the product's own pre-pass functions, run here on the host ----
    int64_t n_ops = 0, n_chain = 0;
    {
        Ws h; memset(&h, 0, sizeof(h));
        const int64_t V = hdr[2];
        std::vector<int32_t> z1(1, 0), nt1(1, (int32_t)nt), opc((size_t)V + 1, 0), cf((size_t)V + 1, 0), own((size_t)V, -1), ln((size_t)V + 1, 0),
            cs((size_t)V, 0);
        std::vector<int64_t> opo((size_t)V + 2, 0), co((size_t)V + 2, 0), lp((size_t)V + 2, 0);
        std::vector<VInfo> hv((size_t)V); memcpy(hv.data(), vi.data(), sizeof(VInfo) * (size_t)nt);
        h.C = 1; h.vtx_off = h_voff; h.hmode = z1.data(); h.status = z1.data(); h.ntree = nt1.data(); h.vinfo = hv.data(); h.ins = ins.data();
        h.op_cnt = opc.data(); h.op_off = opo.data(); h.chain_flag = cf.data(); h.chain_ord = co.data(); h.owner = own.data();
        h.leaf_need = ln.data(); h.leaf_lp = lp.data(); h.chain_slot = cs.data();
        for (int64_t i = 0; i < V; i++) f_ops_class(h, i);
        for (int64_t i = 0; i < V; i++) { opo[(size_t)i + 1] = opo[(size_t)i] + opc[(size_t)i]; co[(size_t)i + 1] = co[(size_t)i] + cf[(size_t)i];
                                           lp[(size_t)i + 1] = lp[(size_t)i] + ln[(size_t)i]; }
        for (int64_t i = 0; i < V; i++) f_chain_slot(h, i);
        n_ops = opo[(size_t)V]; n_chain = co[(size_t)V];
        // pointer jumping, worst order for an in-place update (descending), until nothing changes
        for (int round = 0;; round++) {
            std::vector<int32_t> before = own;
            for (int64_t i = V - 1; i >= 0; i--) f_owner_jump(h, i);
            if (before == own) { printf("owner: %d jumping rounds\n", round); break; }
        }
        std::vector<HOp> hops((size_t)n_ops);
        h.ops = hops.data();
        for (int64_t i = 0; i < V; i++) f_ops_fill(h, i);
        w.ops = (HOp *)D(sizeof(HOp) * (size_t)n_ops); CK(cudaMemcpy(w.ops, hops.data(), sizeof(HOp) * (size_t)n_ops, cudaMemcpyHostToDevice));
        w.op_cnt = (int32_t *)D(4 * opc.size()); CK(cudaMemcpy(w.op_cnt, opc.data(), 4 * opc.size(), cudaMemcpyHostToDevice));
        w.op_off = (int64_t *)D(8 * opo.size()); CK(cudaMemcpy(w.op_off, opo.data(), 8 * opo.size(), cudaMemcpyHostToDevice));
        w.chain_flag = (int32_t *)D(4 * cf.size()); CK(cudaMemcpy(w.chain_flag, cf.data(), 4 * cf.size(), cudaMemcpyHostToDevice));
        w.chain_ord = (int64_t *)D(8 * co.size()); CK(cudaMemcpy(w.chain_ord, co.data(), 8 * co.size(), cudaMemcpyHostToDevice));
        w.owner = (int32_t *)D(4 * own.size()); CK(cudaMemcpy(w.owner, own.data(), 4 * own.size(), cudaMemcpyHostToDevice));
        w.chain_root = (int32_t *)D(4 * (size_t)n_chain + 4);
        w.leaf_need = (int32_t *)D(4 * ln.size()); CK(cudaMemcpy(w.leaf_need, ln.data(), 4 * ln.size(), cudaMemcpyHostToDevice));
        w.leaf_lp = (int64_t *)D(8 * lp.size()); CK(cudaMemcpy(w.leaf_lp, lp.data(), 8 * lp.size(), cudaMemcpyHostToDevice));
        w.chain_slot = (int32_t *)D(4 * cs.size()); CK(cudaMemcpy(w.chain_slot, cs.data(), 4 * cs.size(), cudaMemcpyHostToDevice));
        w.resv_base = (int32_t *)D(4 * (size_t)n_chain + 64);
        printf("ops: %ld (chain vertices %ld)\n", (long)n_ops, (long)n_chain);
    }
    std::vector<int> variants;
    for (int i = 2; i < argc; i++) variants.push_back(atoi(argv[i]));
    if (variants.empty()) variants.push_back(110);
    cudaEvent_t e0, e1, e2; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2));
    std::vector<HNode> hn((size_t)Hcap); std::vector<int32_t> heid((size_t)Hcap), hroot_at((size_t)nt);
    std::vector<uint64_t> gh((size_t)Hcap);
    int rc = 0;
    for (int var : variants) {
        float best = 1e30f, bestl = 0;
        // variant 100*flags + bits: f_heaps_chain with a node cache of 1 << bits entries (bits 0: none)
        const int bits = var % 100;
        size_t smem = heaps_chain_smem_bytes(bits);
        w.heap_cache_bits = bits;
        for (int rep = 0; rep < 3; rep++) {
            CK(cudaMemset(w.heap_top, 0, 8)); CK(cudaMemset(w.status, 0, 4)); CK(cudaMemset(w.heap_used, 0, 8));
            CK(cudaMemset(w.hn, 0xff, sizeof(HNode) * (size_t)Hcap));
            CK(cudaMemset(w.vcnt, 0, 4 * ((size_t)hdr[2] + 1)));
            w.heaps_variant = var / 100;
            CK(cudaEventRecord(e0));
            CK(cudaFuncSetAttribute(k_build<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_build<1><<<1, 32, smem>>>(w);
            k_root_fill<<<(unsigned)((hdr[2] + 255) / 256), 256>>>(w, hdr[2]);
            CK(cudaEventRecord(e1));
            k_leaf<<<(unsigned)leaf_list.size(), 32>>>(w, (int64_t)leaf_list.size());
            CK(cudaEventRecord(e2));
            CK(cudaDeviceSynchronize());
            float ms, msl; CK(cudaEventElapsedTime(&ms, e0, e1)); CK(cudaEventElapsedTime(&msl, e1, e2));
            if (ms < best) { best = ms; bestl = msl; }
        }
        int32_t st; int64_t used; unsigned long long top;
        CK(cudaMemcpy(&st, w.status, 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&used, w.heap_used, 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(&top, w.heap_top, 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hn.data(), w.hn, sizeof(HNode) * (size_t)top, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(heid.data(), w.hn_eid, 4 * (size_t)top, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hroot_at.data(), w.root_at, 4 * (size_t)nt, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < (size_t)top; i++) {
            const HNode &n = hn[i];
            uint64_t h = mix(1, (uint64_t)n.sum); h = mix(h, (uint32_t)n.anom); h = mix(h, (uint32_t)n.nz); h = mix(h, (uint32_t)n.tot);
            h = mix(h, (uint32_t)heid[i]); h = mix(h, (uint16_t)n.rank); h = mix(h, (uint16_t)n.lrank);
            const bool okl = n.left < (int32_t)i, okr = n.right < (int32_t)i;
            h = mix(h, n.left >= 0 ? (okl ? gh[(size_t)n.left] : 13) : 7); h = mix(h, n.right >= 0 ? (okr ? gh[(size_t)n.right] : 17) : 11);
            gh[i] = h;
        }
        int64_t bad = 0, firstbad = -1;
        for (int64_t pos = 0; pos < nt; pos++) {
            const int32_t a = croot[(size_t)pos], b = hroot_at[(size_t)pos];
            const bool same = (a < 0 && b < 0) || (a >= 0 && b >= 0 && (uint64_t)b < top && chash[(size_t)a] == gh[(size_t)b]);
            if (!same) { if (firstbad < 0) firstbad = pos; bad++; }
        }
        printf("variant %d: %.3f ms (+ leaves %.3f ms)  status %d  used %ld (cpu %zu)  top %llu  smem %zu  mismatching vertices %ld (first %ld)  %s\n", var, best, bestl, st,
               (long)used, cn.size(), top, smem, (long)bad, (long)firstbad, bad == 0 ? "OK" : "FAIL");
#ifdef AA_HEAP_TIMERS
        {
            int64_t t[16]; CK(cudaMemcpy(t, w.vbase, 128, cudaMemcpyDeviceToHost));
            const char *nm[8] = {"op-read", "reserve", "switch", "descent", "ranks", "stores", "newspine", "extend(count)"};
            for (int i = 0; i < 8; i++) printf("   %-14s cycles %10ld  n %8ld  avg %7.1f\n", nm[i], (long)t[2 * i], (long)t[2 * i + 1], t[2 * i + 1] ? (double)t[2 * i] / t[2 * i + 1] : 0.0);
        }
#endif
        if (bad) rc = 1;
    }
    return rc;
}
