// aa_backend_cuda.cuh — the product Backend: sm_100a kernels, CUB scan / radix sort, pooled HBM workspace.
//
// Launch shapes:
//   for_each         one thread per item (blocks, candidate slots, vertices), 256-thread CTAs
//   for_each_contig  one warp per contig, one CTA per warp so that the block scheduler spreads the
//                    (largest-first ordered) contigs over all 148 SMs
//   workers          persistent one-warp CTAs pulling walk tasks from a global counter
// The workspace is a bump allocator over a few large cudaMalloc blocks that live in the aa_ctx, so a
// steady-state solve performs no cudaMalloc at all.
#pragma once
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>
#include <thrust/iterator/reverse_iterator.h>

#include <atomic>
#include <chrono>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

#include "aa_pipeline.cuh"

namespace aa {

template <class F>
__global__ void __launch_bounds__(256) k_items(int64_t n, F f) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) f(i, nullptr);
}
// one warp per item (contig / worker slot)
template <class F>
__global__ void __launch_bounds__(32) k_warp_items(int64_t n, F f) {
    extern __shared__ __align__(16) unsigned char aa_smem[];
    int64_t i = blockIdx.x;
    if (i < n) f(i, (void *)aa_smem);  // all 32 lanes enter; sequential phases continue on lane 0 only
}

struct MaxI64 {
    __host__ __device__ __forceinline__ int64_t operator()(int64_t a, int64_t b) const { return a > b ? a : b; }
};
struct MaxI32 {
    __host__ __device__ __forceinline__ int32_t operator()(int32_t a, int32_t b) const { return a > b ? a : b; }
};
struct MinI32 {
    __host__ __device__ __forceinline__ int32_t operator()(int32_t a, int32_t b) const { return a < b ? a : b; }
};
struct CastI32 {
    __host__ __device__ __forceinline__ int64_t operator()(const int32_t &x) const { return (int64_t)x; }
};

// Pinned host slabs for result arrays, shared by every context of the process (results outlive solves and contexts).
// A released slab is kept for the next result of about the same size; at most KEEP_BYTES stay cached.
class ResultSlabs {
  public:
    static ResultSlabs &get() {
        static ResultSlabs *p = new ResultSlabs();  // never destroyed: results may be freed during process exit
        return *p;
    }
    void *acquire(size_t bytes) {
        if (bytes == 0) bytes = 256;
        if (bytes > MAX_SLAB) return nullptr;  // (.aln.all rows of a large input: tens of GB are not worth pinning)
        {
            std::lock_guard<std::mutex> g(mu);
            size_t best = SIZE_MAX;
            for (size_t i = 0; i < idle.size(); i++)
                if (idle[i].cap >= bytes && idle[i].cap <= 4 * bytes + (1u << 20) && (best == SIZE_MAX || idle[i].cap < idle[best].cap)) best = i;
            if (best != SIZE_MAX) {
                Slab sl = idle[best];
                idle.erase(idle.begin() + (long)best);
                idle_bytes -= sl.cap;
                busy[sl.base] = sl.cap;
                return sl.base;
            }
        }
        void *base = nullptr;
        const size_t cap = bytes + bytes / 8;
        if (cudaHostAlloc(&base, cap, cudaHostAllocPortable) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;  // the caller falls back to pageable arrays
        }
        std::lock_guard<std::mutex> g(mu);
        busy[base] = cap;
        return base;
    }
    bool release(void *base) {
        std::vector<void *> drop;
        {
            std::lock_guard<std::mutex> g(mu);
            auto it = busy.find(base);
            if (it == busy.end()) return false;
            idle.push_back({base, it->second});
            idle_bytes += it->second;
            busy.erase(it);
            while (idle_bytes > KEEP_BYTES && !idle.empty()) {  // oldest first
                drop.push_back(idle.front().base);
                idle_bytes -= idle.front().cap;
                idle.erase(idle.begin());
            }
        }
        for (void *q : drop)
            if (cudaFreeHost(q) != cudaSuccess) cudaGetLastError();
        return true;
    }

  private:
    struct Slab {
        void *base;
        size_t cap;
    };
    static constexpr size_t KEEP_BYTES = (size_t)2 << 30;
    static constexpr size_t MAX_SLAB = (size_t)1 << 30;
    std::mutex mu;
    std::vector<Slab> idle;
    size_t idle_bytes = 0;
    std::unordered_map<void *, size_t> busy;
};
inline bool result_slab_release(void *base) { return ResultSlabs::get().release(base); }

struct CudaBackend {
    int device = 0;
    cudaStream_t stream = nullptr;   // the stream launches currently go to
    cudaStream_t main_stream = nullptr, side_stream = nullptr, aux_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_side = nullptr, ev_aux0 = nullptr, ev_aux1 = nullptr;
    bool side_pending = false;
    bool failed = false;
    bool oom = false;  // the failure was an allocation: recoverable
    std::string errmsg;
    int64_t n_launch = 0;
    int sm_count = 148;
    size_t free_at_init = (size_t)16 << 30;

    struct Block {
        char *base;
        size_t cap, top;
    };
    std::vector<Block> pool;
    struct Mark {
        int block;
        size_t top;
    };
    std::vector<Mark> log;  // one entry per allocation (for release_last)
    cudaEvent_t ev[PH_COUNT][2];
    bool ev_on[PH_COUNT];
    cudaEvent_t ev_total[2];
    // AA_TRACE=1: one event after every launch, printed as a timeline (ms since the start of the solve) by end_solve
    bool trace = false;
    struct TraceMark {
        const char *name;
        cudaEvent_t ev;
        bool side;
        double host_ms;  // host clock when the launch was issued
    };
    std::chrono::steady_clock::time_point t_host0;
    std::vector<TraceMark> marks;
    size_t n_marks = 0;
    void mark(const char *name) {
        if (!trace || failed) return;
        if (n_marks == marks.size()) {
            TraceMark m{name, nullptr, false, 0.0};
            cudaEventCreate(&m.ev);
            marks.push_back(m);
        }
        marks[n_marks].name = name;
        marks[n_marks].side = stream == side_stream;
        marks[n_marks].host_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_host0).count();
        cudaEventRecord(marks[n_marks].ev, stream);
        n_marks++;
    }

    bool ok() const { return !failed; }
    const std::string &error() const { return errmsg; }
    void fail(const char *what, cudaError_t e) {
        if (!failed) {
            failed = true;
            errmsg = std::string(what) + ": " + cudaGetErrorString(e);
        }
    }
#define AA_CUDA(call)                          \
    do {                                       \
        cudaError_t e__ = (call);              \
        if (e__ != cudaSuccess) fail(#call, e__); \
    } while (0)

    bool init(int dev) {
        device = dev;
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n <= 0 || dev < 0 || dev >= n) {
            errmsg = e != cudaSuccess ? std::string("no usable CUDA device: ") + cudaGetErrorString(e)
                                      : "no usable CUDA device (count " + std::to_string(n) + ", asked for " + std::to_string(dev) + ")";
            failed = true;
            return false;
        }
        AA_CUDA(cudaSetDevice(dev));
        cudaDeviceProp prop;
        AA_CUDA(cudaGetDeviceProperties(&prop, dev));
        sm_count = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : 148;
        {
            size_t fr = 0, tot = 0;
            if (cudaMemGetInfo(&fr, &tot) == cudaSuccess) free_at_init = fr;
            else cudaGetLastError();
        }
        AA_CUDA(cudaStreamCreateWithFlags(&main_stream, cudaStreamNonBlocking));
        AA_CUDA(cudaStreamCreateWithFlags(&side_stream, cudaStreamNonBlocking));
        AA_CUDA(cudaStreamCreateWithFlags(&aux_stream, cudaStreamNonBlocking));
        AA_CUDA(cudaEventCreateWithFlags(&ev_aux0, cudaEventDisableTiming));
        AA_CUDA(cudaEventCreateWithFlags(&ev_aux1, cudaEventDisableTiming));
        AA_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
        AA_CUDA(cudaEventCreateWithFlags(&ev_side, cudaEventDisableTiming));
        stream = main_stream;
        for (int p = 0; p < PH_COUNT; p++) {
            AA_CUDA(cudaEventCreate(&ev[p][0]));
            AA_CUDA(cudaEventCreate(&ev[p][1]));
            ev_on[p] = false;
        }
        AA_CUDA(cudaEventCreate(&ev_total[0]));
        AA_CUDA(cudaEventCreate(&ev_total[1]));
        const char *tr = std::getenv("AA_TRACE");
        trace = tr && tr[0] && tr[0] != '0';
        return !failed;
    }
    void shutdown() {
        if (stream) cudaStreamSynchronize(stream);
        for (auto &b : pool) cudaFree(b.base);
        pool.clear();
        if (pinned) cudaFreeHost(pinned);
        pinned = nullptr;
        if (stream) {
            for (int p = 0; p < PH_COUNT; p++) {
                cudaEventDestroy(ev[p][0]);
                cudaEventDestroy(ev[p][1]);
            }
            cudaEventDestroy(ev_total[0]);
            cudaEventDestroy(ev_total[1]);
            cudaStreamSynchronize(side_stream);
            cudaEventDestroy(ev_fork);
            cudaEventDestroy(ev_side);
            cudaStreamDestroy(side_stream);
            if (aux_stream) {
                cudaStreamSynchronize(aux_stream);
                cudaEventDestroy(ev_aux0);
                cudaEventDestroy(ev_aux1);
                cudaStreamDestroy(aux_stream);
                aux_stream = nullptr;
            }
            cudaStreamDestroy(main_stream);
            stream = main_stream = side_stream = nullptr;
        }
    }

    // ---- pooled workspace ----
    void *alloc_bytes(size_t n) {
        if (failed) return nullptr;
        n = (n + 255) & ~(size_t)255;
        if (n == 0) n = 256;
        if (pool.empty() || pool.back().top + n > pool.back().cap) {
            // geometric growth up to 4 GB per block; larger requests get a block of exactly their size (no waste)
            size_t last = pool.empty() ? 0 : pool.back().cap;
            size_t want = std::max<size_t>(n, std::max<size_t>(std::min<size_t>(2 * last, (size_t)4 << 30), (size_t)256 << 20));
            char *p = nullptr;
            cudaError_t e = cudaMalloc(&p, want);
            if (e != cudaSuccess && want > n) {
                cudaGetLastError();
                want = n;
                e = cudaMalloc(&p, want);
            }
            if (e != cudaSuccess) {
                cudaGetLastError();
                oom = true;  // the context stays usable: the next solve starts from an empty pool
                fail("cudaMalloc(workspace)", e);
                return nullptr;
            }
            pool.push_back({p, want, 0});
        }
        Block &b = pool.back();
        log.push_back({(int)pool.size() - 1, b.top});
        void *r = b.base + b.top;
        b.top += n;
        return r;
    }
    size_t alloc_mark() const { return log.size(); }
    void release_to(size_t mark) { release_last((int)(log.size() - mark)); }
    void release_last(int k) {
        while (k-- > 0 && !log.empty()) {
            Mark m = log.back();
            log.pop_back();
            // allocations are a stack: the one being released lives in the newest block
            if (m.block == (int)pool.size() - 1) {
                pool[(size_t)m.block].top = m.top;
                if (m.top == 0 && pool.size() > 1) {  // the block is empty again: hand it back (a regrown arena needs the room)
                    cudaStreamSynchronize(main_stream);
                    cudaFree(pool.back().base);
                    pool.pop_back();
                }
            }
        }
    }
    void *alloc_persistent(size_t n) {
        void *p = nullptr;
        cudaError_t e = cudaMalloc(&p, n ? n : 256);
        if (e != cudaSuccess) {
            cudaGetLastError();
            oom = true;  // recoverable: the next solve / upload starts from an empty pool
            fail("cudaMalloc(batch)", e);
            return nullptr;
        }
        return p;
    }
    // after a solve that stopped half-way: nothing may still be running on either stream when the pool is reused
    void quiesce() {
        cudaStreamSynchronize(main_stream);
        cudaStreamSynchronize(side_stream);
        if (aux_stream) cudaStreamSynchronize(aux_stream);
        side_pending = false;
        stream = main_stream;
    }
    void free_persistent(void *p) { cudaFree(p); }

    // give the whole workspace back to the bump allocator (coalescing a fragmented pool into one block)
    void reset_pool() {
        if (failed && oom) {  // recover from an out-of-memory solve: drop the whole pool and start over
            cudaDeviceSynchronize();
            cudaGetLastError();
            for (auto &b : pool) cudaFree(b.base);
            pool.clear();
            failed = oom = false;
            errmsg.clear();
        }
        AA_CUDA(cudaSetDevice(device));
        if (pool.size() > 1) {
            AA_CUDA(cudaStreamSynchronize(main_stream));
            AA_CUDA(cudaStreamSynchronize(side_stream));
            size_t total = 0;
            for (auto &b : pool) {
                total += b.cap;
                cudaFree(b.base);
            }
            pool.clear();
            char *p = nullptr;
            if (cudaMalloc(&p, total) == cudaSuccess) pool.push_back({p, total, 0});
            else cudaGetLastError();
        }
        for (auto &b : pool) b.top = 0;
        log.clear();
    }
    void begin_solve(bool keep_pool = false) {
        AA_CUDA(cudaSetDevice(device));
        stream = main_stream;
        if (side_pending) AA_CUDA(cudaStreamSynchronize(side_stream));
        side_pending = false;
        if (!keep_pool) reset_pool();
        n_launch = 0;
        n_marks = 0;
        t_host0 = std::chrono::steady_clock::now();
        for (int p = 0; p < PH_COUNT; p++) ev_on[p] = false;
        AA_CUDA(cudaEventRecord(ev_total[0], stream));
    }
    void end_solve(aa_stats &st) {
        AA_CUDA(cudaEventRecord(ev_total[1], stream));
        AA_CUDA(cudaStreamSynchronize(stream));
        float ms = 0;
        if (!failed) {
            cudaEventElapsedTime(&ms, ev_total[0], ev_total[1]);
            st.ms_total = ms;
            for (int p = 0; p < PH_COUNT; p++)
                if (ev_on[p]) {
                    cudaEventElapsedTime(&ms, ev[p][0], ev[p][1]);
                    st.ms_phase[p] = ms;
                }
        }
        st.n_launch = n_launch;
        if (trace && !failed) {
            cudaStreamSynchronize(side_stream);
            float prev[2] = {0, 0};
            for (size_t i = 0; i < n_marks; i++) {
                float t = 0;
                cudaEventElapsedTime(&t, ev_total[0], marks[i].ev);
                const int s = marks[i].side ? 1 : 0;
                std::fprintf(stderr, "[aa_trace] %-6s %-14s end %9.3f ms  (+%.3f)  issued at host %9.3f ms\n", s ? "side" : "main", marks[i].name, t, t - prev[s], marks[i].host_ms);
                prev[s] = t;
            }
        }
    }
    void phase_begin(int p) { AA_CUDA(cudaEventRecord(ev[p][0], stream)); }
    void phase_end(int p) {
        AA_CUDA(cudaEventRecord(ev[p][1], stream));
        ev_on[p] = true;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) fail(phase_name(p), e);
    }

    // ---- staged upload of many host arrays: parallel memcpy into one pinned buffer, then asynchronous copies.  A pageable
    // cudaMemcpy runs at ~6 GB/s on the calling thread; this runs at the host's memcpy bandwidth + PCIe / C2C speed
    char *pinned = nullptr;
    size_t pinned_cap = 0;
    struct Piece {
        void *dst;
        const void *src;
        size_t n, off;
    };
    std::vector<Piece> pieces;
    size_t staged_bytes = 0, staged_tail = 0;  // padded total / end of the last piece
    // page-locked source (aa_host_alloc, cudaHostRegister, a pinned torch tensor ...): the DMA engine reads it as it is
    static bool host_pinned(const void *h) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, h) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        return at.type == cudaMemoryTypeHost;
    }
    void stage(void *d, const void *h, size_t n) {
        if (!n) return;
        if (n >= ((size_t)256 << 10) && host_pinned(h)) {  // (small pieces are not worth the query)
            // (every entry point that stages synchronises the stream before it returns: aa_solve / aa_solve_subset when they
            // fetch the rows or quiesce after a failure, aa_upload explicitly)
            if (!failed) AA_CUDA(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, stream));
            return;
        }
        // a piece that continues the previous one on the device also continues it in the staging buffer, so that the two
        // go up in ONE copy (a shard staged contig by contig is thousands of pieces per array)
        const bool cont = !pieces.empty() && (char *)pieces.back().dst + pieces.back().n == (char *)d &&
                          pieces.back().off + pieces.back().n == staged_tail;
        const size_t off = cont ? staged_tail : ((staged_tail + 255) & ~(size_t)255);
        pieces.push_back({d, h, n, off});
        staged_tail = off + n;
        staged_bytes = (staged_tail + 255) & ~(size_t)255;
    }
    void flush_staged() {
        if (failed || pieces.empty()) {
            pieces.clear();
            staged_bytes = staged_tail = 0;
            return;
        }
        if (staged_bytes > pinned_cap) {
            if (pinned) cudaFreeHost(pinned);
            pinned = nullptr;
            pinned_cap = 0;
            const size_t want = staged_bytes + staged_bytes / 4;
            if (cudaHostAlloc((void **)&pinned, want, cudaHostAllocDefault) == cudaSuccess) pinned_cap = want;
            else cudaGetLastError();
        }
        if (!pinned) {  // no pinned memory to be had: plain copies
            for (auto &p : pieces) h2d(p.dst, p.src, p.n);
        } else {
            const size_t CH = (size_t)4 << 20;  // work items of 4 MB, spread over the host threads
            struct Item {
                char *d;
                const char *s;
                size_t n;
            };
            std::vector<Item> items;
            for (auto &p : pieces)
                for (size_t o = 0; o < p.n; o += CH) items.push_back({pinned + p.off + o, (const char *)p.src + o, std::min(CH, p.n - o)});
            const int nt = (int)std::min<size_t>((size_t)std::min(host_threads(), 8), items.size());
            std::atomic<size_t> next{0};
            auto work = [&]() {
                for (;;) {
                    const size_t i = next.fetch_add(1);
                    if (i >= items.size()) break;
                    std::memcpy(items[i].d, items[i].s, items[i].n);
                }
            };
            std::vector<std::thread> pool;
            for (int t = 1; t < nt; t++) pool.emplace_back(work);
            work();
            for (auto &t : pool) t.join();
            for (size_t i = 0; i < pieces.size();) {  // one copy per run of contiguous pieces
                size_t j = i + 1, n = pieces[i].n;
                while (j < pieces.size() && (char *)pieces[j - 1].dst + pieces[j - 1].n == (char *)pieces[j].dst &&
                       pieces[j - 1].off + pieces[j - 1].n == pieces[j].off) {
                    n += pieces[j].n;
                    j++;
                }
                AA_CUDA(cudaMemcpyAsync(pieces[i].dst, pinned + pieces[i].off, n, cudaMemcpyHostToDevice, stream));
                i = j;
            }
        }
        pieces.clear();
        staged_bytes = staged_tail = 0;
    }
    // ---- staged download (result rows): asynchronous copies into the pinned buffer, one synchronisation, then a parallel
    // memcpy into the caller's (fresh, pageable) arrays: their pages are first touched by all host threads
    std::vector<Piece> dpieces;
    size_t dstaged_bytes = 0;
    void *result_slab(size_t bytes) { return ResultSlabs::get().acquire(bytes); }
    void stage_d2h(void *h, const void *d, size_t n) {
        if (!n) return;
        dpieces.push_back({h, d, n, dstaged_bytes});
        dstaged_bytes += (n + 255) & ~(size_t)255;
    }
    void flush_d2h(bool dst_pinned = false) {
        if (failed || dpieces.empty()) {
            dpieces.clear();
            dstaged_bytes = 0;
            return;
        }
        if (dst_pinned) {  // the destinations are pinned (a result slab): straight copies, one synchronisation
            for (auto &p : dpieces) AA_CUDA(cudaMemcpyAsync(p.dst, p.src, p.n, cudaMemcpyDeviceToHost, stream));
            AA_CUDA(cudaStreamSynchronize(stream));
            dpieces.clear();
            dstaged_bytes = 0;
            return;
        }
        if (dstaged_bytes > pinned_cap) {
            AA_CUDA(cudaStreamSynchronize(stream));  // the upload of this solve may still read the old buffer
            if (pinned) cudaFreeHost(pinned);
            pinned = nullptr;
            pinned_cap = 0;
            const size_t want = dstaged_bytes + dstaged_bytes / 4;
            if (cudaHostAlloc((void **)&pinned, want, cudaHostAllocDefault) == cudaSuccess) pinned_cap = want;
            else cudaGetLastError();
        }
        if (!pinned) {
            for (auto &p : dpieces) d2h(p.dst, p.src, p.n);
        } else {
            for (auto &p : dpieces) AA_CUDA(cudaMemcpyAsync(pinned + p.off, p.src, p.n, cudaMemcpyDeviceToHost, stream));
            AA_CUDA(cudaStreamSynchronize(stream));
            const size_t CH = (size_t)2 << 20;
            struct Item {
                char *d;
                const char *s;
                size_t n;
            };
            std::vector<Item> items;
            for (auto &p : dpieces)
                for (size_t o = 0; o < p.n; o += CH) items.push_back({(char *)p.dst + o, pinned + p.off + o, std::min(CH, p.n - o)});
            const int nt = (int)std::min<size_t>((size_t)std::min(host_threads(), 8), items.size());
            std::atomic<size_t> next{0};
            auto work = [&]() {
                for (;;) {
                    const size_t i = next.fetch_add(1);
                    if (i >= items.size()) break;
                    std::memcpy(items[i].d, items[i].s, items[i].n);
                }
            };
            std::vector<std::thread> pool;
            for (int t = 1; t < nt; t++) pool.emplace_back(work);
            work();
            for (auto &t : pool) t.join();
        }
        dpieces.clear();
        dstaged_bytes = 0;
    }
    void h2d(void *d, const void *h, size_t n) {
        if (n && !failed) AA_CUDA(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, stream));
    }
    void d2h(void *h, const void *d, size_t n) {
        if (n && !failed) {
            AA_CUDA(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, stream));
            AA_CUDA(cudaStreamSynchronize(stream));
        }
    }
    void zero(void *p, size_t n) {
        if (n && !failed) AA_CUDA(cudaMemsetAsync(p, 0, n, stream));
    }
    void sync() { AA_CUDA(cudaStreamSynchronize(stream)); }
    int64_t read_i64(const int64_t *p) {
        int64_t v = 0;
        d2h(&v, p, 8);
        return v;
    }

    template <class F>
    void for_each(const char *name, int64_t n, F f) {
        if (n <= 0 || failed) return;
        int64_t grid = (n + 255) / 256;
        k_items<F><<<(unsigned)grid, 256, 0, stream>>>(n, f);
        n_launch++;
        mark(name);
    }
    template <class F>
    void for_each_contig(const char *name, int64_t n, F f, size_t smem = 0) {
        if (n <= 0 || failed) return;
        if (smem > 48 * 1024) AA_CUDA(cudaFuncSetAttribute(k_warp_items<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_warp_items<F><<<(unsigned)n, 32, smem, stream>>>(n, f);
        n_launch++;
        mark(name);
    }
    template <class F>
    void workers(const char *name, int64_t n, F f, size_t smem = 0) {
        if (n <= 0 || failed) return;
        k_warp_items<F><<<(unsigned)n, 32, smem, stream>>>(n, f);
        n_launch++;
        mark(name);
    }
    void scan_i32(const int32_t *in, int64_t *out, int64_t n) {
        if (n <= 0 || failed) return;
        cub::TransformInputIterator<int64_t, CastI32, const int32_t *> it(in, CastI32());
        size_t tmp = 0;
        AA_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, it, out, n, stream));
        void *t = alloc_bytes(tmp);
        if (!t) return;
        AA_CUDA(cub::DeviceScan::ExclusiveSum(t, tmp, it, out, n, stream));
        n_launch++;
        mark("scan");
    }
    // segmented scans (segments = runs of equal keys): the parts phase
    void seg_excl_max_i64(const int32_t *keys, const int64_t *in, int64_t *out, int64_t n) {
        if (n <= 0 || failed) return;
        size_t tmp = 0;
        AA_CUDA(cub::DeviceScan::ExclusiveScanByKey(nullptr, tmp, keys, in, out, MaxI64(), (int64_t)-1, n, cub::Equality(), stream));
        void *t = alloc_bytes(tmp);
        if (!t) return;
        AA_CUDA(cub::DeviceScan::ExclusiveScanByKey(t, tmp, keys, in, out, MaxI64(), (int64_t)-1, n, cub::Equality(), stream));
        n_launch++;
        mark("seg_scan");
    }
    void seg_incl_max_i32(const int32_t *keys, const int32_t *in, int32_t *out, int64_t n) {
        if (n <= 0 || failed) return;
        size_t tmp = 0;
        AA_CUDA(cub::DeviceScan::InclusiveScanByKey(nullptr, tmp, keys, in, out, MaxI32(), n, cub::Equality(), stream));
        void *t = alloc_bytes(tmp);
        if (!t) return;
        AA_CUDA(cub::DeviceScan::InclusiveScanByKey(t, tmp, keys, in, out, MaxI32(), n, cub::Equality(), stream));
        n_launch++;
        mark("seg_scan");
    }
    void seg_rexcl_min_i32(const int32_t *keys, const int32_t *in, int32_t *out, int64_t n) {  // from the right
        if (n <= 0 || failed) return;
        auto rk = thrust::make_reverse_iterator(keys + n);
        auto ri = thrust::make_reverse_iterator(in + n);
        auto ro = thrust::make_reverse_iterator(out + n);
        size_t tmp = 0;
        AA_CUDA(cub::DeviceScan::ExclusiveScanByKey(nullptr, tmp, rk, ri, ro, MinI32(), (int32_t)0x7fffffff, n, cub::Equality(), stream));
        void *t = alloc_bytes(tmp);
        if (!t) return;
        AA_CUDA(cub::DeviceScan::ExclusiveScanByKey(t, tmp, rk, ri, ro, MinI32(), (int32_t)0x7fffffff, n, cub::Equality(), stream));
        n_launch++;
        mark("seg_scan");
    }
    void sort_pairs_u32(const uint32_t *kin, uint32_t *kout, const uint32_t *vin, uint32_t *vout, int64_t n, int end_bit) {
        if (n <= 0 || failed) return;
        size_t tmp = 0;
        AA_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp, kin, kout, vin, vout, n, 0, end_bit, stream));
        void *t = alloc_bytes(tmp);
        if (!t) return;
        AA_CUDA(cub::DeviceRadixSort::SortPairs(t, tmp, kin, kout, vin, vout, n, 0, end_bit, stream));
        n_launch++;
        mark("radix_sort");
    }
    void sort_pairs_u64(const uint64_t *kin, uint64_t *kout, const uint32_t *vin, uint32_t *vout, int64_t n, int end_bit) {
        if (n <= 0 || failed) return;
        size_t tmp = 0;
        AA_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp, kin, kout, vin, vout, n, 0, end_bit, stream));
        void *t = alloc_bytes(tmp);
        if (!t) return;
        AA_CUDA(cub::DeviceRadixSort::SortPairs(t, tmp, kin, kout, vin, vout, n, 0, end_bit, stream));
        n_launch++;
        mark("radix_sort64");
    }
    void fill_ff(void *p, size_t n) {
        if (n && !failed) AA_CUDA(cudaMemsetAsync(p, 0xff, n, stream));
    }
    // ---- side stream: work that is off the critical path runs concurrently with the main stream ----
    bool device_kahn() const { return true; }
    // a second, short-lived fork: the launches between aux_begin and aux_end run concurrently with what follows on the main
    // stream until aux_end's matching join (used to build the heaps of the large and the small contigs side by side)
    void aux_begin() {
        AA_CUDA(cudaEventRecord(ev_aux0, main_stream));
        AA_CUDA(cudaStreamWaitEvent(aux_stream, ev_aux0, 0));
        stream = aux_stream;
    }
    void aux_enter() { stream = aux_stream; }  // back onto the aux stream after an aux_end (no new dependency on the main stream)
    void aux_end() {
        AA_CUDA(cudaEventRecord(ev_aux1, aux_stream));
        stream = main_stream;
    }
    void aux_join() { AA_CUDA(cudaStreamWaitEvent(main_stream, ev_aux1, 0)); }
    void side_begin() {
        AA_CUDA(cudaEventRecord(ev_fork, main_stream));
        AA_CUDA(cudaStreamWaitEvent(side_stream, ev_fork, 0));
        stream = side_stream;
    }
    void side_end() {
        AA_CUDA(cudaEventRecord(ev_side, side_stream));
        stream = main_stream;
        side_pending = true;
    }
    void side_join() {
        if (side_pending) AA_CUDA(cudaStreamWaitEvent(main_stream, ev_side, 0));
        side_pending = false;
    }
    int host_threads() {
        unsigned h = std::thread::hardware_concurrency();
        return (int)std::min<unsigned>(h ? h : 1, 32);
    }
    int64_t max_workers() { return (int64_t)sm_count * 32; }
    // memory the walk scratch may take: a quarter of what was free when the context was created, minus nothing
    // else -- no driver query on the solve path (cudaMemGetInfo stalls the host for milliseconds)
    int64_t scratch_budget() { return (int64_t)std::min<size_t>(free_at_init / 4, (size_t)24 << 30); }
};

}  // namespace aa
