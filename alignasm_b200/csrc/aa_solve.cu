// aa_solve.cu — C ABI of the hot path (include/alignasm_b200.h) on top of the CUDA backend.
// There is no CPU path in this library: without a usable CUDA device every entry point fails.
#include "aa_backend_cuda.cuh"

#include <new>

struct aa_ctx {
    aa::CudaBackend bk;
    aa::Pipeline<aa::CudaBackend> pipe;
    std::string err;
    aa_ctx() : pipe(bk) {}
};
struct aa_dev_batch {
    aa::DevBatch *d;
};

namespace {
thread_local std::string g_create_err = "";
}

extern "C" {

const char *aa_version(void) { return "alignasm_b200 0.1.0 (sm_100a)"; }
const char *aa_phase_name(int phase) { return aa::phase_name(phase); }

// results of this library may live in pinned slabs (aa_backend_cuda.cuh): aa_result_free gives them back
static const bool g_slab_hook = (aa::g_result_slab_release = aa::result_slab_release, true);

aa_status aa_create(aa_ctx **ctx, int device) {
    if (!ctx) return AA_ERR_INVALID;
    *ctx = nullptr;
    aa_ctx *c = new (std::nothrow) aa_ctx();
    if (!c) return AA_ERR_NOMEM;
    if (!c->bk.init(device)) {
        g_create_err = c->bk.error();
        delete c;
        return AA_ERR_NO_DEVICE;
    }
    *ctx = c;
    return AA_OK;
}
void aa_destroy(aa_ctx *ctx) {
    if (!ctx) return;
    ctx->bk.shutdown();
    delete ctx;
}
int aa_ctx_device(const aa_ctx *ctx) { return ctx ? ctx->bk.device : -1; }

void *aa_host_alloc(int64_t bytes) {
    void *p = nullptr;
    if (bytes <= 0) bytes = 1;
    if (cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void aa_host_free(void *p) {
    if (p && cudaFreeHost(p) != cudaSuccess) cudaGetLastError();
}
const char *aa_last_error(const aa_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

aa_status aa_upload(aa_ctx *ctx, const aa_batch *batch, aa_dev_batch **dev) {
    if (!ctx || !dev) return AA_ERR_INVALID;
    *dev = nullptr;
    if (!ctx->bk.ok() && ctx->bk.oom) ctx->bk.reset_pool();  // an allocation failure is recoverable: start from an empty pool
    if (!ctx->bk.ok()) {
        ctx->err = ctx->bk.error();
        return AA_ERR_CUDA;
    }
    cudaSetDevice(ctx->bk.device);
    aa::DevBatch *d = nullptr;
    aa_status st = ctx->pipe.upload(batch, d);
    if (st != AA_OK) {
        ctx->err = ctx->pipe.err;
        return st;
    }
    if (!ctx->bk.ok()) {
        ctx->err = ctx->bk.error();
        ctx->pipe.free_batch(d);
        return AA_ERR_CUDA;
    }
    *dev = new aa_dev_batch{d};
    return AA_OK;
}
void aa_dev_batch_free(aa_ctx *ctx, aa_dev_batch *dev) {
    if (!ctx || !dev) return;
    ctx->pipe.free_batch(dev->d);
    delete dev;
}
aa_status aa_solve_device(aa_ctx *ctx, aa_dev_batch *dev, const aa_opts *opts, aa_result *res) {
    if (!ctx || !dev || !dev->d) return AA_ERR_INVALID;
    if (!ctx->bk.ok() && !ctx->bk.oom) {  // (an out-of-memory failure is cleared by the pool reset of the next solve)
        ctx->err = ctx->bk.error();
        return AA_ERR_CUDA;
    }
    aa_opts o{};
    if (opts) o = *opts;
    aa_status st = ctx->pipe.solve(*dev->d, o, res);
    if (st != AA_OK) {
        ctx->err = ctx->pipe.err;
        if (st != AA_ERR_UNSOLVABLE) {
            if (res) aa::result_free_host(res);
            ctx->bk.quiesce();  // side-stream kernels (topo, walk-0 trace) may still be running
        }
    }
    return st;
}
aa_status aa_solve(aa_ctx *ctx, const aa_batch *batch, const aa_opts *opts, aa_result *res) {
    if (!ctx || !res) return AA_ERR_INVALID;
    // one-shot: the batch is staged in the pooled workspace, so a steady-state call makes no cudaMalloc / cudaFree
    ctx->bk.reset_pool();  // (also recovers a context whose last solve ran out of memory)
    if (!ctx->bk.ok()) {
        ctx->err = ctx->bk.error();
        return AA_ERR_CUDA;
    }
    aa::DevBatch *d = nullptr;
    aa_status st = ctx->pipe.upload(batch, d, /*pooled=*/true);
    if (st != AA_OK) {
        ctx->err = ctx->pipe.err;
        return st;
    }
    if (!ctx->bk.ok()) {
        ctx->err = ctx->bk.error();
        ctx->pipe.free_batch(d);
        return AA_ERR_CUDA;
    }
    aa_opts o{};
    if (opts) o = *opts;
    st = ctx->pipe.solve(*d, o, res, /*keep_pool=*/true);
    if (st != AA_OK) {
        ctx->err = ctx->pipe.err;
        if (st != AA_ERR_UNSOLVABLE) {
            aa::result_free_host(res);
            ctx->bk.quiesce();  // side-stream kernels (topo, walk-0 trace) may still be running
        }
    }
    ctx->pipe.free_batch(d);
    return st;
}

aa_status aa_solve_subset(aa_ctx *ctx, const aa_batch *batch, const int64_t *ctgs, int64_t n_ctgs, const aa_opts *opts,
                          aa_result *res) {
    if (!ctx || !res) return AA_ERR_INVALID;
    // one-shot: the batch is staged in the pooled workspace, so a steady-state call makes no cudaMalloc / cudaFree
    ctx->bk.reset_pool();  // (also recovers a context whose last solve ran out of memory)
    if (!ctx->bk.ok()) {
        ctx->err = ctx->bk.error();
        return AA_ERR_CUDA;
    }
    aa::DevBatch *d = nullptr;
    aa_status st = ctx->pipe.upload_subset(batch, ctgs, n_ctgs, d);
    if (st != AA_OK) {
        ctx->err = ctx->pipe.err;
        return st;
    }
    if (!ctx->bk.ok()) {
        ctx->err = ctx->bk.error();
        ctx->pipe.free_batch(d);
        return AA_ERR_CUDA;
    }
    aa_opts o{};
    if (opts) o = *opts;
    st = ctx->pipe.solve(*d, o, res, /*keep_pool=*/true);
    if (st != AA_OK) {
        ctx->err = ctx->pipe.err;
        if (st != AA_ERR_UNSOLVABLE) {
            aa::result_free_host(res);
            ctx->bk.quiesce();  // side-stream kernels (topo, walk-0 trace) may still be running
        }
    }
    ctx->pipe.free_batch(d);
    return st;
}
void aa_result_free(aa_result *res) { aa::result_free_host(res); }
aa_status aa_get_stats(const aa_ctx *ctx, aa_stats *stats) {
    if (!ctx || !stats) return AA_ERR_INVALID;
    *stats = ctx->pipe.last_stats;
    return AA_OK;
}

}  // extern "C"
