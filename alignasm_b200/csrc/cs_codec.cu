// cs_codec.cu — the cs:Z: codec on the device.
//
//   aa_cs_runs_device   parse_short_cs + get_overlap_range (reference src/paf_data.cpp:29-72, 90-123): the exact-match runs of
//                       every alignment row, in query orientation, with the reference's consistency check
//   aa_cs_edit_device   get_edited_paf_data (reference src/paf_data.cpp:125-220): the cs:Z: field, mat_num and aln_len of
//                       every OUTPUT row, re-cut to the edited query interval
//
// Both walk a row's cs string ONCE, forward, whatever the strand.  The reference applies the operations of a '-' row in
// reverse order; the query position at which operation k is applied then is qs + (Q - Qincl(k)) (Q = total query consumption,
// Qincl = inclusive prefix), so a forward walk with running prefixes gives the same positions, and the runs of a '-' row are
// written back to front.  The re-cut output is in FILE orientation, i.e. forward token order for both strands.
// One thread per row: a cs field is ~100 bytes (a few thousand for long alignments), the rows are independent, and the
// whole file image is uploaded once, so the kernels read it where it lies.  Two passes each (count / size, scan, fill).
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/alignasm_b200.h"



namespace {

struct Tok {
    char type;
    int64_t len;     // ':' match length, '+' / '-' indel length, '*' 1
    int32_t at, n;   // slice of the cs string holding the operation's text
};
__device__ __forceinline__ bool is_alpha_d(char c) { return (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z'); }
// one operation of a short-form cs string (parse_short_cs, paf_data.cpp:29-72); 0 or the error code of the reference's throw site
__device__ __forceinline__ int32_t next_tok(const char *__restrict__ s, int32_t n, int32_t &pos, Tok &o) {
    const int32_t start = pos;
    const char t = s[pos++];
    int64_t len = 0;
    if (t == ':') {
        bool neg = false;
        if (pos < n && s[pos] == '-') {  // std::from_chars accepts a sign; the value must be > 0 anyway
            neg = true;
            pos++;
        }
        const int32_t d0 = pos;
        while (pos < n && s[pos] >= '0' && s[pos] <= '9') {
            len = len * 10 + (s[pos] - '0');
            pos++;
        }
        if (pos == d0 || neg || len <= 0) return AA_CS_ERR_LENGTH;
    } else if (t == '*') {
        if (pos + 2 > n || !is_alpha_d(s[pos]) || !is_alpha_d(s[pos + 1])) return AA_CS_ERR_SUBST;
        pos += 2;
        len = 1;
    } else if (t == '+' || t == '-') {
        const int32_t s0 = pos;
        while (pos < n && is_alpha_d(s[pos])) pos++;
        len = pos - s0;
        if (len == 0) return AA_CS_ERR_INDEL;
    } else {
        return AA_CS_ERR_OP;
    }
    o.type = t;
    o.len = len;
    o.at = start;
    o.n = pos - start;
    return 0;
}
__device__ __forceinline__ bool has_prefix(const char *s, int32_t n) {
    return n >= 5 && s[0] == 'c' && s[1] == 's' && s[2] == ':' && s[3] == 'Z' && s[4] == ':';
}

// ---- get_overlap_range -------------------------------------------------------------------------------------------
__global__ void k_cs_count(int64_t n, const char *__restrict__ text, const int64_t *__restrict__ off, const int32_t *__restrict__ len,
                           const int64_t *__restrict__ qs, const int64_t *__restrict__ qe, const int64_t *__restrict__ rs,
                           const int64_t *__restrict__ re, const uint8_t *__restrict__ fwd, int32_t *__restrict__ nrun,
                           int32_t *__restrict__ err) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const char *s = text + off[i];
    const int32_t m = len[i];
    int32_t e = 0, runs = 0;
    int64_t Q = 0, R = 0;
    if (!has_prefix(s, m)) {
        e = AA_CS_ERR_TAG;
    } else {
        int32_t pos = 5;
        Tok o;
        while (pos < m && (e = next_tok(s, m, pos, o)) == 0) {
            if (o.type == ':') {
                runs++;
                Q += o.len;
                R += o.len;
            } else if (o.type == '+') {
                Q += o.len;
            } else if (o.type == '-') {
                R += o.len;
            } else {
                Q += 1;
                R += 1;
            }
        }
        if (e == 0) {  // consumption must match the coordinates (paf_data.cpp:119-122)
            const int64_t step = fwd[i] ? 1 : -1;
            if (qs[i] + Q != qe[i] + 1 || rs[i] + step * R != re[i] + step) e = AA_CS_ERR_CONSUME;
        }
    }
    err[i] = e;
    nrun[i] = e ? 0 : runs;
}
__global__ void k_cs_fill(int64_t n, const char *__restrict__ text, const int64_t *__restrict__ off, const int32_t *__restrict__ len,
                          const int64_t *__restrict__ qs, const int64_t *__restrict__ qe, const int64_t *__restrict__ rs,
                          const int64_t *__restrict__ re, const uint8_t *__restrict__ fwd, const int32_t *__restrict__ err,
                          const int64_t *__restrict__ run_off, int64_t *__restrict__ run_ql, int64_t *__restrict__ run_qr,
                          int64_t *__restrict__ run_rl) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || err[i]) return;
    const char *s = text + off[i];
    const int32_t m = len[i];
    const bool f = fwd[i] != 0;
    const int64_t r0 = run_off[i], nr = run_off[i + 1] - r0;
    const int64_t q0 = qs[i], Qtot = qe[i] + 1 - q0, ref0 = rs[i], Rtot = f ? re[i] + 1 - ref0 : ref0 - re[i] + 1;
    int64_t Q = 0, R = 0, k = 0;  // exclusive prefixes of the consumption, run counter
    int32_t pos = 5;
    Tok o;
    while (pos < m && next_tok(s, m, pos, o) == 0) {
        const int64_t ql = (o.type == ':' || o.type == '+') ? o.len : (o.type == '*' ? 1 : 0);
        const int64_t rl = (o.type == ':' || o.type == '-') ? o.len : (o.type == '*' ? 1 : 0);
        if (o.type == ':') {
            // position at which the reference applies this operation: forward rows in order, '-' rows from the last one back
            const int64_t qi = f ? q0 + Q : q0 + (Qtot - (Q + ql));
            const int64_t ri = f ? ref0 + R : ref0 - (Rtot - (R + rl));
            const int64_t at = r0 + (f ? k : nr - 1 - k);
            run_ql[at] = qi;
            run_qr[at] = qi + o.len - 1;
            run_rl[at] = ri;
            k++;
        }
        Q += ql;
        R += rl;
    }
}

// ---- get_edited_paf_data ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int32_t dec_digits(int64_t v) {
    int32_t d = 1;
    while (v >= 10) {
        v /= 10;
        d++;
    }
    return d;
}
template <bool FILL>
__global__ void k_cs_edit(int64_t n_out, const char *__restrict__ text, const int64_t *__restrict__ off, const int32_t *__restrict__ len,
                          const int64_t *__restrict__ qs, const int64_t *__restrict__ qe, const uint8_t *__restrict__ fwd,
                          const int32_t *__restrict__ row_mat, const int32_t *__restrict__ row_aln, const int64_t *__restrict__ out_row,
                          const int64_t *__restrict__ eqs_a, const int64_t *__restrict__ eqe_a, const int64_t *__restrict__ ers_a,
                          const int64_t *__restrict__ ere_a, int32_t *__restrict__ out_len, const int64_t *__restrict__ out_off,
                          char *__restrict__ out_text, int32_t *__restrict__ out_mat, int32_t *__restrict__ out_aln,
                          int32_t *__restrict__ out_err) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_out) return;
    const int64_t g = out_row[k];
    const char *s = text + off[g];
    const int32_t m = len[g];
    const int64_t eqs = eqs_a[k], eqe = eqe_a[k];
    char *dst = FILL ? out_text + out_off[k] : nullptr;
    if (FILL && out_err[k]) return;
    if (eqs == qs[g] && eqe == qe[g]) {  // untouched row: the field as it was read (paf_data.cpp:128-135)
        if (FILL) {
            for (int32_t j = 0; j < m; j++) dst[j] = s[j];
        } else {
            out_len[k] = m;
            out_mat[k] = row_mat[g];
            out_aln[k] = row_aln[g];
            out_err[k] = 0;
        }
        return;
    }
    int32_t e = 0, w = 5, mat = 0, aln = 0;
    int64_t qb = 0, rb = 0;
    if (!has_prefix(s, m)) e = AA_CS_ERR_TAG;
    if (FILL) {
        dst[0] = 'c';
        dst[1] = 's';
        dst[2] = ':';
        dst[3] = 'Z';
        dst[4] = ':';
    }
    const bool f = fwd[g] != 0;
    const int64_t q0 = qs[g], Qtot = qe[g] + 1 - q0;
    int64_t Q = 0;
    int32_t pos = 5;
    Tok o;
    while (e == 0 && pos < m && (e = next_tok(s, m, pos, o)) == 0) {
        const int64_t ql = (o.type == ':' || o.type == '+') ? o.len : (o.type == '*' ? 1 : 0);
        const int64_t qi = f ? q0 + Q : q0 + (Qtot - (Q + ql));  // query position at which the reference applies the operation
        Q += ql;
        if (o.type == ':') {
            const int64_t oe = qi + o.len - 1;
            const int64_t a = qi > eqs ? qi : eqs, b = oe < eqe ? oe : eqe;
            if (a <= b) {
                const int64_t l = b - a + 1;
                const int32_t d = dec_digits(l);
                if (FILL) {
                    dst[w] = ':';
                    int64_t v = l;
                    for (int32_t j = d; j >= 1; j--) {
                        dst[w + j] = (char)('0' + v % 10);
                        v /= 10;
                    }
                }
                w += 1 + d;
                mat += (int32_t)l;
                aln += (int32_t)l;
                qb += l;
                rb += l;
            }
        } else {
            bool keep;
            if (o.type == '+') {
                const int64_t oe = qi + o.len - 1;
                keep = qi <= eqe && eqs <= oe;
                if (keep && (qi < eqs || eqe < oe)) e = AA_CS_ERR_CLIP_INS;
            } else if (o.type == '*') {
                keep = eqs <= qi && qi <= eqe;
            } else {
                keep = eqs < qi && qi <= eqe;
            }
            if (keep && e == 0) {
                if (FILL)
                    for (int32_t j = 0; j < o.n; j++) dst[w + j] = s[o.at + j];
                w += o.n;
                aln += (int32_t)o.len;
                if (o.type == '+') qb += o.len;
                else if (o.type == '-') rb += o.len;
                else {
                    qb += 1;
                    rb += 1;
                }
            }
        }
    }
    if (!FILL) {
        const int64_t ers = ers_a[k], ere = ere_a[k];
        const int64_t want_r = ere > ers ? ere - ers : ers - ere;
        if (e == 0 && (qb != eqe - eqs + 1 || rb != want_r + 1)) e = AA_CS_ERR_EDIT;
        out_len[k] = e ? 0 : w;
        out_mat[k] = mat;
        out_aln[k] = aln;
        out_err[k] = e;
    }
}

// Device scratch of one call: ONE cudaMalloc per stage (cudaFree synchronises the device and costs a millisecond a piece, and
// this path is about latency), carved into 256-byte aligned pieces.  This is not the solve path: no pool.
struct DevArena {
    char *base = nullptr;
    size_t cap = 0, top = 0;
    bool reserve(size_t bytes) {
        cap = bytes + 4096;
        return cudaMalloc((void **)&base, cap) == cudaSuccess;
    }
    void *take(size_t n) {
        n = (n + 255) & ~(size_t)255;
        if (top + n > cap) return nullptr;
        void *p = base + top;
        top += n;
        return p;
    }
    ~DevArena() {
        if (base) cudaFree(base);
    }
};
struct DevBuf {
    void *p = nullptr;
    bool alloc(DevArena &a, size_t n) {
        p = a.take(n ? n : 16);
        return p != nullptr;
    }
    template <class T>
    T *as() const {
        return (T *)p;
    }
};
inline size_t pad256(size_t n) { return ((n ? n : 16) + 255) & ~(size_t)255; }
thread_local std::string g_cs_err;
aa_status cs_fail(const char *what, cudaError_t e) {
    g_cs_err = std::string(what) + ": " + cudaGetErrorString(e);
    cudaGetLastError();
    return e == cudaErrorMemoryAllocation ? AA_ERR_NOMEM : AA_ERR_CUDA;
}
#define CSCK(call)                                         \
    do {                                                   \
        cudaError_t e__ = (call);                          \
        if (e__ != cudaSuccess) return cs_fail(#call, e__); \
    } while (0)
template <class T>
bool up(DevArena &a, DevBuf &b, const T *h, int64_t n, cudaStream_t st) {
    if (!b.alloc(a, (size_t)(n > 0 ? n : 1) * sizeof(T))) return false;
    return n <= 0 || cudaMemcpyAsync(b.p, h, (size_t)n * sizeof(T), cudaMemcpyHostToDevice, st) == cudaSuccess;
}
struct CastI64 {
    __host__ __device__ __forceinline__ int64_t operator()(const int32_t &x) const { return (int64_t)x; }
};
size_t scan_tmp_bytes(int64_t n) {
    cub::TransformInputIterator<int64_t, CastI64, const int32_t *> it(nullptr, CastI64());
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, it, (int64_t *)nullptr, n + 1, nullptr);
    return tmp;
}
aa_status scan_i32(DevArena &a, const int32_t *in, int64_t *out, int64_t n, cudaStream_t st) {  // out[0..n]: exclusive prefix sums + total
    cub::TransformInputIterator<int64_t, CastI64, const int32_t *> it(in, CastI64());
    size_t tmp = 0;
    CSCK(cub::DeviceScan::ExclusiveSum(nullptr, tmp, it, out, n + 1, st));
    DevBuf t;
    if (!t.alloc(a, tmp)) return cs_fail("cudaMalloc(scan)", cudaErrorMemoryAllocation);
    CSCK(cub::DeviceScan::ExclusiveSum(t.p, tmp, it, out, n + 1, st));
    return AA_OK;
}
}  // namespace

extern "C" {

const char *aa_cs_last_error(void) { return g_cs_err.c_str(); }

const char *aa_cs_error_text(int32_t code) {  // the reference's exception texts (paf_data.cpp:31-122, 140-218)
    switch (code) {
        case 0: return "";
        case AA_CS_ERR_TAG: return "PAF record does not contain a short-form cs:Z tag";
        case AA_CS_ERR_LENGTH: return "Invalid :length operation in cs tag";
        case AA_CS_ERR_SUBST: return "Invalid substitution operation in cs tag";
        case AA_CS_ERR_INDEL: return "Empty indel operation in cs tag";
        case AA_CS_ERR_OP: return "Unsupported operation in short-form cs tag";
        case AA_CS_ERR_CONSUME: return "cs tag consumption does not match PAF coordinates";
        case AA_CS_ERR_CLIP_INS: return "Alignment was clipped inside a cs insertion";
        case AA_CS_ERR_EDIT: return "Edited cs tag does not match edited PAF coordinates";
    }
    return "unknown cs error";
}

aa_status aa_cs_runs_device(aa_ctx *ctx, const char *text, int64_t text_len, const aa_cs_rows *rows, aa_cs_runs *out) {
    if (!ctx || !text || !rows || !out || rows->n < 0) return AA_ERR_INVALID;
    std::memset(out, 0, sizeof *out);
    const int64_t n = rows->n;
    CSCK(cudaSetDevice(aa_ctx_device(ctx)));
    cudaStream_t st = nullptr;  // the legacy default stream: this call is synchronous anyway
    DevBuf d_text, d_off, d_len, d_qs, d_qe, d_rs, d_re, d_fwd, d_nrun, d_err, d_roff, d_ql, d_qr, d_rl;
    DevArena A, A2;
    const size_t N1 = (size_t)(n + 2);
    if (!A.reserve(pad256((size_t)text_len) + 6 * pad256(N1 * 8) + 3 * pad256(N1 * 4) + pad256(N1) + pad256(scan_tmp_bytes(n)) + 16 * 256))
        return cs_fail("cudaMalloc(cs fields)", cudaErrorMemoryAllocation);
    if (!up(A, d_text, text, text_len, st) || !up(A, d_off, rows->cs_off, n, st) || !up(A, d_len, rows->cs_len, n, st) ||
        !up(A, d_qs, rows->qry_str, n, st) || !up(A, d_qe, rows->qry_end, n, st) || !up(A, d_rs, rows->ref_str, n, st) ||
        !up(A, d_re, rows->ref_end, n, st) || !up(A, d_fwd, rows->aln_fwd, n, st) || !d_nrun.alloc(A, (size_t)(n + 1) * 4) ||
        !d_err.alloc(A, (size_t)(n + 1) * 4) || !d_roff.alloc(A, (size_t)(n + 2) * 8))
        return cs_fail("staging the cs fields", cudaGetLastError() == cudaSuccess ? cudaErrorMemoryAllocation : cudaErrorUnknown);
    CSCK(cudaMemsetAsync(d_nrun.p, 0, (size_t)(n + 1) * 4, st));
    const unsigned grid = (unsigned)((n + 127) / 128);
    if (n > 0)
        k_cs_count<<<grid, 128, 0, st>>>(n, d_text.as<char>(), d_off.as<int64_t>(), d_len.as<int32_t>(), d_qs.as<int64_t>(), d_qe.as<int64_t>(),
                                         d_rs.as<int64_t>(), d_re.as<int64_t>(), d_fwd.as<uint8_t>(), d_nrun.as<int32_t>(), d_err.as<int32_t>());
    aa_status s = scan_i32(A, d_nrun.as<int32_t>(), d_roff.as<int64_t>(), n, st);
    if (s != AA_OK) return s;
    int64_t R = 0;
    CSCK(cudaMemcpy(&R, d_roff.as<int64_t>() + n, 8, cudaMemcpyDeviceToHost));
    if (!A2.reserve(3 * pad256((size_t)R * 8)) || !d_ql.alloc(A2, (size_t)R * 8) || !d_qr.alloc(A2, (size_t)R * 8) || !d_rl.alloc(A2, (size_t)R * 8))
        return cs_fail("cudaMalloc(runs)", cudaErrorMemoryAllocation);
    if (n > 0)
        k_cs_fill<<<grid, 128, 0, st>>>(n, d_text.as<char>(), d_off.as<int64_t>(), d_len.as<int32_t>(), d_qs.as<int64_t>(), d_qe.as<int64_t>(),
                                        d_rs.as<int64_t>(), d_re.as<int64_t>(), d_fwd.as<uint8_t>(), d_err.as<int32_t>(), d_roff.as<int64_t>(),
                                        d_ql.as<int64_t>(), d_qr.as<int64_t>(), d_rl.as<int64_t>());
    CSCK(cudaGetLastError());
    out->n_rows = n;
    out->n_run = R;
    out->run_off = (int64_t *)std::malloc((size_t)(n + 1) * 8);
    out->run_ql = (int64_t *)std::malloc((size_t)(R > 0 ? R : 1) * 8);
    out->run_qr = (int64_t *)std::malloc((size_t)(R > 0 ? R : 1) * 8);
    out->run_rl = (int64_t *)std::malloc((size_t)(R > 0 ? R : 1) * 8);
    out->err = (int32_t *)std::malloc((size_t)(n > 0 ? n : 1) * 4);
    if (!out->run_off || !out->run_ql || !out->run_qr || !out->run_rl || !out->err) {
        aa_cs_runs_free(out);
        g_cs_err = "out of host memory";
        return AA_ERR_NOMEM;
    }
    CSCK(cudaMemcpy(out->run_off, d_roff.p, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost));
    if (R > 0) {
        CSCK(cudaMemcpy(out->run_ql, d_ql.p, (size_t)R * 8, cudaMemcpyDeviceToHost));
        CSCK(cudaMemcpy(out->run_qr, d_qr.p, (size_t)R * 8, cudaMemcpyDeviceToHost));
        CSCK(cudaMemcpy(out->run_rl, d_rl.p, (size_t)R * 8, cudaMemcpyDeviceToHost));
    }
    if (n > 0) CSCK(cudaMemcpy(out->err, d_err.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
    return AA_OK;
}
void aa_cs_runs_free(aa_cs_runs *r) {
    if (!r) return;
    std::free(r->run_off);
    std::free(r->run_ql);
    std::free(r->run_qr);
    std::free(r->run_rl);
    std::free(r->err);
    std::memset(r, 0, sizeof *r);
}

aa_status aa_cs_edit_device(aa_ctx *ctx, const char *text, int64_t text_len, const aa_cs_rows *rows, const int32_t *row_mat,
                            const int32_t *row_aln, int64_t n_out, const int64_t *out_row, const int64_t *eqs, const int64_t *eqe,
                            const int64_t *ers, const int64_t *ere, aa_cs_edits *out) {
    if (!ctx || !text || !rows || !out || n_out < 0 || (n_out > 0 && (!out_row || !eqs || !eqe || !ers || !ere)) || !row_mat || !row_aln)
        return AA_ERR_INVALID;
    std::memset(out, 0, sizeof *out);
    const int64_t n = rows->n;
    CSCK(cudaSetDevice(aa_ctx_device(ctx)));
    cudaStream_t st = nullptr;
    DevBuf d_text, d_off, d_len, d_qs, d_qe, d_fwd, d_mat, d_aln, d_row, d_eqs, d_eqe, d_ers, d_ere, d_olen, d_ooff, d_omat, d_oaln, d_oerr, d_otext;
    DevArena A, A2;
    const size_t N1 = (size_t)(n + 2), M1 = (size_t)(n_out + 2);
    if (!A.reserve(pad256((size_t)text_len) + 3 * pad256(N1 * 8) + 3 * pad256(N1 * 4) + pad256(N1) + 6 * pad256(M1 * 8) + 4 * pad256(M1 * 4) +
                   pad256(scan_tmp_bytes(n_out)) + 24 * 256))
        return cs_fail("cudaMalloc(cs fields)", cudaErrorMemoryAllocation);
    if (!up(A, d_text, text, text_len, st) || !up(A, d_off, rows->cs_off, n, st) || !up(A, d_len, rows->cs_len, n, st) ||
        !up(A, d_qs, rows->qry_str, n, st) || !up(A, d_qe, rows->qry_end, n, st) || !up(A, d_fwd, rows->aln_fwd, n, st) ||
        !up(A, d_mat, row_mat, n, st) || !up(A, d_aln, row_aln, n, st) || !up(A, d_row, out_row, n_out, st) || !up(A, d_eqs, eqs, n_out, st) ||
        !up(A, d_eqe, eqe, n_out, st) || !up(A, d_ers, ers, n_out, st) || !up(A, d_ere, ere, n_out, st) || !d_olen.alloc(A, (size_t)(n_out + 1) * 4) ||
        !d_ooff.alloc(A, (size_t)(n_out + 2) * 8) || !d_omat.alloc(A, (size_t)(n_out + 1) * 4) || !d_oaln.alloc(A, (size_t)(n_out + 1) * 4) ||
        !d_oerr.alloc(A, (size_t)(n_out + 1) * 4))
        return cs_fail("staging the cs fields", cudaGetLastError() == cudaSuccess ? cudaErrorMemoryAllocation : cudaErrorUnknown);
    CSCK(cudaMemsetAsync(d_olen.p, 0, (size_t)(n_out + 1) * 4, st));
    const unsigned grid = (unsigned)((n_out + 127) / 128);
    if (n_out > 0)
        k_cs_edit<false><<<grid, 128, 0, st>>>(n_out, d_text.as<char>(), d_off.as<int64_t>(), d_len.as<int32_t>(), d_qs.as<int64_t>(), d_qe.as<int64_t>(),
                                               d_fwd.as<uint8_t>(), d_mat.as<int32_t>(), d_aln.as<int32_t>(), d_row.as<int64_t>(), d_eqs.as<int64_t>(),
                                               d_eqe.as<int64_t>(), d_ers.as<int64_t>(), d_ere.as<int64_t>(), d_olen.as<int32_t>(), nullptr, nullptr,
                                               d_omat.as<int32_t>(), d_oaln.as<int32_t>(), d_oerr.as<int32_t>());
    aa_status s = scan_i32(A, d_olen.as<int32_t>(), d_ooff.as<int64_t>(), n_out, st);
    if (s != AA_OK) return s;
    int64_t bytes = 0;
    CSCK(cudaMemcpy(&bytes, d_ooff.as<int64_t>() + n_out, 8, cudaMemcpyDeviceToHost));
    if (!A2.reserve(pad256((size_t)bytes)) || !d_otext.alloc(A2, (size_t)bytes)) return cs_fail("cudaMalloc(edited cs)", cudaErrorMemoryAllocation);
    if (n_out > 0)
        k_cs_edit<true><<<grid, 128, 0, st>>>(n_out, d_text.as<char>(), d_off.as<int64_t>(), d_len.as<int32_t>(), d_qs.as<int64_t>(), d_qe.as<int64_t>(),
                                              d_fwd.as<uint8_t>(), d_mat.as<int32_t>(), d_aln.as<int32_t>(), d_row.as<int64_t>(), d_eqs.as<int64_t>(),
                                              d_eqe.as<int64_t>(), d_ers.as<int64_t>(), d_ere.as<int64_t>(), d_olen.as<int32_t>(), d_ooff.as<int64_t>(),
                                              d_otext.as<char>(), d_omat.as<int32_t>(), d_oaln.as<int32_t>(), d_oerr.as<int32_t>());
    CSCK(cudaGetLastError());
    out->n = n_out;
    out->n_bytes = bytes;
    out->off = (int64_t *)std::malloc((size_t)(n_out + 1) * 8);
    out->text = (char *)std::malloc((size_t)(bytes > 0 ? bytes : 1));
    out->mat_num = (int32_t *)std::malloc((size_t)(n_out > 0 ? n_out : 1) * 4);
    out->aln_len = (int32_t *)std::malloc((size_t)(n_out > 0 ? n_out : 1) * 4);
    out->err = (int32_t *)std::malloc((size_t)(n_out > 0 ? n_out : 1) * 4);
    if (!out->off || !out->text || !out->mat_num || !out->aln_len || !out->err) {
        aa_cs_edits_free(out);
        g_cs_err = "out of host memory";
        return AA_ERR_NOMEM;
    }
    CSCK(cudaMemcpy(out->off, d_ooff.p, (size_t)(n_out + 1) * 8, cudaMemcpyDeviceToHost));
    if (bytes > 0) CSCK(cudaMemcpy(out->text, d_otext.p, (size_t)bytes, cudaMemcpyDeviceToHost));
    if (n_out > 0) {
        CSCK(cudaMemcpy(out->mat_num, d_omat.p, (size_t)n_out * 4, cudaMemcpyDeviceToHost));
        CSCK(cudaMemcpy(out->aln_len, d_oaln.p, (size_t)n_out * 4, cudaMemcpyDeviceToHost));
        CSCK(cudaMemcpy(out->err, d_oerr.p, (size_t)n_out * 4, cudaMemcpyDeviceToHost));
    }
    return AA_OK;
}
void aa_cs_edits_free(aa_cs_edits *e) {
    if (!e) return;
    std::free(e->off);
    std::free(e->text);
    std::free(e->mat_num);
    std::free(e->aln_len);
    std::free(e->err);
    std::memset(e, 0, sizeof *e);
}

}  // extern "C"
