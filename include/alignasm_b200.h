/* alignasm_b200.h — C ABI of the B200-native alignasm hot path.
 *
 * This is the drop-in boundary for the one data-parallel path of ACCtools/alignasm:
 *
 *     void solve_ctg_read(std::vector<PafReadData>&, std::vector<PafOutputData>& out,
 *                         std::vector<PafOutputData>& alt_out,
 *                         std::vector<std::vector<PafOutputData>>& max_out);
 *                                  (reference: src/paf_data.hpp:193, called at src/alignasm.cpp:357,373,391)
 *
 * The C++ seam is not ABI-stable (std::vector / std::string), so the C layer is batch-oriented and
 * carries exactly the fields that cross it, as plain pointers and sizes:
 *   in : all blocks of every contig in FILE order (ctg_index = position inside the contig,
 *        alignasm.cpp:138-139), closed int64 intervals (alignasm.cpp:141-151), ref_str > ref_end for
 *        '-' rows (alignasm.cpp:155-159), dense chromosome ids (alignasm.cpp:153), mapq, query length,
 *        and the exact-match runs that get_overlap_range() derives from cs:Z: (paf_data.cpp:90-123).
 *   out: the three per-contig lists of PafOutputData (paf_data.hpp:90-105) — primary chain
 *        (.aln.paf), alternative chain (.aln.alt.paf) and the equal-coverage list (.aln.all.paf) —
 *        plus ctg_sorted_index, the one side effect the reference writes into its input
 *        (paf_data.cpp:236,244).
 *
 * Error behaviour: the reference throws / asserts (never caught); this ABI returns a
 * non-zero aa_status and a message from aa_last_error().  There is NO CPU fallback: every aa_solve*
 * entry point fails with AA_ERR_NO_DEVICE when no CUDA device is usable.
 *
 * Threading: every function that takes an aa_ctx is re-entrant per context (one host thread per context at a time; two
 * contexts never share state).  The library holds ONE piece of process-wide state: the cache of warm contexts behind
 * aa_solve_multi (csrc/aa_multi.cpp).  It is mutex-guarded and contexts are checked out exclusively for the length of a
 * shard solve, so concurrent aa_solve_multi calls are safe (a device whose cached contexts are all in use gets a new
 * one); aa_multi_release() waits for running solves before it frees them.  aa_multi_last_error() is per thread.
 */
#ifndef ALIGNASM_B200_H
#define ALIGNASM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum aa_status {
    AA_OK = 0,
    AA_ERR_INVALID = 1,    /* malformed batch (offsets not monotone, empty contig, ...)            */
    AA_ERR_NO_DEVICE = 2,  /* no usable CUDA device: the product has no CPU path                    */
    AA_ERR_CUDA = 3,       /* CUDA runtime error, message in aa_last_error                          */
    AA_ERR_NOMEM = 4,      /* device or host allocation failed / contig too large for one GPU       */
    AA_ERR_IO = 5,         /* PAF file could not be read / written                                  */
    AA_ERR_FORMAT = 6,     /* PAF / cs:Z: syntax error (reference: std::invalid_argument)           */
    AA_ERR_UNSOLVABLE = 7  /* a contig has no src->dest walk (reference: assert, paf_data.cpp:732)  */
} aa_status;

/* ---- input: one batch of contigs, structure-of-arrays, host memory ------------------------- */
typedef struct aa_batch {
    int64_t n_ctg;            /* number of contigs                                                  */
    int64_t n_blk;            /* total alignment blocks (PAF rows)                                  */
    int64_t n_run;            /* total exact-match runs over all blocks                             */
    const int64_t *ctg_off;   /* [n_ctg+1] block offsets; blocks of a contig are in file order     */
    const int64_t *qry_str;   /* [n_blk] closed interval on the query                               */
    const int64_t *qry_end;
    const int64_t *ref_str;   /* [n_blk] closed interval on the target; ref_str > ref_end on '-'    */
    const int64_t *ref_end;
    const int64_t *qry_total; /* [n_blk] query (contig) length, PAF column 2                        */
    const int32_t *ref_chr;   /* [n_blk] dense target id                                            */
    const uint8_t *aln_fwd;   /* [n_blk] 1 = '+', 0 = '-'                                           */
    const uint8_t *map_qul;   /* [n_blk] mapq                                                       */
    const int64_t *run_off;   /* [n_blk+1] offsets into the run arrays                              */
    const int64_t *run_ql;    /* [n_run] query-closed [l,r] of each exact-match run (':' op)        */
    const int64_t *run_qr;
    const int64_t *run_rl;    /* [n_run] target coordinate of the run's first query base           */
} aa_batch;

typedef struct aa_opts {
    int32_t non_skip_linkable; /* --non_skip_linkable (alignasm.cpp:54-57,74)                        */
    int32_t want_all;          /* materialise the .aln.all.paf list (can be very large)              */
    int32_t max_walks;         /* MAX_PATH_COUNT, paf_data.cpp:729; 0 = reference value 10000         */
    int32_t keep_debug;        /* fill aa_result.dbg (graph / d / best / walk distances)             */
} aa_opts;

/* rows of PafOutputData (paf_data.hpp:90-105), structure-of-arrays */
typedef struct aa_rows {
    int64_t n;
    int32_t *ctg_index;
    int64_t *qry_str, *qry_end, *ref_str, *ref_end;
    uint8_t *is_alt;
} aa_rows;

/* intermediate state for parity tests (only when aa_opts.keep_debug) — vertex ids are
 * contig-local: singles 0..n-1, pair vertices n.., src = V-2, dest = V-1 (paf_data.cpp:286-291,
 * 371-372, 699-700) */
typedef struct aa_debug {
    int64_t *vtx_off;   /* [n_ctg+1] */
    int64_t *edge_off;  /* [n_ctg+1] */
    int64_t *walk_off;  /* [n_ctg+1] */
    int32_t *e_src, *e_dst;          /* [E] adjacency order (OC3)                                 */
    int64_t *e_qry, *e_ref;          /* [E]                                                       */
    int32_t *e_anom, *e_qnz, *e_qtot;
    uint8_t *d_reach;                /* [V] */
    int64_t *d_sum;                  /* [V] qry_score + ref_score of d[v] (CALC_SUM view)          */
    int32_t *d_anom, *d_qnz, *d_qtot;
    int32_t *best;                   /* [V] */
    int32_t *order;                  /* [V] forward Kahn position of each vertex                   */
    int64_t *w_sum;                  /* [W] walk distances, pop order                              */
    int32_t *w_anom, *w_qnz, *w_qtot;
    int64_t *anom_dis;               /* [n_ctg] anom_dis[dest], paf_data.cpp:713                   */
} aa_debug;

typedef struct aa_stats {
    int64_t n_ctg, n_blk, n_run;
    int64_t n_pair;      /* pair vertices P                                                          */
    int64_t n_vtx;       /* V summed over solved contigs (n + P + 2 each)                            */
    int64_t n_edge;      /* E                                                                        */
    int64_t n_heap;      /* persistent leftist-heap nodes H                                          */
    int64_t n_walk;      /* walks enumerated K                                                       */
    int64_t n_task;      /* walks recovered + upgraded (edge_path_to_paf_path calls)                 */
    int64_t n_launch;    /* kernels launched by the last solve                                       */
    double ms_total;     /* device time of the last solve (CUDA events)                              */
    double ms_phase[16]; /* per phase, see aa_phase_name()                                           */
    double algo_bytes;   /* algorithmic bytes of the last solve (DESIGN.md formula)                  */
    double algo_bytes_phase[16];
} aa_stats;

typedef struct aa_result {
    int64_t n_ctg;
    int64_t *out_off;      /* [n_ctg+1] rows of the primary chain per contig                        */
    aa_rows out;
    int64_t *alt_off;      /* [n_ctg+1]                                                             */
    aa_rows alt;
    int64_t *all_path_off; /* [n_ctg+1] -> paths (only with want_all, else all zero)                */
    int64_t *all_row_off;  /* [n_paths+1] -> rows                                                   */
    aa_rows all;
    int32_t *sorted_index; /* [n_blk] ctg_sorted_index of every input block                         */
    aa_debug *dbg;         /* NULL unless keep_debug                                                */
    aa_stats stats;
} aa_result;

typedef struct aa_ctx aa_ctx;             /* one per (host thread, device)                          */
typedef struct aa_dev_batch aa_dev_batch; /* a batch staged in HBM                                  */

/* ---- the hot path ----------------------------------------------------------------------- */
aa_status aa_create(aa_ctx **ctx, int device);
void aa_destroy(aa_ctx *ctx);
const char *aa_last_error(const aa_ctx *ctx);

/* solve_ctg_read over a whole batch, host buffers in, host buffers out (replaces the loop at
 * alignasm.cpp:346-379).  `res` is filled with library-owned memory: release with aa_result_free. */
aa_status aa_solve(aa_ctx *ctx, const aa_batch *batch, const aa_opts *opts, aa_result *res);

/* the same in three steps, so that a caller can keep a batch resident in HBM */
aa_status aa_upload(aa_ctx *ctx, const aa_batch *batch, aa_dev_batch **dev);
aa_status aa_solve_device(aa_ctx *ctx, aa_dev_batch *dev, const aa_opts *opts, aa_result *res /* may be NULL: results stay on the device */);
void aa_dev_batch_free(aa_ctx *ctx, aa_dev_batch *dev);

void aa_result_free(aa_result *res);

/* The same for a subset of the batch's contigs (ascending contig ids; rows of the result are indexed by position in
 * `ctgs`): the blocked_range a TBB worker gets (alignasm.cpp:354-359).  Staged straight from the caller's arrays. */
aa_status aa_solve_subset(aa_ctx *ctx, const aa_batch *batch, const int64_t *ctgs, int64_t n_ctgs, const aa_opts *opts,
                          aa_result *res);
/* ---- the same over several GPUs of one box: contigs are independent (tbb::parallel_for over contigs, alignasm.cpp:351-359),
 * so they are partitioned by a cost estimate (longest processing time first), every shard is solved on its own device from
 * its own host thread, and the rows are merged back in input order.  No collective.  `devices` may name a device twice. */
aa_status aa_solve_multi(const int32_t *devices, int32_t n_dev, const aa_batch *batch, const aa_opts *opts, aa_result *res);
void aa_shard_contigs(const aa_batch *batch, int32_t max_walks, int32_t n_shards, int32_t *shard_of /* [n_ctg] */);
const char *aa_multi_last_error(void);
void aa_multi_release(void); /* aa_solve_multi keeps warm contexts between calls (as many per device as were ever in use at once): free them */
/* statistics (sizes, per-phase CUDA-event times, algorithmic bytes) of the last solve on this context */
aa_status aa_get_stats(const aa_ctx *ctx, aa_stats *stats);
const char *aa_phase_name(int phase); /* NULL past the last phase */
const char *aa_version(void);

/* ---- host side of the drop-in: PAF reader / cs:Z: codec / writers ------------------------- */
typedef struct aa_paf aa_paf;
/* reader + contig bucketing + get_overlap_range (alignasm.cpp:110-181, paf_data.cpp:90-123) */
aa_status aa_paf_read(const char *path, aa_paf **paf, char *err, int64_t err_cap);
/* --alt ingestion (alignasm.cpp:186-332): rows of an alternative PAF whose query names are `<contig>:<START>-<END>`
 * segments are shifted into contig coordinates and appended to that contig's blocks (every row of a segment whose
 * aln_len / segment length exceeds `alt_baseline`, else the segment's best row); they are written back with
 * `xi:Z:A_<row>`.  Call between aa_paf_read and aa_paf_batch; pointers from an earlier aa_paf_batch are invalidated. */
aa_status aa_paf_read_alt(aa_paf *paf, const char *alt_path, double alt_baseline, char *err, int64_t err_cap);
const aa_batch *aa_paf_batch(const aa_paf *paf);
/* writers incl. get_edited_paf_data (alignasm.cpp:398-490, paf_data.cpp:125-220); writes
 * <prefix>.aln.paf, <prefix>.aln.alt.paf, <prefix>.aln.all.paf */
aa_status aa_paf_write(const aa_paf *paf, const aa_result *res, const char *out_prefix, char *err, int64_t err_cap);
void aa_paf_free(aa_paf *paf);

/* ---- the cs:Z: codec on the device (csrc/cs_codec.cu) -------------------------------------------------------------
 * The same two functions of the reference as above, computed by CUDA kernels over the file image (one thread per row, one
 * forward walk of the cs string whatever the strand), bit-identical to the host codec:
 *   aa_cs_runs_device   parse_short_cs + get_overlap_range   (paf_data.cpp:29-72, 90-123)
 *   aa_cs_edit_device   get_edited_paf_data                  (paf_data.cpp:125-220)
 * aa_paf_read_device / aa_paf_write_device are the reader and the writers with those stages on the device of `ctx`
 * (`alignasm --cs_device`); everything else (TSV fields, contig bucketing, row formatting) is the host code above.
 * Per-row error codes stand for the reference's exception sites; aa_cs_error_text gives the text. */
enum {
    AA_CS_ERR_TAG = 1,      /* no short-form cs:Z: tag                      paf_data.cpp:31-33   */
    AA_CS_ERR_LENGTH = 2,   /* invalid :length operation                    :44-47               */
    AA_CS_ERR_SUBST = 3,    /* invalid substitution operation               :50-53               */
    AA_CS_ERR_INDEL = 4,    /* empty indel operation                        :58-61               */
    AA_CS_ERR_OP = 5,       /* unsupported operation                        :65-67               */
    AA_CS_ERR_CONSUME = 6,  /* consumption does not match the coordinates   :119-122             */
    AA_CS_ERR_CLIP_INS = 7, /* alignment clipped inside an insertion        :160-163             */
    AA_CS_ERR_EDIT = 8      /* edited cs does not match edited coordinates  :214-217             */
};
typedef struct aa_cs_rows {  /* n alignment rows: where each cs:Z: field lies in `text`, and the row's (closed) coordinates */
    int64_t n;
    const int64_t *cs_off;   /* [n] byte offset of the field (it starts with "cs:Z:")            */
    const int32_t *cs_len;   /* [n] its length                                                     */
    const int64_t *qry_str, *qry_end, *ref_str, *ref_end; /* [n] like the aa_batch arrays: ref_str > ref_end on '-' */
    const uint8_t *aln_fwd;  /* [n]                                                                */
} aa_cs_rows;
typedef struct aa_cs_runs {  /* library-owned; release with aa_cs_runs_free */
    int64_t n_rows, n_run;
    int64_t *run_off;        /* [n_rows+1] */
    int64_t *run_ql, *run_qr, *run_rl; /* [n_run], query orientation, as aa_batch wants them */
    int32_t *err;            /* [n_rows] 0 or AA_CS_ERR_* (a row in error has no runs) */
} aa_cs_runs;
typedef struct aa_cs_edits { /* library-owned; release with aa_cs_edits_free */
    int64_t n, n_bytes;
    int64_t *off;            /* [n+1] -> text */
    char *text;              /* the edited cs:Z: fields, back to back (no terminators) */
    int32_t *mat_num, *aln_len, *err; /* [n] */
} aa_cs_edits;
aa_status aa_cs_runs_device(aa_ctx *ctx, const char *text, int64_t text_len, const aa_cs_rows *rows, aa_cs_runs *out);
void aa_cs_runs_free(aa_cs_runs *runs);
/* one edited field per output row k: the row `out_row[k]` of `rows` cut to the query interval [eqs[k], eqe[k]]; ers / ere are the
 * edited reference ends (for the consistency check); row_mat / row_aln are the columns 10 / 11 of the rows as read */
aa_status aa_cs_edit_device(aa_ctx *ctx, const char *text, int64_t text_len, const aa_cs_rows *rows, const int32_t *row_mat,
                            const int32_t *row_aln, int64_t n_out, const int64_t *out_row, const int64_t *eqs, const int64_t *eqe,
                            const int64_t *ers, const int64_t *ere, aa_cs_edits *out);
void aa_cs_edits_free(aa_cs_edits *edits);
const char *aa_cs_error_text(int32_t code);
const char *aa_cs_last_error(void); /* per thread: the CUDA / allocation failure behind a non-zero status of the two calls above */
int aa_ctx_device(const aa_ctx *ctx);
aa_status aa_paf_read_device(const char *path, aa_ctx *ctx, aa_paf **paf, char *err, int64_t err_cap);
aa_status aa_paf_write_device(const aa_paf *paf, aa_ctx *ctx, const aa_result *res, const char *out_prefix, char *err, int64_t err_cap);

/* ---- page-locked host memory for the arrays of a batch (optional) ---------------------------
 * aa_solve / aa_upload / aa_solve_subset copy arrays that live in page-locked memory to the device directly; pageable
 * arrays go through the library's staging buffer first (one more pass over the input at host memcpy speed).  The
 * reference keeps its blocks in std::vector (paf_data.hpp:96-107); a caller that fills the aa_batch arrays itself can
 * place them here instead.  NULL when no page-locked memory can be had (use malloc then). */
void *aa_host_alloc(int64_t bytes);
void aa_host_free(void *p);

#ifdef __cplusplus
}
#endif
#endif /* ALIGNASM_B200_H */
