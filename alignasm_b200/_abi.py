"""ctypes mirror of include/alignasm_b200.h (the C ABI).  Structure layouts only — no logic."""
import ctypes as C

import numpy as np

i64p = C.POINTER(C.c_int64)
i32p = C.POINTER(C.c_int32)
u8p = C.POINTER(C.c_uint8)


class aa_batch(C.Structure):
    _fields_ = [
        ("n_ctg", C.c_int64), ("n_blk", C.c_int64), ("n_run", C.c_int64),
        ("ctg_off", i64p),
        ("qry_str", i64p), ("qry_end", i64p), ("ref_str", i64p), ("ref_end", i64p), ("qry_total", i64p),
        ("ref_chr", i32p), ("aln_fwd", u8p), ("map_qul", u8p),
        ("run_off", i64p), ("run_ql", i64p), ("run_qr", i64p), ("run_rl", i64p),
    ]


class aa_opts(C.Structure):
    _fields_ = [("non_skip_linkable", C.c_int32), ("want_all", C.c_int32), ("max_walks", C.c_int32),
                ("keep_debug", C.c_int32)]


class aa_rows(C.Structure):
    _fields_ = [("n", C.c_int64), ("ctg_index", i32p), ("qry_str", i64p), ("qry_end", i64p), ("ref_str", i64p),
                ("ref_end", i64p), ("is_alt", u8p)]


class aa_debug(C.Structure):
    _fields_ = [
        ("vtx_off", i64p), ("edge_off", i64p), ("walk_off", i64p),
        ("e_src", i32p), ("e_dst", i32p), ("e_qry", i64p), ("e_ref", i64p),
        ("e_anom", i32p), ("e_qnz", i32p), ("e_qtot", i32p),
        ("d_reach", u8p), ("d_sum", i64p), ("d_anom", i32p), ("d_qnz", i32p), ("d_qtot", i32p),
        ("best", i32p), ("order", i32p),
        ("w_sum", i64p), ("w_anom", i32p), ("w_qnz", i32p), ("w_qtot", i32p),
        ("anom_dis", i64p),
    ]


class aa_stats(C.Structure):
    _fields_ = [
        ("n_ctg", C.c_int64), ("n_blk", C.c_int64), ("n_run", C.c_int64), ("n_pair", C.c_int64),
        ("n_vtx", C.c_int64), ("n_edge", C.c_int64), ("n_heap", C.c_int64), ("n_walk", C.c_int64),
        ("n_task", C.c_int64), ("n_launch", C.c_int64),
        ("ms_total", C.c_double), ("ms_phase", C.c_double * 16),
        ("algo_bytes", C.c_double), ("algo_bytes_phase", C.c_double * 16),
    ]


class aa_result(C.Structure):
    _fields_ = [
        ("n_ctg", C.c_int64),
        ("out_off", i64p), ("out", aa_rows),
        ("alt_off", i64p), ("alt", aa_rows),
        ("all_path_off", i64p), ("all_row_off", i64p), ("all", aa_rows),
        ("sorted_index", i32p),
        ("dbg", C.POINTER(aa_debug)),
        ("stats", aa_stats),
    ]


STATUS = {0: "AA_OK", 1: "AA_ERR_INVALID", 2: "AA_ERR_NO_DEVICE", 3: "AA_ERR_CUDA", 4: "AA_ERR_NOMEM",
          5: "AA_ERR_IO", 6: "AA_ERR_FORMAT", 7: "AA_ERR_UNSOLVABLE"}

# every symbol include/alignasm_b200.h declares (tests check that the built library exports them all)
EXPORTS = ["aa_create", "aa_destroy", "aa_last_error", "aa_solve", "aa_upload", "aa_solve_device",
           "aa_dev_batch_free", "aa_result_free", "aa_get_stats", "aa_phase_name", "aa_version",
           "aa_paf_read", "aa_paf_read_alt", "aa_paf_batch", "aa_paf_write", "aa_paf_free", "aa_solve_subset", "aa_solve_multi", "aa_shard_contigs",
           "aa_multi_last_error", "aa_multi_release",
           # cs:Z: codec on the device (csrc/cs_codec.cu)
           "aa_cs_runs_device", "aa_cs_runs_free", "aa_cs_edit_device", "aa_cs_edits_free", "aa_cs_error_text", "aa_cs_last_error",
           "aa_ctx_device", "aa_paf_read_device", "aa_paf_write_device", "aa_host_alloc", "aa_host_free"]


class aa_cs_rows(C.Structure):
    _fields_ = [("n", C.c_int64), ("cs_off", C.POINTER(C.c_int64)), ("cs_len", C.POINTER(C.c_int32)), ("qry_str", C.POINTER(C.c_int64)),
                ("qry_end", C.POINTER(C.c_int64)), ("ref_str", C.POINTER(C.c_int64)), ("ref_end", C.POINTER(C.c_int64)),
                ("aln_fwd", C.POINTER(C.c_uint8))]


class aa_cs_runs(C.Structure):
    _fields_ = [("n_rows", C.c_int64), ("n_run", C.c_int64), ("run_off", C.POINTER(C.c_int64)), ("run_ql", C.POINTER(C.c_int64)),
                ("run_qr", C.POINTER(C.c_int64)), ("run_rl", C.POINTER(C.c_int64)), ("err", C.POINTER(C.c_int32))]


class aa_cs_edits(C.Structure):
    _fields_ = [("n", C.c_int64), ("n_bytes", C.c_int64), ("off", C.POINTER(C.c_int64)), ("text", C.POINTER(C.c_char)),
                ("mat_num", C.POINTER(C.c_int32)), ("aln_len", C.POINTER(C.c_int32)), ("err", C.POINTER(C.c_int32))]

_NP = {C.c_int64: np.int64, C.c_int32: np.int32, C.c_uint8: np.uint8}


def np_from(ptr, n):
    """Copy n elements out of a ctypes pointer into a fresh numpy array."""
    dt = _NP[ptr._type_]
    if n <= 0 or not ptr:
        return np.zeros(0, dtype=dt)
    return np.ctypeslib.as_array(ptr, shape=(int(n),)).astype(dt, copy=True)


def np_view(ptr, n):
    """n elements behind a ctypes pointer as a numpy array WITHOUT copying: valid only while the C memory lives."""
    dt = _NP[ptr._type_]
    if n <= 0 or not ptr:
        return np.zeros(0, dtype=dt)
    return np.ctypeslib.as_array(ptr, shape=(int(n),))


def ptr_of(arr, ctype):
    return arr.ctypes.data_as(C.POINTER(ctype))
