"""alignasm_b200 — B200-native implementation of the alignasm hot path (solve_ctg_read).

Python is only the binding: every entry point calls the C ABI of ``libalignasm_b200.so``
(include/alignasm_b200.h), whose hot path is hand-written sm_100a CUDA.  There is no CPU
fallback: if the library is missing or no CUDA device is usable, calls raise.

Mirror of the reference interface (reference src/paf_data.hpp:191-193, src/alignasm.cpp):
    read_paf(path)                         reader + bucketing + get_overlap_range
    Solver(device).solve(batch, ...)       the solve_ctg_read loop over all contigs
    solve_ctg_read(blocks, ...)            one contig, same argument meaning as the reference
    PafFile.write(result, prefix)          the three writers incl. get_edited_paf_data
"""
import ctypes as C
import os

import numpy as np

from . import _abi
from ._abi import aa_batch, aa_opts, aa_result, np_from, np_view, ptr_of

__all__ = ["Batch", "PafFile", "Result", "Solver", "AlignasmError", "read_paf", "solve_ctg_read", "solve_multi",
           "shard_contigs", "lib_path", "load_library"]

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class AlignasmError(RuntimeError):
    def __init__(self, status, msg=""):
        self.status = int(status)
        super().__init__(f"{_abi.STATUS.get(int(status), status)}: {msg}")


def lib_path():
    return os.path.join(_HERE, "libalignasm_b200.so")


def load_library():
    """Load the C-ABI library.  Fails loudly when it has not been built (no fallback)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError(f"{path} is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a).  alignasm_b200 has no CPU fallback.")
    lib = C.CDLL(path)
    vp = C.c_void_p
    lib.aa_create.argtypes = [C.POINTER(vp), C.c_int]
    lib.aa_create.restype = C.c_int
    lib.aa_destroy.argtypes = [vp]
    lib.aa_destroy.restype = None
    lib.aa_last_error.argtypes = [vp]
    lib.aa_last_error.restype = C.c_char_p
    lib.aa_solve.argtypes = [vp, C.POINTER(aa_batch), C.POINTER(aa_opts), C.POINTER(aa_result)]
    lib.aa_solve.restype = C.c_int
    lib.aa_upload.argtypes = [vp, C.POINTER(aa_batch), C.POINTER(vp)]
    lib.aa_upload.restype = C.c_int
    lib.aa_solve_device.argtypes = [vp, vp, C.POINTER(aa_opts), C.POINTER(aa_result)]
    lib.aa_solve_device.restype = C.c_int
    lib.aa_dev_batch_free.argtypes = [vp, vp]
    lib.aa_dev_batch_free.restype = None
    lib.aa_result_free.argtypes = [C.POINTER(aa_result)]
    lib.aa_result_free.restype = None
    lib.aa_get_stats.argtypes = [vp, C.POINTER(_abi.aa_stats)]
    lib.aa_get_stats.restype = C.c_int
    lib.aa_phase_name.argtypes = [C.c_int]
    lib.aa_phase_name.restype = C.c_char_p
    lib.aa_version.argtypes = []
    lib.aa_version.restype = C.c_char_p
    lib.aa_paf_read.argtypes = [C.c_char_p, C.POINTER(vp), C.c_char_p, C.c_int64]
    lib.aa_paf_read.restype = C.c_int
    lib.aa_paf_read_alt.argtypes = [vp, C.c_char_p, C.c_double, C.c_char_p, C.c_int64]
    lib.aa_paf_read_alt.restype = C.c_int
    lib.aa_paf_batch.argtypes = [vp]
    lib.aa_paf_batch.restype = C.POINTER(aa_batch)
    lib.aa_paf_write.argtypes = [vp, C.POINTER(aa_result), C.c_char_p, C.c_char_p, C.c_int64]
    lib.aa_paf_write.restype = C.c_int
    lib.aa_paf_free.argtypes = [vp]
    lib.aa_paf_free.restype = None
    lib.aa_paf_read_device.argtypes = [C.c_char_p, vp, C.POINTER(vp), C.c_char_p, C.c_int64]
    lib.aa_paf_read_device.restype = C.c_int
    lib.aa_paf_write_device.argtypes = [vp, vp, C.POINTER(aa_result), C.c_char_p, C.c_char_p, C.c_int64]
    lib.aa_paf_write_device.restype = C.c_int
    lib.aa_host_alloc.argtypes = [C.c_int64]
    lib.aa_host_alloc.restype = vp
    lib.aa_host_free.argtypes = [vp]
    lib.aa_host_free.restype = None
    lib.aa_solve_multi.argtypes = [C.POINTER(C.c_int32), C.c_int32, C.POINTER(aa_batch), C.POINTER(aa_opts), C.POINTER(aa_result)]
    lib.aa_solve_multi.restype = C.c_int
    lib.aa_shard_contigs.argtypes = [C.POINTER(aa_batch), C.c_int32, C.c_int32, C.POINTER(C.c_int32)]
    lib.aa_shard_contigs.restype = None
    lib.aa_multi_last_error.argtypes = []
    lib.aa_multi_last_error.restype = C.c_char_p
    _LIB = lib
    return lib


class _Pins:
    """Owner of aa_host_alloc blocks (freed with the batch that views them)."""

    def __init__(self, ptrs):
        self.ptrs = ptrs

    def __del__(self):
        try:
            lib = load_library()
            for p in self.ptrs:
                lib.aa_host_free(p)
        except Exception:
            pass
        self.ptrs = []


class Batch:
    """A batch of contigs as host structure-of-arrays (the aa_batch of the C ABI)."""

    FIELDS = [("ctg_off", np.int64), ("qry_str", np.int64), ("qry_end", np.int64), ("ref_str", np.int64),
              ("ref_end", np.int64), ("qry_total", np.int64), ("ref_chr", np.int32), ("aln_fwd", np.uint8),
              ("map_qul", np.uint8), ("run_off", np.int64), ("run_ql", np.int64), ("run_qr", np.int64),
              ("run_rl", np.int64)]

    def __init__(self, **arrays):
        for name, dt in self.FIELDS:
            setattr(self, name, np.ascontiguousarray(arrays[name], dtype=dt))
        self.n_ctg = len(self.ctg_off) - 1
        self.n_blk = len(self.qry_str)
        self.n_run = len(self.run_ql)
        assert len(self.run_off) == self.n_blk + 1 and int(self.ctg_off[-1]) == self.n_blk
        self._c = None

    @classmethod
    def from_c(cls, cb):
        """Deep-copy an aa_batch (e.g. the one a PafFile owns)."""
        nb, nc, nr = cb.n_blk, cb.n_ctg, cb.n_run
        sizes = {"ctg_off": nc + 1, "run_off": nb + 1, "run_ql": nr, "run_qr": nr, "run_rl": nr}
        return cls(**{name: np_from(getattr(cb, name), sizes.get(name, nb)) for name, _ in cls.FIELDS})

    def c_struct(self):
        if self._c is None:
            b = aa_batch()
            b.n_ctg, b.n_blk, b.n_run = self.n_ctg, self.n_blk, self.n_run
            ct = {np.int64: C.c_int64, np.int32: C.c_int32, np.uint8: C.c_uint8}
            for name, dt in self.FIELDS:
                setattr(b, name, ptr_of(getattr(self, name), ct[dt]))
            self._c = b
        return self._c

    def pinned(self):
        """A copy of the batch whose arrays live in page-locked host memory (aa_host_alloc): solve() then copies them to the
        device directly instead of through the library's staging buffer."""
        lib = load_library()
        arrays, holds = {}, []
        for name, dt in self.FIELDS:
            src = getattr(self, name)
            p = lib.aa_host_alloc(max(src.nbytes, 1))
            if not p:
                for q in holds:
                    lib.aa_host_free(q)
                raise MemoryError("aa_host_alloc: no page-locked memory")
            holds.append(p)
            dst = np.frombuffer((C.c_char * max(src.nbytes, 1)).from_address(p), dtype=dt, count=src.size)
            dst[:] = src
            arrays[name] = dst
        out = Batch(**arrays)
        for name, _ in self.FIELDS:  # np.ascontiguousarray keeps an already contiguous array of the right dtype as it is
            assert getattr(out, name).ctypes.data == arrays[name].ctypes.data
        out._pins = _Pins(holds)
        return out

    def select(self, contigs):
        """Sub-batch holding the given contigs (used for sharding across GPUs / ranks)."""
        contigs = np.asarray(contigs, dtype=np.int64)

        def ranges(lo, cnt):  # concatenation of arange(lo[i], lo[i] + cnt[i]) without a Python loop
            total = int(cnt.sum())
            if total == 0:
                return np.zeros(0, np.int64)
            start = np.cumsum(cnt) - cnt
            return np.repeat(lo - start, cnt) + np.arange(total, dtype=np.int64)

        lo, hi = self.ctg_off[contigs], self.ctg_off[contigs + 1]
        blk = ranges(lo, hi - lo)
        rlo = self.run_off[blk]
        rcnt = self.run_off[blk + 1] - rlo
        run = ranges(rlo, rcnt)
        arrays = {name: getattr(self, name)[blk] for name, _ in self.FIELDS
                  if name not in ("ctg_off", "run_off", "run_ql", "run_qr", "run_rl")}
        arrays["ctg_off"] = np.concatenate([[0], np.cumsum(hi - lo)])
        arrays["run_off"] = np.concatenate([[0], np.cumsum(rcnt)])
        for name in ("run_ql", "run_qr", "run_rl"):
            arrays[name] = getattr(self, name)[run]
        return Batch(**arrays)


def _rows(r, get=None):
    n = r.n
    get = get or np_from
    return {"ctg_index": get(r.ctg_index, n), "qry_str": get(r.qry_str, n), "qry_end": get(r.qry_end, n),
            "ref_str": get(r.ref_str, n), "ref_end": get(r.ref_end, n), "is_alt": get(r.is_alt, n)}


def _stats_dict(st):
    d = {k: getattr(st, k) for k in ("n_ctg", "n_blk", "n_run", "n_pair", "n_vtx", "n_edge", "n_heap", "n_walk", "n_task",
                                     "n_launch", "ms_total", "algo_bytes")}
    d["ms_phase"] = list(st.ms_phase)
    d["algo_bytes_phase"] = list(st.algo_bytes_phase)
    return d


class Result:
    """Host copy of an aa_result: three per-contig CSR lists of PafOutputData rows + statistics."""

    def __init__(self, cres, n_blk, free_fn=None, copy=True):
        """copy=False: the row arrays are views of the library's buffers (as a C caller sees them), valid until close()."""
        nc = cres.n_ctg
        get = np_from if copy else np_view
        self.n_ctg = nc
        self.out_off = get(cres.out_off, nc + 1)
        self.alt_off = get(cres.alt_off, nc + 1)
        self.all_path_off = get(cres.all_path_off, nc + 1)
        npaths = int(self.all_path_off[-1]) if nc >= 0 and len(self.all_path_off) else 0
        self.all_row_off = get(cres.all_row_off, npaths + 1)
        self.out, self.alt, self.all = _rows(cres.out, get), _rows(cres.alt, get), _rows(cres.all, get)
        self.sorted_index = get(cres.sorted_index, n_blk)
        self.stats = _stats_dict(cres.stats)
        self.dbg = None
        if cres.dbg:
            g = cres.dbg.contents
            vo, eo, wo = np_from(g.vtx_off, nc + 1), np_from(g.edge_off, nc + 1), np_from(g.walk_off, nc + 1)
            nv, ne, nw = int(vo[-1]), int(eo[-1]), int(wo[-1])
            self.dbg = {"vtx_off": vo, "edge_off": eo, "walk_off": wo, "anom_dis": np_from(g.anom_dis, nc)}
            for k in ("e_src", "e_dst", "e_qry", "e_ref", "e_anom", "e_qnz", "e_qtot"):
                self.dbg[k] = np_from(getattr(g, k), ne)
            for k in ("d_reach", "d_sum", "d_anom", "d_qnz", "d_qtot", "best", "order"):
                self.dbg[k] = np_from(getattr(g, k), nv)
            for k in ("w_sum", "w_anom", "w_qnz", "w_qtot"):
                self.dbg[k] = np_from(getattr(g, k), nw)
        self._c = cres
        self._free = free_fn

    def rows_of(self, which, contig):
        """List of (ctg_index, qs, qe, rs, re, is_alt) tuples of one contig's chain."""
        off = {"out": self.out_off, "alt": self.alt_off}[which]
        r = getattr(self, which)
        a, b = int(off[contig]), int(off[contig + 1])
        return [tuple(int(r[k][i]) for k in ("ctg_index", "qry_str", "qry_end", "ref_str", "ref_end", "is_alt"))
                for i in range(a, b)]

    def all_of(self, contig):
        res = []
        for m in range(int(self.all_path_off[contig]), int(self.all_path_off[contig + 1])):
            a, b = int(self.all_row_off[m]), int(self.all_row_off[m + 1])
            res.append([tuple(int(self.all[k][i]) for k in ("ctg_index", "qry_str", "qry_end", "ref_str", "ref_end", "is_alt"))
                        for i in range(a, b)])
        return res

    def close(self):
        if self._free is not None and self._c is not None:
            self._free(C.byref(self._c))
        self._c = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PafFile:
    """A parsed PAF (reader + bucketing + cs:Z: -> runs), owned by the C library."""

    def __init__(self, path, alt=None, alt_baseline=0.5, solver=None):
        """`alt`: an alternative PAF (reference CLI `--alt`, `--alt_baseline`; alignasm.cpp:186-332).
        `solver`: parse the cs:Z: tags on that Solver's device (aa_paf_read_device) instead of on the host."""
        lib = load_library()
        self._lib = lib
        h = C.c_void_p()
        err = C.create_string_buffer(512)
        if solver is not None:
            st = lib.aa_paf_read_device(os.fsencode(path), solver._h, C.byref(h), err, 512)
        else:
            st = lib.aa_paf_read(os.fsencode(path), C.byref(h), err, 512)
        if st != 0:
            raise AlignasmError(st, err.value.decode())
        self._h = h
        self.path = path
        self._batch = None
        if alt is not None:
            st = lib.aa_paf_read_alt(h, os.fsencode(alt), float(alt_baseline), err, 512)
            if st != 0:
                lib.aa_paf_free(h)
                self._h = None
                raise AlignasmError(st, err.value.decode())

    @property
    def batch(self):
        if self._batch is None:
            self._batch = Batch.from_c(self._lib.aa_paf_batch(self._h).contents)
        return self._batch

    def write(self, result, out_prefix, solver=None):
        """`solver`: re-cut the cs:Z: tags of the primary / alternative rows on that Solver's device (aa_paf_write_device)."""
        err = C.create_string_buffer(512)
        if solver is not None:
            st = self._lib.aa_paf_write_device(self._h, solver._h, C.byref(result._c), os.fsencode(out_prefix), err, 512)
        else:
            st = self._lib.aa_paf_write(self._h, C.byref(result._c), os.fsencode(out_prefix), err, 512)
        if st != 0:
            raise AlignasmError(st, err.value.decode())

    def close(self):
        if self._h:
            self._lib.aa_paf_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def read_paf(path, alt=None, alt_baseline=0.5, solver=None):
    return PafFile(path, alt=alt, alt_baseline=alt_baseline, solver=solver)


def _opts(non_skip_linkable=False, want_all=False, max_walks=0, keep_debug=False):
    o = aa_opts()
    o.non_skip_linkable, o.want_all, o.max_walks, o.keep_debug = int(non_skip_linkable), int(want_all), int(max_walks), int(keep_debug)
    return o


class Solver:
    """One aa_ctx: a CUDA device plus its workspace.  Raises when no device is usable."""

    def __init__(self, device=0):
        self._lib = load_library()
        h = C.c_void_p()
        st = self._lib.aa_create(C.byref(h), int(device))
        if st != 0:
            msg = self._lib.aa_last_error(h if h else None).decode()
            raise AlignasmError(st, msg)
        self._h = h
        self.device = device

    def _check(self, st, res=None):
        if st != 0:
            msg = self._lib.aa_last_error(self._h).decode()
            if res is not None:  # AA_ERR_UNSOLVABLE leaves a filled result behind: it is the library's memory, give it back
                self._lib.aa_result_free(C.byref(res))
            raise AlignasmError(st, msg)

    def solve(self, batch, copy=True, **kw):
        """Host buffers in, host result out (the e2e path: H2D + kernels + D2H).  copy=False: see Result."""
        res = aa_result()
        o = _opts(**kw)
        self._check(self._lib.aa_solve(self._h, C.byref(batch.c_struct()), C.byref(o), C.byref(res)), res)
        return Result(res, batch.n_blk, self._lib.aa_result_free, copy=copy)

    def upload(self, batch):
        d = C.c_void_p()
        self._check(self._lib.aa_upload(self._h, C.byref(batch.c_struct()), C.byref(d)))
        return _DevBatch(self, d, batch.n_blk)

    def solve_device(self, dev, fetch=True, copy=True, **kw):
        """Solve a batch that is already resident in HBM.  fetch=False leaves the result on the device; copy=False: see Result."""
        o = _opts(**kw)
        if not fetch:
            self._check(self._lib.aa_solve_device(self._h, dev._h, C.byref(o), None))
            return None
        res = aa_result()
        self._check(self._lib.aa_solve_device(self._h, dev._h, C.byref(o), C.byref(res)), res)
        return Result(res, dev.n_blk, self._lib.aa_result_free, copy=copy)

    def stats(self):
        """Sizes, per-phase device times and algorithmic bytes of the last solve on this context."""
        st = _abi.aa_stats()
        self._check(self._lib.aa_get_stats(self._h, C.byref(st)))
        return _stats_dict(st)

    def phase_names(self):
        names, i = [], 0
        while True:
            s = self._lib.aa_phase_name(i)
            if not s:
                return names
            names.append(s.decode())
            i += 1

    def close(self):
        if self._h:
            self._lib.aa_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _DevBatch:
    def __init__(self, solver, h, n_blk):
        self._solver, self._h, self.n_blk = solver, h, n_blk

    def free(self):
        if self._h:
            self._solver._lib.aa_dev_batch_free(self._solver._h, self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def solve_multi(batch, devices, copy=True, **kw):
    """The batch sharded by contig over several GPUs of this box (aa_solve_multi: LPT shards, one host thread and
    one context per device, rows merged in input order; no collective)."""
    lib = load_library()
    dev = (C.c_int32 * len(devices))(*[int(d) for d in devices])
    res = aa_result()
    o = _opts(**kw)
    st = lib.aa_solve_multi(dev, len(devices), C.byref(batch.c_struct()), C.byref(o), C.byref(res))
    if st != 0:
        msg = (lib.aa_multi_last_error() or b"").decode()
        lib.aa_result_free(C.byref(res))  # (a no-op on an empty result; AA_ERR_UNSOLVABLE leaves a filled one)
        raise AlignasmError(st, msg)
    return Result(res, batch.n_blk, lib.aa_result_free, copy=copy)


def cs_runs_device(solver, text, cs_off, cs_len, qry_str, qry_end, ref_str, ref_end, aln_fwd):
    """parse_short_cs + get_overlap_range on the device (aa_cs_runs_device): `text` holds the cs:Z: fields of the rows at
    cs_off / cs_len.  Returns (run_off, run_ql, run_qr, run_rl, err) as numpy arrays."""
    lib = load_library()
    lib.aa_cs_runs_device.argtypes = [C.c_void_p, C.c_char_p, C.c_int64, C.POINTER(_abi.aa_cs_rows), C.POINTER(_abi.aa_cs_runs)]
    lib.aa_cs_runs_device.restype = C.c_int
    lib.aa_cs_runs_free.argtypes = [C.POINTER(_abi.aa_cs_runs)]
    lib.aa_cs_last_error.restype = C.c_char_p
    arr = {"cs_off": np.ascontiguousarray(cs_off, dtype=np.int64), "cs_len": np.ascontiguousarray(cs_len, dtype=np.int32),
           "qry_str": np.ascontiguousarray(qry_str, dtype=np.int64), "qry_end": np.ascontiguousarray(qry_end, dtype=np.int64),
           "ref_str": np.ascontiguousarray(ref_str, dtype=np.int64), "ref_end": np.ascontiguousarray(ref_end, dtype=np.int64),
           "aln_fwd": np.ascontiguousarray(aln_fwd, dtype=np.uint8)}
    rows = _abi.aa_cs_rows(len(arr["cs_off"]), ptr_of(arr["cs_off"], C.c_int64), ptr_of(arr["cs_len"], C.c_int32), ptr_of(arr["qry_str"], C.c_int64),
                           ptr_of(arr["qry_end"], C.c_int64), ptr_of(arr["ref_str"], C.c_int64), ptr_of(arr["ref_end"], C.c_int64),
                           ptr_of(arr["aln_fwd"], C.c_uint8))
    out = _abi.aa_cs_runs()
    st = lib.aa_cs_runs_device(solver._h, text, len(text), C.byref(rows), C.byref(out))
    if st != 0:
        raise AlignasmError(st, (lib.aa_cs_last_error() or b"").decode())
    n, r = out.n_rows, out.n_run
    res = (np_from(out.run_off, n + 1), np_from(out.run_ql, r), np_from(out.run_qr, r), np_from(out.run_rl, r), np_from(out.err, n))
    lib.aa_cs_runs_free(C.byref(out))
    return res


def cs_error_text(code):
    lib = load_library()
    lib.aa_cs_error_text.argtypes = [C.c_int32]
    lib.aa_cs_error_text.restype = C.c_char_p
    return lib.aa_cs_error_text(int(code)).decode()


def shard_contigs(batch, n_shards, max_walks=0):
    """Shard id of every contig under the library's cost model (aa_shard_contigs)."""
    lib = load_library()
    out = np.zeros(batch.n_ctg, dtype=np.int32)
    lib.aa_shard_contigs(C.byref(batch.c_struct()), int(max_walks), int(n_shards), ptr_of(out, C.c_int32))
    return out


def solve_ctg_read(blocks, solver=None, non_skip_linkable=False):
    """One contig, the reference's seam (src/paf_data.hpp:193).

    blocks: list of dicts in file order with qry_str, qry_end, ref_str, ref_end (closed; ref_str > ref_end
    on '-'), qry_total_length, ref_chr, aln_fwd, map_qul, runs=[(q_l, q_r, r_l), ...].
    Returns (out, alt_out, max_out, sorted_index) with rows as (ctg_index, qs, qe, rs, re, is_alt).
    """
    solver = solver or Solver()
    n = len(blocks)
    runs = [r for b in blocks for r in b["runs"]]
    batch = Batch(
        ctg_off=[0, n], qry_str=[b["qry_str"] for b in blocks], qry_end=[b["qry_end"] for b in blocks],
        ref_str=[b["ref_str"] for b in blocks], ref_end=[b["ref_end"] for b in blocks],
        qry_total=[b["qry_total_length"] for b in blocks], ref_chr=[b["ref_chr"] for b in blocks],
        aln_fwd=[1 if b["aln_fwd"] else 0 for b in blocks], map_qul=[b["map_qul"] for b in blocks],
        run_off=np.concatenate([[0], np.cumsum([len(b["runs"]) for b in blocks])]),
        run_ql=[r[0] for r in runs], run_qr=[r[1] for r in runs], run_rl=[r[2] for r in runs])
    res = solver.solve(batch, non_skip_linkable=non_skip_linkable, want_all=True)
    return res.rows_of("out", 0), res.rows_of("alt", 0), res.all_of(0), res.sorted_index
