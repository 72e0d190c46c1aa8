"""Size-independent fingerprint of a Result, shared by tests/golden/make_fullsize.py (CPU restatement, committed as
tests/golden/fullsize_<tag>.json) and the GPU parity test that solves the same synthetic workload."""
import hashlib

import numpy as np

ROW_KEYS = ("ctg_index", "qry_str", "qry_end", "ref_str", "ref_end", "is_alt")
DBG_WALK = ("walk_off", "w_sum", "w_anom", "w_qnz", "w_qtot")
STAT_KEYS = ("n_blk", "n_ctg", "n_pair", "n_vtx", "n_edge", "n_heap", "n_walk", "n_task")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def digest(res):
    """Size-independent fingerprint of a Result (needs keep_debug for the walk lists and the per-contig shapes)."""
    d = {"out_off": sha(res.out_off), "alt_off": sha(res.alt_off), "sorted_index": sha(res.sorted_index)}
    for which in ("out", "alt"):
        for k in ROW_KEYS:
            d[f"{which}.{k}"] = sha(getattr(res, which)[k])
    for k in DBG_WALK:
        d["dbg." + k] = sha(res.dbg[k])
    d["dbg.edges"] = sha(np.stack([np.asarray(res.dbg[k], dtype=np.int64) for k in ("e_src", "e_dst", "e_qry", "e_ref", "e_anom", "e_qnz", "e_qtot")]))
    d["dbg.d"] = sha(np.stack([np.asarray(res.dbg[k], dtype=np.int64) for k in ("d_reach", "d_sum", "d_anom", "d_qnz", "d_qtot", "best")]))
    per = np.stack([np.diff(np.asarray(res.dbg["vtx_off"])), np.diff(np.asarray(res.dbg["edge_off"])), np.diff(np.asarray(res.dbg["walk_off"]))])
    d["per_contig(V,E,K)"] = sha(per.astype(np.int64))
    return d, per



# tag -> (tools/synth_paf.cpp arguments, solve options): the inputs pinned at full size by tests/golden/make_fullsize.py
PINS = {
    "c2": (["--preset", "c2", "--seed", 2], {}),
    "c3": (["--preset", "c3", "--seed", 3], {}),
    # BASELINE config 4 ladder (one dense contig): the sizes the CPU restatement still finishes
    "dense845": (["--preset", "c4", "--n", 845], {}),
    "dense845.nsl": (["--preset", "c4", "--n", 845], {"non_skip_linkable": True}),
    "dense1645": (["--preset", "c4", "--n", 1645], {}),
    "dense1645.nsl": (["--preset", "c4", "--n", 1645], {"non_skip_linkable": True}),
    "dense3290.nsl": (["--preset", "c4", "--n", 3290], {"non_skip_linkable": True}),
}
