"""Compare a Result (oracle port / emulated core / CUDA product) with the committed golden fixtures."""
import os

import numpy as np

import parity_util as pu

CASES = ["micro", "tiny", "ties", "dense", "withalt"]


def open_case(case):
    """The golden input as a parsed PafFile; case `withalt` carries an alternative PAF (`--alt`, alignasm.cpp:186-332)."""
    import alignasm_b200 as aa
    alt = os.path.join(pu.GOLDEN, case + ".altin.paf")
    return aa.read_paf(os.path.join(pu.GOLDEN, case + ".paf"), alt=alt if os.path.exists(alt) else None)


def check_against_golden(case, nsl, solve_fn, paf_file, workdir):
    """solve_fn(batch, non_skip_linkable=..., want_all=True, keep_debug=True) -> Result."""
    tag = case + (".nsl" if nsl else "")
    res = solve_fn(paf_file.batch, non_skip_linkable=nsl, want_all=True, keep_debug=True)
    pre = os.path.join(workdir, tag + "_got")
    paf_file.write(res, pre)
    for ext in ("aln.paf", "aln.alt.paf", "aln.all.paf"):
        want = os.path.join(pu.GOLDEN, tag + "." + ext)
        assert pu.files_equal(pre + "." + ext, want), f"{tag}.{ext}:\n" + pu.first_diff(pre + "." + ext, want)
    gold = np.load(os.path.join(pu.GOLDEN, tag + ".dump.npz"))
    dbg = res.dbg
    for c, n in enumerate(gold["n"].tolist()):
        if n == 1:
            continue
        e0, e1 = int(dbg["edge_off"][c]), int(dbg["edge_off"][c + 1])
        v0, v1 = int(dbg["vtx_off"][c]), int(dbg["vtx_off"][c + 1])
        w0, w1 = int(dbg["walk_off"][c]), int(dbg["walk_off"][c + 1])
        edges = np.stack([dbg[k][e0:e1].astype(np.int64) for k in ("e_src", "e_dst", "e_qry", "e_ref", "e_anom", "e_qnz", "e_qtot")], axis=1)
        assert np.array_equal(edges, gold[f"edges{c}"]), f"{tag} contig {c}: ordered edge list"
        assert int(dbg["anom_dis"][c]) == int(gold[f"anom{c}"][0]), f"{tag} contig {c}: anom_dis"
        d = gold[f"d{c}"]  # v, reach, qry, ref, anom, qnz, qtot, best
        assert v1 - v0 == len(d)
        reach = d[:, 1].astype(bool)
        assert np.array_equal(dbg["d_reach"][v0:v1].astype(bool), reach), f"{tag} contig {c}: reachability"
        assert np.array_equal(dbg["d_sum"][v0:v1][reach], (d[:, 2] + d[:, 3])[reach]), f"{tag} contig {c}: d.sum"
        for k, col in (("d_anom", 4), ("d_qnz", 5), ("d_qtot", 6)):
            assert np.array_equal(dbg[k][v0:v1][reach], d[:, col][reach]), f"{tag} contig {c}: {k}"
        assert np.array_equal(dbg["best"][v0:v1], d[:, 7]), f"{tag} contig {c}: best"
        order = np.empty(v1 - v0, dtype=np.int64)
        order[gold[f"order{c}"]] = np.arange(v1 - v0)
        assert np.array_equal(dbg["order"][v0:v1], order), f"{tag} contig {c}: forward Kahn order"
        wk = gold[f"walks{c}"]  # qry, ref, anom, qnz, qtot
        assert w1 - w0 == len(wk), f"{tag} contig {c}: walk count"
        assert np.array_equal(dbg["w_sum"][w0:w1], wk[:, 0] + wk[:, 1]), f"{tag} contig {c}: walk sums"
        for k, col in (("w_anom", 2), ("w_qnz", 3), ("w_qtot", 4)):
            assert np.array_equal(dbg[k][w0:w1], wk[:, col]), f"{tag} contig {c}: {k}"
    return res
