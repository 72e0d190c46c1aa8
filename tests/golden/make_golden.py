#!/usr/bin/env python
"""Generates tests/golden/*: small PAF inputs and the outputs of the REFERENCE ITSELF on them.

Run here (container with /root/reference):   python tests/golden/make_golden.py
It needs oracle/_ref (built by `make -C oracle ref` from the reference sources where they lie) and writes, per case
and per mode (default / --non_skip_linkable):
    <case>.paf                               input (hand-built micro cases or tools/synth_paf.cpp output)
    <case>.altin.paf                         (case `withalt`) the alternative PAF given as `--alt` (tests/alt_util.py)
    <case>[.nsl].aln.paf / .aln.alt.paf / .aln.all.paf   reference outputs, canonical allocator
    <case>[.nsl].dump.npz                    graph edges / d / best / forward order / walk distances / anom_dis of the
                                             first DUMP_CONTIGS contigs, from the hook build (oracle/ref_dump_tu.cpp)
The GPU box has no /root/reference: tests only read these committed files.
"""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle_py  # noqa: E402
import parity_util as pu  # noqa: E402
import alt_util  # noqa: E402

DUMP_CONTIGS = 6
CHR = {"chr1": 248956422, "chr2": 242193529, "chr3": 198295559}


def row(q, qlen, qs, strand, chrom, ts, mapq, ops):
    """ops in QUERY orientation: [(':', n) | ('*', 1) | ('+', n) | ('-', n)]; returns one PAF line."""
    qn = sum(n for t, n in ops if t in ":*+")
    rn = sum(n for t, n in ops if t in ":*-")
    nm = sum(n for t, n in ops if t == ":")
    al = sum(n for t, n in ops)
    txt = {":": lambda n: f":{n}", "*": lambda n: "*ag", "+": lambda n: "+" + "a" * n, "-": lambda n: "-" + "t" * n}
    seq = ops if strand == "+" else ops[::-1]
    cs = "cs:Z:" + "".join(txt[t](n) for t, n in seq)
    return "\t".join(str(x) for x in (q, qlen, qs, qs + qn, strand, chrom, CHR[chrom], ts, ts + rn, nm, al, mapq, "tp:A:P", cs)) + "\n"


def micro_cases():
    M = [(":", 1000)]
    L = []
    L.append(row("m_single", 5000, 100, "+", "chr1", 1000, 60, M))
    # colinear, positive reference gaps
    for k in range(3):
        L.append(row("m_colinear", 10000, 100 + 2000 * k, "+", "chr1", 50000 + 2100 * k, 60, M))
    # reference overlap (negative gap is doubled)
    L.append(row("m_refoverlap", 10000, 0, "+", "chr1", 50000, 60, M))
    L.append(row("m_refoverlap", 10000, 1500, "+", "chr1", 50500, 30, M))
    # same strand, gap beyond SV_BASELINE: anom + cap
    L.append(row("m_farjump", 10000, 0, "+", "chr1", 50000, 60, M))
    L.append(row("m_farjump", 10000, 1500, "+", "chr1", 5000000, 60, M))
    L.append(row("m_farjump", 10000, 3000, "+", "chr1", 5002000, 0, M))
    # inversions, both orders, near and far
    L.append(row("m_inv_fwd", 10000, 0, "+", "chr1", 50000, 60, M))
    L.append(row("m_inv_fwd", 10000, 1500, "-", "chr1", 52000, 60, M))
    L.append(row("m_inv_fwd", 10000, 3000, "+", "chr1", 54000, 60, M))
    L.append(row("m_inv_rev", 10000, 0, "-", "chr1", 50000, 60, M))
    L.append(row("m_inv_rev", 10000, 1500, "+", "chr1", 48000, 60, M))
    L.append(row("m_inv_far", 10000, 0, "+", "chr1", 50000, 60, M))
    L.append(row("m_inv_far", 10000, 1500, "-", "chr1", 9000000, 1, M))
    # translocation
    L.append(row("m_trans", 10000, 0, "+", "chr1", 50000, 60, M))
    L.append(row("m_trans", 10000, 1500, "+", "chr2", 70000, 60, M))
    L.append(row("m_trans", 10000, 3000, "+", "chr1", 53000, 60, M))
    # all minus, colinear in query order
    for k in range(3):
        L.append(row("m_minus", 10000, 100 + 2000 * k, "-", "chr3", 90000 - 2100 * k, 60, M))
    # cut point: equal run starts (paf_data.cpp:315-327), incl. the single-base run that is skipped
    L.append(row("m_cut_eq", 10000, 0, "+", "chr1", 50000, 60, [(":", 100), ("*", 1), (":", 100)]))
    L.append(row("m_cut_eq", 10000, 101, "+", "chr1", 50101, 60, [(":", 300)]))
    L.append(row("m_cut_eq1", 10000, 0, "+", "chr1", 50000, 60, [(":", 100), ("*", 1), (":", 100)]))
    L.append(row("m_cut_eq1", 10000, 101, "+", "chr1", 50101, 60, [(":", 1), ("*", 1), (":", 300)]))
    # cut point: j's run starts inside i's run (paf_data.cpp:328-338), on both strands
    L.append(row("m_cut_in", 10000, 0, "+", "chr1", 50000, 60, [(":", 500)]))
    L.append(row("m_cut_in", 10000, 300, "+", "chr1", 50300, 60, [(":", 500)]))
    L.append(row("m_cut_in_rev", 10000, 0, "-", "chr1", 50000, 60, [(":", 200), ("-", 3), (":", 300)]))
    L.append(row("m_cut_in_rev", 10000, 300, "-", "chr1", 49500, 60, [(":", 250), ("+", 2), (":", 250)]))
    # cut point: i's run starts inside j's run (paf_data.cpp:347-356)
    L.append(row("m_cut_iinj", 10000, 0, "+", "chr1", 50000, 60, [(":", 100), ("+", 50), (":", 51)]))
    L.append(row("m_cut_iinj", 10000, 120, "+", "chr1", 50120, 60, [(":", 181)]))
    # cut point: no run touches: minimum gap (paf_data.cpp:339-345, 360-370)
    L.append(row("m_cut_gap", 10000, 0, "+", "chr1", 50000, 60, [(":", 100), ("+", 100), (":", 1)]))
    L.append(row("m_cut_gap", 10000, 150, "+", "chr1", 50150, 60, [(":", 40), ("*", 1), (":", 200)]))
    # contained block + chain of overlaps (pair vertices (i,j)->(j,k))
    L.append(row("m_chain", 20000, 0, "+", "chr1", 50000, 60, [(":", 1000)]))
    L.append(row("m_chain", 20000, 200, "+", "chr1", 50200, 60, [(":", 300)]))
    L.append(row("m_chain", 20000, 800, "+", "chr1", 50800, 60, [(":", 1000)]))
    L.append(row("m_chain", 20000, 1500, "+", "chr1", 51500, 30, [(":", 1000)]))
    L.append(row("m_chain", 20000, 2400, "+", "chr2", 1000, 0, [(":", 1000)]))
    L.append(row("m_chain", 20000, 5000, "+", "chr1", 55000, 60, [(":", 1000)]))
    # duplicate sort keys on more than 16 rows (std::sort is unstable beyond its insertion-sort threshold: H2),
    # duplicated loci make exact distance ties (H1)
    for k in range(12):
        L.append(row("m_dupkeys", 40000, 1000 + 2500 * k, "+", "chr1", 60000 + 2500 * k, 60, [(":", 2000)]))
        if k % 2 == 0:
            L.append(row("m_dupkeys", 40000, 1000 + 2500 * k, "-", "chr2", 80000 + 100 * k, 60 if k % 4 else 0, [(":", 2000)]))
        if k % 3 == 0:
            L.append(row("m_dupkeys", 40000, 1000 + 2500 * k, "+", "chr3", 7000 + 2500 * k, 30, [(":", 2000)]))
    return L


def gen_outputs(case, paf, tmp):
    alt = paf[:-4] + ".altin.paf"
    alt = alt if os.path.exists(alt) else None
    for nsl in (False, True):
        tag = case + (".nsl" if nsl else "")
        pre = os.path.join(tmp, tag)
        oracle_py.run_ref(paf, pre, variant="canon", non_skip_linkable=nsl, alt=alt)
        for ext in ("aln.paf", "aln.alt.paf", "aln.all.paf"):
            shutil.copy(pre + "." + ext, os.path.join(HERE, tag + "." + ext))
        dump = pre + ".dump"
        oracle_py.run_ref(paf, pre + "_d", variant="dump", non_skip_linkable=nsl, dump=dump, limit_contigs=DUMP_CONTIGS, alt=alt)
        ctgs = oracle_py.parse_dump(dump)
        arrays = {"n": np.array([c["n"] for c in ctgs], dtype=np.int64)}
        for i, c in enumerate(ctgs):
            if c["n"] == 1:
                continue
            arrays[f"edges{i}"] = np.array(c["edges"], dtype=np.int64).reshape(-1, 7)
            arrays[f"d{i}"] = np.array(c["d"], dtype=np.int64).reshape(-1, 8)
            arrays[f"walks{i}"] = np.array(c["walks"], dtype=np.int64).reshape(-1, 5)
            arrays[f"order{i}"] = np.array(c["order"], dtype=np.int32)
            arrays[f"anom{i}"] = np.array([c["anom_dis"]], dtype=np.int64)
        np.savez_compressed(os.path.join(HERE, tag + ".dump.npz"), **arrays)


def main():
    if oracle_py.ref_binary("canon") is None:
        oracle_py.build(ref=True)
    with tempfile.TemporaryDirectory() as tmp:
        cases = {}
        p = os.path.join(HERE, "micro.paf")
        with open(p, "w") as f:
            f.writelines(micro_cases())
        cases["micro"] = p
        cases["tiny"] = pu.synth(os.path.join(HERE, "tiny.paf"), "--contigs", 24, "--blocks", 8, "--sd", 6, "--p_dup", 0.2,
                                 "--p_trans", 0.2, "--p_inv", 0.2, "--seed", 13, "--lmin", 500, "--lmax", 5000, "--gap_max", 200)
        cases["ties"] = pu.synth(os.path.join(HERE, "ties.paf"), "--contigs", 10, "--blocks", 30, "--sd", 10, "--p_dup", 0.15,
                                 "--p_trans", 0.15, "--p_inv", 0.15, "--seed", 11, "--lmin", 2000, "--lmax", 9000)
        cases["dense"] = pu.synth(os.path.join(HERE, "dense.paf"), "--preset", "c4", "--n", 40, "--lmin", 2000, "--lmax", 20000)
        cases["withalt"] = pu.synth(os.path.join(HERE, "withalt.paf"), "--contigs", 12, "--blocks", 14, "--sd", 6, "--p_dup", 0.1,
                                    "--p_trans", 0.15, "--p_inv", 0.15, "--seed", 17, "--lmin", 1500, "--lmax", 8000)
        alt_util.make_alt(cases["withalt"], os.path.join(HERE, "withalt.altin.paf"), seed=5)
        for case, paf in cases.items():
            gen_outputs(case, paf, tmp)
            print("golden", case, os.path.getsize(paf), "bytes")


if __name__ == "__main__":
    main()
