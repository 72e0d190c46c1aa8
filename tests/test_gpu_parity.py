"""GPU parity: the CUDA path, called through the C ABI, against the oracle on the same seeded inputs.
Bit-exact: output rows, ctg_sorted_index, ordered edge lists, d/best, forward order, walk distances."""
import os

import numpy as np
import pytest

import parity_util as pu
from shapes import SMALL

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def solver(product_lib):
    import alignasm_b200 as aa
    return aa.Solver(0)


@pytest.mark.parametrize("name", list(SMALL))
def test_matches_oracle(name, solver, workdir):
    import alignasm_b200 as aa
    from oracle import oracle_py
    args, variants = SMALL[name]
    paf = pu.synth(os.path.join(workdir, name + ".paf"), *args)
    pf = aa.read_paf(paf)
    for nsl in variants:
        got = solver.solve(pf.batch, non_skip_linkable=nsl, want_all=True, keep_debug=True)
        want = oracle_py.oracle_solve(pf.batch, threads=8, non_skip_linkable=nsl, want_all=True, keep_debug=True)
        assert pu.debug_equal(got.dbg, want.dbg) is None
        assert pu.result_rows_equal(got, want) is None
        for k in ("n_pair", "n_vtx", "n_edge", "n_heap", "n_walk", "n_task"):
            assert got.stats[k] == want.stats[k], k
        assert got.stats["n_launch"] > 0


@pytest.mark.parametrize("nsl", [False, True])
@pytest.mark.parametrize("case", ["micro", "tiny", "ties", "dense", "withalt"])
def test_matches_reference_golden(case, nsl, solver, workdir):
    """The CUDA path against the golden vectors the reference itself produced (tests/golden/make_golden.py)."""
    import alignasm_b200 as aa
    from golden_util import check_against_golden, open_case
    pf = open_case(case)
    res = check_against_golden(case, nsl, solver.solve, pf, workdir)
    assert res.stats["n_launch"] > 0


def test_resident_batch_and_repeatability(solver, workdir):
    """aa_upload + aa_solve_device (batch resident in HBM) gives the same rows as aa_solve, run after run
    (the pooled workspace is reused between solves)."""
    import alignasm_b200 as aa
    args, _ = SMALL["c1_small"]
    pf = aa.read_paf(pu.synth(os.path.join(workdir, "resident.paf"), *args))
    first = solver.solve(pf.batch, want_all=True)
    dev = solver.upload(pf.batch)
    for _ in range(3):
        again = solver.solve_device(dev, want_all=True)
        assert pu.result_rows_equal(first, again) is None
    assert solver.solve_device(dev, fetch=False) is None
    assert solver.stats()["n_blk"] == pf.batch.n_blk
    dev.free()


def test_full_size_properties(solver, workdir):
    """BASELINE config 2 at full size (~500k blocks; the oracle needs minutes there): size-independent
    properties of the result, plus exact agreement with the oracle on a sample of its contigs."""
    import alignasm_b200 as aa
    from oracle import oracle_py
    pf = aa.read_paf(pu.synth(os.path.join(workdir, "c2_full.paf"), "--preset", "c2"))
    b = pf.batch
    res = solver.solve(b)
    n = np.diff(b.ctg_off)
    # every contig has a primary chain; rows reference blocks of their own contig
    assert np.all(np.diff(res.out_off) >= 1)
    for which, off in (("out", res.out_off), ("alt", res.alt_off)):
        r = getattr(res, which)
        ctg = np.repeat(np.arange(b.n_ctg), np.diff(off))
        assert np.all(r["ctg_index"] >= 0) and np.all(r["ctg_index"] < n[ctg])
        g = b.ctg_off[ctg] + r["ctg_index"]
        # trimmed intervals stay inside the block they came from and are non-empty
        assert np.all(r["qry_str"] >= b.qry_str[g]) and np.all(r["qry_end"] <= b.qry_end[g])
        assert np.all(r["qry_str"] <= r["qry_end"])
        lo, hi = np.minimum(b.ref_str[g], b.ref_end[g]), np.maximum(b.ref_str[g], b.ref_end[g])
        assert np.all(np.minimum(r["ref_str"], r["ref_end"]) >= lo) and np.all(np.maximum(r["ref_str"], r["ref_end"]) <= hi)
        # a chain is strictly increasing and non-overlapping on the query
        same = ctg[1:] == ctg[:-1]
        assert np.all(r["qry_str"][1:][same] > r["qry_end"][:-1][same])
    # the primary chain never carries the secondary flag (its blocks are marked by its own walk or an earlier one)
    # ctg_sorted_index is a permutation per contig that sorts (qry_str, qry_end)
    for c in np.random.default_rng(0).choice(b.n_ctg, 40, replace=False):
        a, e = int(b.ctg_off[c]), int(b.ctg_off[c + 1])
        si = res.sorted_index[a:e]
        assert np.array_equal(np.sort(si), np.arange(e - a))
        inv = np.empty(e - a, dtype=np.int64)
        inv[si] = np.arange(e - a)
        keys = list(zip(b.qry_str[a:e][inv].tolist(), b.qry_end[a:e][inv].tolist()))
        assert keys == sorted(keys)
    # idempotence: a second solve of the same batch gives identical rows
    assert pu.result_rows_equal(res, solver.solve(b), check_all=False) is None
    # exact agreement with the oracle on contigs it can do in seconds
    pick = np.nonzero(n <= 600)[0][:60]
    sub = b.select(pick)
    got, want = solver.solve(sub, want_all=True), oracle_py.oracle_solve(sub, threads=8, want_all=True)
    assert pu.result_rows_equal(got, want) is None
    full_rows = [res.rows_of("out", int(c)) for c in pick]
    assert full_rows == [got.rows_of("out", k) for k in range(len(pick))]


def test_cancer_full_size(solver, workdir):
    """BASELINE config 3 (cancer karyotype: dense translocation / inversion / duplicate blocks, exact ties, the
    alt loop active on most contigs) at full size: invariants on every contig + exact agreement with the oracle
    on a sample of contigs, including the tie-sensitive .alt / .all lists."""
    import alignasm_b200 as aa
    from oracle import oracle_py
    pf = aa.read_paf(pu.synth(os.path.join(workdir, "c3_full.paf"), "--preset", "c3"))
    b = pf.batch
    res = solver.solve(b, want_all=True)
    n = np.diff(b.ctg_off)
    assert np.all(np.diff(res.out_off) >= 1)
    assert res.stats["n_walk"] <= 10000 * b.n_ctg and res.stats["n_task"] >= b.n_ctg - int(np.sum(n == 1))
    for which, off in (("out", res.out_off), ("alt", res.alt_off)):
        r = getattr(res, which)
        ctg = np.repeat(np.arange(b.n_ctg), np.diff(off))
        g = b.ctg_off[ctg] + r["ctg_index"]
        assert np.all(r["ctg_index"] >= 0) and np.all(r["ctg_index"] < n[ctg])
        assert np.all(r["qry_str"] >= b.qry_str[g]) and np.all(r["qry_end"] <= b.qry_end[g])
        same = ctg[1:] == ctg[:-1]
        assert np.all(r["qry_str"][1:][same] > r["qry_end"][:-1][same])
    # the alt chain exists only where it has fewer anomalies than the primary could avoid: it is never longer than the contig
    assert np.all(np.diff(res.alt_off) <= n)
    pick = np.nonzero(n <= 500)[0][:80]
    sub = b.select(pick)
    got = solver.solve(sub, want_all=True, keep_debug=True)
    want = oracle_py.oracle_solve(sub, threads=8, want_all=True, keep_debug=True)
    assert pu.debug_equal(got.dbg, want.dbg) is None
    assert pu.result_rows_equal(got, want) is None
    assert [res.rows_of("alt", int(c)) for c in pick] == [got.rows_of("alt", k) for k in range(len(pick))]


def test_sharded_solve_equals_whole(solver, workdir):
    """Contig sharding: LPT shards solved one after the other on this GPU and merged in input
    order give exactly the rows of the unsharded solve."""
    import alignasm_b200 as aa
    from alignasm_b200 import sharding
    args, _ = SMALL["cancer_small"]
    b = aa.read_paf(pu.synth(os.path.join(workdir, "shard.paf"), *args)).batch
    whole = sharding.rows_by_contig(solver.solve(b, want_all=True))
    for world in (2, 3, 8):
        shards = sharding.lpt_shards(sharding.contig_costs(b), world)
        rows = [sharding.rows_by_contig(solver.solve(b.select(ids), want_all=True)) if len(ids) else [] for ids in shards]
        assert sharding.merge_shards(b.n_ctg, shards, rows) == whole


def test_native_sharded_solve(solver, workdir):
    """aa_solve_multi (C ABI): contigs sharded over several contexts (here all on device 0, which exercises the
    partition, the per-shard host threads and the merge) return exactly the rows of the single-device solve."""
    import alignasm_b200 as aa
    args, _ = SMALL["cancer_small"]
    b = aa.read_paf(pu.synth(os.path.join(workdir, "multi.paf"), *args)).batch
    whole = solver.solve(b, want_all=True)
    for devs in ([0, 0], [0, 0, 0, 0, 0]):
        got = aa.solve_multi(b, devs, want_all=True)
        assert pu.result_rows_equal(got, whole) is None
        for k in ("n_ctg", "n_blk", "n_pair", "n_vtx", "n_edge", "n_heap", "n_walk", "n_task"):
            assert got.stats[k] == whole.stats[k], k
    shard = aa.shard_contigs(b, 4)
    assert shard.min() == 0 and shard.max() == 3 and len(shard) == b.n_ctg


def test_walk_limit_option(solver, workdir):
    """aa_opts.max_walks (MAX_PATH_COUNT, paf_data.cpp:729): small limits exercise the early stop of the batched
    enumeration, a large one its backlog / refill / spill paths; all against the oracle."""
    import alignasm_b200 as aa
    from oracle import oracle_py
    args, _ = SMALL["ties"]
    b = aa.read_paf(pu.synth(os.path.join(workdir, "klimit.paf"), *args)).batch
    for k in (1, 2, 37, 1000, 60000):
        got = solver.solve(b, want_all=True, keep_debug=True, max_walks=k)
        want = oracle_py.oracle_solve(b, threads=8, want_all=True, keep_debug=True, max_walks=k)
        assert pu.debug_equal(got.dbg, want.dbg) is None, k
        assert pu.result_rows_equal(got, want) is None, k


def test_cli_writes_reference_bytes(product_lib, workdir):
    """`alignasm <input.paf>` (the drop-in surface) reproduces the reference's three files byte for byte."""
    import shutil
    import subprocess
    paf = os.path.join(workdir, "cli_ties.paf")
    shutil.copy(os.path.join(pu.GOLDEN, "ties.paf"), paf)
    exe = os.path.join(pu.ROOT, "alignasm_b200", "alignasm")
    out = subprocess.run([exe, paf], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "File read complete" in out.stdout and "Write output PAF file" in out.stdout
    for ext in ("aln.paf", "aln.alt.paf", "aln.all.paf"):
        assert pu.files_equal(paf[:-4] + "." + ext, os.path.join(pu.GOLDEN, "ties." + ext)), ext
    assert subprocess.run([exe, os.path.join(workdir, "nope.txt")], capture_output=True).returncode == 1
    # the same through the sharded path (two contexts on device 0)
    paf2 = os.path.join(workdir, "cli_ties2.paf")
    shutil.copy(os.path.join(pu.GOLDEN, "ties.paf"), paf2)
    out = subprocess.run([exe, "--devices", "0,0", paf2], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    for ext in ("aln.paf", "aln.alt.paf", "aln.all.paf"):
        assert pu.files_equal(paf2[:-4] + "." + ext, os.path.join(pu.GOLDEN, "ties." + ext)), ext


def test_cli_alt_ingestion(product_lib, workdir):
    """`alignasm --alt ALT.paf [--alt_baseline B] <input.paf>` (alignasm.cpp:186-332): byte-identical to the reference,
    against the committed golden (baseline 0.5) and against the oracle on a fresh seeded case (baseline 0.25)."""
    import shutil
    import subprocess
    import alignasm_b200 as aa
    import alt_util
    from oracle import oracle_py
    exe = os.path.join(pu.ROOT, "alignasm_b200", "alignasm")
    paf, alt = os.path.join(workdir, "cli_withalt.paf"), os.path.join(workdir, "cli_withalt.altin.paf")
    shutil.copy(os.path.join(pu.GOLDEN, "withalt.paf"), paf)
    shutil.copy(os.path.join(pu.GOLDEN, "withalt.altin.paf"), alt)
    out = subprocess.run([exe, "--alt", alt, paf], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    for ext in ("aln.paf", "aln.alt.paf", "aln.all.paf"):
        assert pu.files_equal(paf[:-4] + "." + ext, os.path.join(pu.GOLDEN, "withalt." + ext)), ext
    with open(paf[:-4] + ".aln.paf") as f:
        assert any("\txi:Z:A_" in line for line in f)
    assert subprocess.run([exe, "--alt", os.path.join(workdir, "nope.txt"), paf], capture_output=True).returncode == 1
    # fresh case, other baseline, --non_skip_linkable: CLI bytes == reader + oracle + writer
    paf = pu.synth(os.path.join(workdir, "cli_alt2.paf"), "--contigs", 40, "--blocks", 30, "--sd", 10, "--p_dup", 0.15,
                   "--p_trans", 0.15, "--p_inv", 0.15, "--seed", 21)
    alt = alt_util.make_alt(paf, os.path.join(workdir, "cli_alt2.altin.paf"), seed=3)
    out = subprocess.run([exe, "-a", alt, "-b", "0.25", "--non_skip_linkable", paf], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    pf = aa.read_paf(paf, alt=alt, alt_baseline=0.25)
    want = oracle_py.oracle_solve(pf.batch, threads=8, non_skip_linkable=True, want_all=True)
    pre = os.path.join(workdir, "cli_alt2_want")
    pf.write(want, pre)
    for ext in ("aln.paf", "aln.alt.paf", "aln.all.paf"):
        assert pu.files_equal(paf[:-4] + "." + ext, pre + "." + ext), pu.first_diff(paf[:-4] + "." + ext, pre + "." + ext)


@pytest.mark.parametrize("guess", ["1", "2"])
def test_segment_sweep_redo_paths(guess, solver, workdir, monkeypatch):
    """The segmented relax guesses the qul counts above each segment and checks its ratio tie-breaks afterwards
    (DESIGN.md 3.1).  AA_SEG_GUESS=1 makes the guess useless, =2 also refuses every recorded condition: tied segments are
    then redone by the sweep with the true seed — the result must not change."""
    import alignasm_b200 as aa
    from oracle import oracle_py
    args, variants = SMALL["segments"]
    pf = aa.read_paf(pu.synth(os.path.join(workdir, "segredo.paf"), *args))
    monkeypatch.setenv("AA_SEG_GUESS", guess)
    for nsl in variants:
        got = solver.solve(pf.batch, non_skip_linkable=nsl, want_all=True, keep_debug=True)
        want = oracle_py.oracle_solve(pf.batch, threads=8, non_skip_linkable=nsl, want_all=True, keep_debug=True)
        assert pu.debug_equal(got.dbg, want.dbg) is None
        assert pu.result_rows_equal(got, want) is None


@pytest.mark.parametrize("p_dup", [0, 0.05])
def test_big_contig_in_shuffled_row_order(p_dup, solver, workdir):
    """Rows of a large contig in arbitrary file order, without and with equal (qry_str, qry_end) keys: the sorted order is
    the one the reference's own unstable std::sort leaves (OC1 / H2), whatever order the file had."""
    import random
    import alignasm_b200 as aa
    from oracle import oracle_py
    src = pu.synth(os.path.join(workdir, f"shuf_src_{p_dup}.paf"), "--contigs", 2, "--blocks", 20000, "--sd", 100, "--p_dup", p_dup,
                   "--p_trans", 0.02, "--p_inv", 0.02, "--seed", 41)
    groups, order = {}, []
    with open(src) as f:
        for line in f:
            q = line.split("\t", 1)[0]
            if q not in groups:
                groups[q] = []
                order.append(q)
            groups[q].append(line)
    rng = random.Random(7)
    paf = os.path.join(workdir, f"shuf_{p_dup}.paf")
    with open(paf, "w") as f:
        for q in order:
            rng.shuffle(groups[q])
            f.writelines(groups[q])
    pf = aa.read_paf(paf)
    assert int(np.diff(pf.batch.ctg_off).max()) >= 16384
    got = solver.solve(pf.batch, want_all=False)
    want = oracle_py.oracle_solve(pf.batch, threads=8, want_all=False)
    assert pu.result_rows_equal(got, want, check_all=False) is None


def _pinned():
    from fullsize_util import PINS
    return [t for t in PINS if os.path.exists(os.path.join(pu.GOLDEN, f"fullsize_{t}.json"))]


@pytest.mark.parametrize("tag", _pinned())
def test_fullsize_pins(tag, solver, workdir):
    """The BENCH workloads themselves, at full size (C2: 260 contigs / 499 997 blocks incl. the 43 099-block contig that is the
    critical path of every step; C3: 549 724 blocks), and the dense-contig ladder of BASELINE config 4 (n = 845, 1 645, and
    3 290 with --non_skip_linkable), against the CPU restatement: tests/golden/fullsize_<tag>.json holds the sha256 of every
    result array, of the ordered edge lists, of d / best and of the walk distances the port produced for the same seed
    (tests/golden/make_fullsize.py, minutes of CPU, run in the CPU container)."""
    import json
    import alignasm_b200 as aa
    from fullsize_util import PINS, STAT_KEYS, digest
    want = json.load(open(os.path.join(pu.GOLDEN, f"fullsize_{tag}.json")))
    args, opts = PINS[tag]
    paf = pu.synth(os.path.join(workdir, f"fullsize_{tag}.paf"), *args)
    pf = aa.read_paf(paf)
    got = solver.solve(pf.batch, want_all=False, keep_debug=True, **opts)
    for k in STAT_KEYS:
        assert int(got.stats[k]) == want["stats"][k], k
    d, per = digest(got)
    big = int(np.argmax(per[0]))
    assert {"index": big, "V": int(per[0][big]), "E": int(per[1][big]), "K": int(per[2][big])} == want["largest_contig"]
    bad = [k for k in want["sha256"] if d[k] != want["sha256"][k]]
    assert not bad, f"{tag}: differs from the CPU restatement in {bad}"


@pytest.mark.parametrize("case,nsl", [("ties", False), ("ties", True), ("withalt", False), ("tiny", True)])
def test_reference_with_binding_writes_golden_bytes(case, nsl, product_lib, workdir):
    """The drop-in, compiled: oracle/_ref/alignasm_ref_b200 is the reference's own reader, get_edited_paf_data and writers
    (paf_data.cpp unmodified) with its solve loop (alignasm.cpp:346-397) replaced by the binding of INTEGRATION.md §2
    (oracle/solve_batch_b200.cpp) calling libalignasm_b200.so.  Its three output files equal the reference's goldens."""
    from oracle import oracle_py
    if oracle_py.ref_binary("b200") is None:
        pytest.skip("oracle/_ref/alignasm_ref_b200 was not built (needs /root/reference at build time)")
    tag = case + (".nsl" if nsl else "")
    alt = os.path.join(pu.GOLDEN, case + ".altin.paf")
    pre = os.path.join(workdir, "bind_" + tag)
    js = oracle_py.run_ref(os.path.join(pu.GOLDEN, case + ".paf"), pre, variant="b200", non_skip_linkable=nsl,
                           alt=alt if os.path.exists(alt) else None)
    assert js["blocks"] > 0
    for ext in ("aln.paf", "aln.alt.paf", "aln.all.paf"):
        want = os.path.join(pu.GOLDEN, tag + "." + ext)
        assert pu.files_equal(pre + "." + ext, want), f"{tag}.{ext}:\n" + pu.first_diff(pre + "." + ext, want)


def _cs_fields(paf_path):
    """(text, cs_off, cs_len, qs, qe, rs, re, fwd) of a PAF file, parsed in Python the way the reader does (alignasm.cpp:141-159)."""
    text = open(paf_path, "rb").read()
    off, ln, qs, qe, rs, re_, fwd = [], [], [], [], [], [], []
    pos = 0
    for line in text.split(b"\n"):
        if line:
            f = line.split(b"\t")
            k = line.index(b"\tcs:Z:") + 1
            end = line.find(b"\t", k)
            off.append(pos + k)
            ln.append((len(line) if end < 0 else end) - k)
            a, b = int(f[2]), int(f[3]) - 1
            c, d = int(f[7]), int(f[8]) - 1
            plus = f[4] == b"+"
            qs.append(a), qe.append(b), rs.append(c if plus else d), re_.append(d if plus else c), fwd.append(1 if plus else 0)
        pos += len(line) + 1
    return text, off, ln, qs, qe, rs, re_, fwd


@pytest.mark.parametrize("case", ["micro", "tiny", "ties", "c1"])
def test_cs_runs_on_device_equal_host_parser(case, solver, workdir):
    """parse_short_cs + get_overlap_range as CUDA kernels (csrc/cs_codec.cu) against the host codec that the
    goldens pin: same run offsets and the same (q_l, q_r, r_l) of every exact-match run, '+' and '-' rows alike."""
    import alignasm_b200 as aa
    paf = (pu.synth(os.path.join(workdir, "cs_c1.paf"), "--preset", "c1", "--scale", 0.2) if case == "c1"
           else os.path.join(pu.GOLDEN, case + ".paf"))
    want = aa.read_paf(paf).batch
    text, off, ln, qs, qe, rs, re_, fwd = _cs_fields(paf)
    assert np.array_equal(qs, want.qry_str) and np.array_equal(rs, want.ref_str) and np.array_equal(re_, want.ref_end)
    run_off, ql, qr, rl, err = aa.cs_runs_device(solver, text, off, ln, qs, qe, rs, re_, fwd)
    assert not err.any()
    assert np.array_equal(run_off, want.run_off)
    assert np.array_equal(ql, want.run_ql) and np.array_equal(qr, want.run_qr) and np.array_equal(rl, want.run_rl)
    # the reader with the device codec gives the same batch
    dev = aa.read_paf(paf, solver=solver).batch
    for name, _ in aa.Batch.FIELDS:
        assert np.array_equal(getattr(dev, name), getattr(want, name)), name


def test_cs_codec_error_sites(solver):
    """The reference's exception sites of the cs codec (paf_data.cpp:31-122) come back as per-row codes."""
    import alignasm_b200 as aa
    from alignasm_b200 import _abi  # noqa: F401
    rows = [(b"cs:Z::10", 0, 9, 100, 109, 1, 0), (b"xx:Z::10", 0, 9, 100, 109, 1, 1), (b"cs:Z::0", 0, 9, 100, 109, 1, 2),
            (b"cs:Z:*a", 0, 0, 100, 100, 1, 3), (b"cs:Z:+", 0, 9, 100, 109, 1, 4), (b"cs:Z:=ACGT", 0, 3, 100, 103, 1, 5),
            (b"cs:Z::9", 0, 9, 100, 109, 1, 6), (b"cs:Z::4*ag-tt+cc:3", 0, 9, 109, 100, 0, 0)]
    text, off = b"", []
    for r in rows:
        off.append(len(text))
        text += r[0] + b"\t"
    run_off, ql, qr, rl, err = aa.cs_runs_device(solver, text, off, [len(r[0]) for r in rows], [r[1] for r in rows], [r[2] for r in rows],
                                                  [r[3] for r in rows], [r[4] for r in rows], [r[5] for r in rows])
    assert err.tolist() == [r[6] for r in rows]
    assert "consumption" in aa.cs_error_text(6)
    # the '-' row: operations applied from the last one back (paf_data.cpp:97-117): run ":3" first, then ":4"
    a = int(run_off[7])
    assert (ql[a:].tolist(), qr[a:].tolist(), rl[a:].tolist()) == ([0, 6], [2, 9], [109, 103])


def test_cli_cs_device_writes_reference_bytes(product_lib, workdir):
    """`alignasm --cs_device`: cs:Z: parsing and re-cutting (get_edited_paf_data, paf_data.cpp:125-220) on the GPU; the three files
    are the reference's goldens byte for byte, with and without --alt."""
    import shutil
    import subprocess
    exe = os.path.join(pu.ROOT, "alignasm_b200", "alignasm")
    for case, extra in (("ties", []), ("tiny", []), ("withalt", ["--alt", os.path.join(pu.GOLDEN, "withalt.altin.paf")])):
        paf = os.path.join(workdir, f"csdev_{case}.paf")
        shutil.copy(os.path.join(pu.GOLDEN, case + ".paf"), paf)
        out = subprocess.run([exe, "--cs_device", *extra, paf], capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
        for ext in ("aln.paf", "aln.alt.paf", "aln.all.paf"):
            assert pu.files_equal(paf[:-4] + "." + ext, os.path.join(pu.GOLDEN, case + "." + ext)), (case, ext)


def test_writer_with_device_codec_on_bench_workload(solver, workdir):
    """The writers with the device cs codec against the host codec on a sizeable input (C2 at 1/5 scale: clipped rows of both
    strands, every operation type)."""
    import alignasm_b200 as aa
    paf = pu.synth(os.path.join(workdir, "csw.paf"), "--preset", "c2", "--scale", 0.2)
    pf = aa.read_paf(paf, solver=solver)
    res = solver.solve(pf.batch)
    pf.write(res, os.path.join(workdir, "csw_dev"), solver=solver)
    pf.write(res, os.path.join(workdir, "csw_host"))
    for ext in ("aln.paf", "aln.alt.paf", "aln.all.paf"):
        assert pu.files_equal(os.path.join(workdir, "csw_dev." + ext), os.path.join(workdir, "csw_host." + ext)), ext


def test_pinned_inputs_and_result_slabs(solver, workdir):
    """A batch in page-locked memory (aa_host_alloc) goes up without the staging pass and gives the same rows; results live
    in pooled pinned slabs: they stay valid after later solves and after their context is gone, and a released slab is reused."""
    import alignasm_b200 as aa
    args, _ = SMALL["c1_small"]
    pf = aa.read_paf(pu.synth(os.path.join(workdir, "pinned.paf"), *args))
    base = solver.solve(pf.batch, want_all=True)
    hb = pf.batch.pinned()
    for name, _ in aa.Batch.FIELDS:
        assert np.array_equal(getattr(hb, name), getattr(pf.batch, name))
    got = solver.solve(hb, want_all=True)
    assert pu.result_rows_equal(base, got) is None
    dev = solver.upload(hb)
    assert pu.result_rows_equal(base, solver.solve_device(dev, want_all=True)) is None
    dev.free()
    # views (copy=False) of a result stay intact while other solves run and after the solver that made them is closed
    other = aa.Solver(0)
    keep = other.solve(pf.batch, copy=False)
    snap = {k: np.array(v) for k, v in keep.out.items()}
    seen = {keep.out["qry_str"].ctypes.data}
    for _ in range(3):
        r = other.solve(pf.batch, copy=False)
        seen.add(r.out["qry_str"].ctypes.data)
        r.close()
    assert len(seen) == 2  # `keep` holds one slab, the three others took turns on a second one
    other.close()
    assert all(np.array_equal(snap[k], keep.out[k]) for k in snap)
    assert pu.result_rows_equal(base, keep, check_all=False) is None
    keep.close()
    again = solver.solve(pf.batch, copy=False)  # a slab that was given back, not a new one
    assert again.out["qry_str"].ctypes.data in seen
    assert pu.result_rows_equal(base, again, check_all=False) is None
    again.close()


@pytest.mark.parametrize("tune", ["8", "12", "1", "2", "32", "42"])
def test_enumeration_variants(tune, solver, workdir, monkeypatch):
    """The enumeration kernel exists in two instantiations (packed keys / wide keys for contigs of 2^20 vertices and more) and
    has optional parts (serial steps, expansion records, the far backlog region).  AA_TUNE forces each variant on inputs every
    variant must solve identically: 8 = wide keys, 12 = wide keys + one backlog region, 1 = no serial steps, 2 = no expansion records,
    32 = the two-group pipelining of heaps and enumeration forced on a small batch (a quarter of the contigs form the "large" group),
    42 = that with wide keys and without expansion records."""
    import alignasm_b200 as aa
    from oracle import oracle_py
    monkeypatch.setenv("AA_TUNE", tune)
    for name in ("ties", "segments", "c1_small"):
        args, variants = SMALL[name]
        pf = aa.read_paf(pu.synth(os.path.join(workdir, "tune_" + name + ".paf"), *args))
        for nsl in variants:
            got = solver.solve(pf.batch, non_skip_linkable=nsl, want_all=True, keep_debug=True)
            want = oracle_py.oracle_solve(pf.batch, threads=8, non_skip_linkable=nsl, want_all=True, keep_debug=True)
            assert pu.debug_equal(got.dbg, want.dbg) is None
            assert pu.result_rows_equal(got, want) is None
