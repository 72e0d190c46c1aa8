"""CPU: the product's own phase functions (alignasm_b200/csrc/aa_core.cuh + aa_pipeline.cuh), instantiated with a
host-loop backend by tests/emul (test-only build, not part of libalignasm_b200.so), against the golden vectors and
against the oracle on seeded synthetic inputs.  This checks the kernels' logic on a box without a GPU; the GPU
parity tests (-m gpu) check the compiled CUDA path itself."""
import os

import pytest

import parity_util as pu
from golden_util import CASES, check_against_golden, open_case
from shapes import SMALL


@pytest.mark.parametrize("nsl", [False, True])
@pytest.mark.parametrize("case", CASES)
def test_core_matches_reference_golden(case, nsl, product_lib, workdir):
    import alignasm_b200 as aa
    import emul_py
    pf = open_case(case)
    check_against_golden(case, nsl, emul_py.emul_solve, pf, workdir)


@pytest.mark.parametrize("name", ["c1_small", "c2_small", "overlappy", "singletons", "dense200", "cancer_small"])
def test_core_matches_oracle(name, product_lib, workdir):
    import alignasm_b200 as aa
    import emul_py
    from oracle import oracle_py
    args, variants = SMALL[name]
    pf = aa.read_paf(pu.synth(os.path.join(workdir, "e_" + name + ".paf"), *args))
    for nsl in variants:
        got = emul_py.emul_solve(pf.batch, non_skip_linkable=nsl, want_all=True, keep_debug=True)
        want = oracle_py.oracle_solve(pf.batch, threads=4, non_skip_linkable=nsl, want_all=True, keep_debug=True)
        assert pu.debug_equal(got.dbg, want.dbg) is None
        assert pu.result_rows_equal(got, want) is None
        for k in ("n_pair", "n_vtx", "n_edge", "n_heap", "n_walk", "n_task"):
            assert got.stats[k] == want.stats[k], k


def test_small_walk_budget(product_lib, workdir):
    """max_walks below the reference's 10000 exercises the early-stop branch of the enumeration."""
    import alignasm_b200 as aa
    import emul_py
    from oracle import oracle_py
    pf = aa.read_paf(os.path.join(pu.GOLDEN, "ties.paf"))
    for k in (1, 2, 7, 100):
        got = emul_py.emul_solve(pf.batch, max_walks=k, want_all=True, keep_debug=True)
        want = oracle_py.oracle_solve(pf.batch, max_walks=k, want_all=True, keep_debug=True)
        assert pu.debug_equal(got.dbg, want.dbg) is None
        assert pu.result_rows_equal(got, want) is None


def test_writer_streams_large_all_lists(product_lib, workdir, monkeypatch):
    """aa_paf_write buffers at most a window of contigs and streams a large .aln.all.paf list path by path when the contig's turn
    comes (the reference streams its rows too, alignasm.cpp:456-482): forced here for every contig, files must not change."""
    import emul_py
    monkeypatch.setenv("AA_WRITE_STREAM_ROWS", "0")
    for case in ("ties", "tiny"):
        pf = open_case(case)
        check_against_golden(case, False, emul_py.emul_solve, pf, workdir)
