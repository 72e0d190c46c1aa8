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
