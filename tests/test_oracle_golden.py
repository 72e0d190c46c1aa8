"""CPU: the oracle (oracle/oracle_port.cpp) against the golden vectors produced by the reference itself
(tests/golden/make_golden.py), byte-for-byte outputs plus graph / d / best / order / walk lists."""
import os

import pytest

import parity_util as pu
from golden_util import CASES, check_against_golden, open_case


@pytest.mark.parametrize("nsl", [False, True])
@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_reference_golden(case, nsl, product_lib, workdir):
    import alignasm_b200 as aa
    from oracle import oracle_py
    pf = open_case(case)
    check_against_golden(case, nsl, oracle_py.oracle_solve, pf, workdir)


def test_oracle_threads_agree(product_lib):
    import alignasm_b200 as aa
    from oracle import oracle_py
    pf = aa.read_paf(os.path.join(pu.GOLDEN, "ties.paf"))
    a = oracle_py.oracle_solve(pf.batch, threads=1, want_all=True)
    b = oracle_py.oracle_solve(pf.batch, threads=4, want_all=True)
    assert pu.result_rows_equal(a, b) is None


def test_fullsize_pin_files_present():
    """The full-size pins of the bench workloads (made by tests/golden/make_fullsize.py) are committed and carry every digest the
    GPU test compares."""
    import json
    for tag, blocks in (("c2", 499997), ("c3", 549724)):
        want = json.load(open(os.path.join(pu.GOLDEN, f"fullsize_{tag}.json")))
        assert want["stats"]["n_blk"] == blocks and want["stats"]["n_walk"] == 2600000
        assert {"out.qry_str", "alt.ref_end", "sorted_index", "dbg.w_sum", "dbg.edges", "dbg.d"} <= set(want["sha256"])
        assert all(len(v) == 64 for v in want["sha256"].values())
