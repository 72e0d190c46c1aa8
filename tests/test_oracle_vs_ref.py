"""CPU, only where oracle/_ref exists (this container: built from /root/reference by oracle/Makefile): the oracle
port against the reference itself on fresh seeded inputs — output files byte-identical (canonical allocator) and
ordered edge lists / d / best / forward order / walk distances equal to the hook dump."""
import os

import pytest

import parity_util as pu
from shapes import SMALL


def _ref():
    from oracle import oracle_py
    return oracle_py if oracle_py.ref_binary("canon") and oracle_py.ref_binary("dump") else None


pytestmark = pytest.mark.skipif(_ref() is None, reason="oracle/_ref not built (no /root/reference on this box)")


@pytest.mark.parametrize("name", ["c1_small", "ties", "overlappy", "tiny", "dense200"])
def test_port_equals_reference(name, product_lib, workdir):
    import alignasm_b200 as aa
    op = _ref()
    args, variants = SMALL[name]
    paf = pu.synth(os.path.join(workdir, "r_" + name + ".paf"), *args)
    pf = aa.read_paf(paf)
    for nsl in variants:
        pre = os.path.join(workdir, "r_" + name + ("_nsl" if nsl else ""))
        res = op.oracle_solve(pf.batch, threads=4, non_skip_linkable=nsl, want_all=True, keep_debug=True)
        pf.write(res, pre + "_port")
        op.run_ref(paf, pre + "_ref", variant="dump", non_skip_linkable=nsl, dump=pre + ".dump")
        for ext in ("aln.paf", "aln.alt.paf", "aln.all.paf"):
            assert pu.files_equal(pre + "_port." + ext, pre + "_ref." + ext), pu.first_diff(pre + "_port." + ext, pre + "_ref." + ext)
        assert pu.debug_vs_dump(res.dbg, op.parse_dump(pre + ".dump")) is None


def test_reference_asserts_hold_on_generator_output(workdir):
    """The assert-enabled reference build accepts the generator's PAF (cs consistent, cuts possible, DAG)."""
    op = _ref()
    if op.ref_binary("dbg") is None:
        pytest.skip("debug variant not built")
    paf = pu.synth(os.path.join(workdir, "dbg.paf"), "--preset", "c1", "--scale", 0.02, "--seed", 77)
    op.run_ref(paf, os.path.join(workdir, "dbg_out"), variant="dbg")
