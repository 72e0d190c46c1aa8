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


# `segments` (1 500-block contigs: several parts, several relax segments, ~126 MB of reference tables per contig) and `chain3000`
# (one 3 000-block chain-like contig, the shape of the bench's big contigs) pin the port to the real reference beyond small shapes
@pytest.mark.parametrize("name", ["c1_small", "ties", "overlappy", "tiny", "dense200", "segments", "chain3000"])
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


@pytest.mark.parametrize("baseline", [0.5, 0.25, 0.9])
def test_alt_ingestion_equals_reference(baseline, product_lib, workdir):
    """`--alt` (alignasm.cpp:186-332): the product's reader + merge, solved by the port, against the reference run with
    the same alternative PAF — output files byte-identical (xi:Z:A_ rows included)."""
    import alignasm_b200 as aa
    import alt_util
    op = _ref()
    paf = pu.synth(os.path.join(workdir, "alt_main.paf"), "--contigs", 40, "--blocks", 30, "--sd", 10, "--p_dup", 0.15,
                   "--p_trans", 0.15, "--p_inv", 0.15, "--seed", 21)
    alt = alt_util.make_alt(paf, os.path.join(workdir, "alt_main.altin.paf"), seed=3)
    pf = aa.read_paf(paf, alt=alt, alt_baseline=baseline)
    assert pf.batch.n_blk > aa.read_paf(paf).batch.n_blk
    pre = os.path.join(workdir, f"alt_{baseline}")
    pf.write(op.oracle_solve(pf.batch, threads=4, want_all=True), pre + "_port")
    op.run_ref(paf, pre + "_ref", variant="canon", alt=alt, alt_baseline=baseline)
    for ext in ("aln.paf", "aln.alt.paf", "aln.all.paf"):
        assert pu.files_equal(pre + "_port." + ext, pre + "_ref." + ext), pu.first_diff(pre + "_port." + ext, pre + "_ref." + ext)


def test_reference_asserts_hold_on_generator_output(workdir):
    """The assert-enabled reference build accepts the generator's PAF (cs consistent, cuts possible, DAG)."""
    op = _ref()
    if op.ref_binary("dbg") is None:
        pytest.skip("debug variant not built")
    paf = pu.synth(os.path.join(workdir, "dbg.paf"), "--preset", "c1", "--scale", 0.02, "--seed", 77)
    op.run_ref(paf, os.path.join(workdir, "dbg_out"), variant="dbg")
