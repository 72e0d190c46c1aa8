"""CPU: the C-ABI library loads, exports every symbol include/alignasm_b200.h declares, matches the ctypes
layouts, and fails loudly (no CPU fallback) when there is no CUDA device.  Host PAF codec error behaviour."""
import ctypes as C
import os
import re
import subprocess

import pytest

import parity_util as pu


def test_exports_every_declared_symbol(product_lib):
    from alignasm_b200 import _abi
    hdr = open(os.path.join(pu.ROOT, "include", "alignasm_b200.h")).read()
    declared = set(re.findall(r"\b(aa_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(_abi.EXPORTS), declared ^ set(_abi.EXPORTS)
    for sym in _abi.EXPORTS:
        assert hasattr(product_lib, sym), sym


def test_struct_layouts_match_header(workdir):
    from alignasm_b200 import _abi
    src = os.path.join(workdir, "sizes.c")
    exe = os.path.join(workdir, "sizes")
    with open(src, "w") as f:
        f.write('#include <stdio.h>\n#include "alignasm_b200.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                "sizeof(aa_batch),sizeof(aa_opts),sizeof(aa_rows),sizeof(aa_debug),sizeof(aa_stats),sizeof(aa_result),"
                "sizeof(aa_cs_rows),sizeof(aa_cs_runs),sizeof(aa_cs_edits));return 0;}\n")
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.run([cc, "-I", os.path.join(pu.ROOT, "include"), "-o", exe, src], check=True)
    got = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(t) for t in (_abi.aa_batch, _abi.aa_opts, _abi.aa_rows, _abi.aa_debug, _abi.aa_stats, _abi.aa_result,
                                  _abi.aa_cs_rows, _abi.aa_cs_runs, _abi.aa_cs_edits)]
    assert got == want


def test_no_device_fails_loudly(product_lib):
    """Without a GPU the solver must refuse to run; with one this test is vacuous (the gpu tests cover it)."""
    import torch
    import alignasm_b200 as aa
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(aa.AlignasmError) as e:
        aa.Solver(0)
    assert e.value.status == 2  # AA_ERR_NO_DEVICE
    assert "CUDA" in str(e.value)


def test_page_locked_batches_need_a_device(product_lib):
    """aa_host_alloc hands out page-locked memory or NULL; Batch.pinned() raises instead of silently keeping pageable arrays."""
    import numpy as np
    import torch
    import alignasm_b200 as aa
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    lib = aa.load_library()
    assert not lib.aa_host_alloc(1 << 20)
    lib.aa_host_free(None)
    z = np.zeros(1, np.int64)
    b = aa.Batch(ctg_off=np.array([0, 1]), qry_str=z, qry_end=z + 5, ref_str=z, ref_end=z + 5, qry_total=z + 5, ref_chr=np.zeros(1, np.int32),
                 aln_fwd=np.ones(1, np.uint8), map_qul=np.ones(1, np.uint8), run_off=np.array([0, 0]), run_ql=np.zeros(0, np.int64),
                 run_qr=np.zeros(0, np.int64), run_rl=np.zeros(0, np.int64))
    with pytest.raises(MemoryError):
        b.pinned()


def test_product_does_not_link_the_oracle(product_lib):
    import alignasm_b200 as aa
    out = subprocess.run(["ldd", aa.lib_path()], capture_output=True, text=True).stdout
    assert "oracle" not in out and "emul" not in out
    syms = subprocess.run(["nm", "-D", "--defined-only", aa.lib_path()], capture_output=True, text=True).stdout
    assert "oracle_solve" not in syms and "emul_solve" not in syms


def _read_err(paf_text, workdir, name):
    import alignasm_b200 as aa
    p = os.path.join(workdir, name + ".paf")
    with open(p, "w") as f:
        f.write(paf_text)
    with pytest.raises(aa.AlignasmError) as e:
        aa.read_paf(p)
    return e.value


ROW = "q\t5000\t100\t1100\t+\tchr1\t248956422\t1000\t2000\t1000\t1000\t60\ttp:A:P\t{cs}\n"


def test_reader_errors(product_lib, workdir):
    """The reference throws std::invalid_argument / returns 1 at these sites (paf_data.cpp:31,46,52,63,66,121;
    alignasm.cpp:165-168); the C ABI reports AA_ERR_FORMAT with the same message text."""
    assert "Missing cs:Z tag" in str(_read_err(ROW.replace("\t{cs}", ""), workdir, "nocs"))
    assert "does not match PAF coordinates" in str(_read_err(ROW.format(cs="cs:Z::999"), workdir, "short"))
    assert "Invalid :length" in str(_read_err(ROW.format(cs="cs:Z::0:1000"), workdir, "zero"))
    assert "Invalid substitution" in str(_read_err(ROW.format(cs="cs:Z::999*a"), workdir, "sub"))
    assert "Empty indel" in str(_read_err(ROW.format(cs="cs:Z::500+:500"), workdir, "indel"))
    assert "Unsupported operation" in str(_read_err(ROW.format(cs="cs:Z::500~gt12ag:500"), workdir, "intron"))
    assert _read_err("", workdir, "empty").status == 6


def test_reader_buckets_on_name_change(product_lib, workdir):
    """A query name that re-appears opens a new contig (alignasm.cpp:125-133)."""
    import alignasm_b200 as aa
    p = os.path.join(workdir, "names.paf")
    with open(p, "w") as f:
        f.write(ROW.format(cs="cs:Z::1000") + ROW.format(cs="cs:Z::1000").replace("q\t", "r\t") + ROW.format(cs="cs:Z::1000"))
    b = aa.read_paf(p).batch
    assert b.n_ctg == 3 and b.ctg_off.tolist() == [0, 1, 2, 3]
    assert b.qry_end.tolist() == [1099] * 3 and b.run_ql.tolist() == [100] * 3 and b.run_qr.tolist() == [1099] * 3


def test_minus_strand_runs(product_lib, workdir):
    """'-' rows swap ref_str/ref_end (alignasm.cpp:155-159) and walk the cs ops backwards (paf_data.cpp:74-86)."""
    import alignasm_b200 as aa
    p = os.path.join(workdir, "minus.paf")
    with open(p, "w") as f:
        f.write("q\t5000\t100\t402\t-\tchr1\t248956422\t1000\t1303\t300\t304\t60\tcs:Z::100-tt*ag:50+a:150\n")
    b = aa.read_paf(p).batch
    assert (int(b.ref_str[0]), int(b.ref_end[0])) == (1302, 1000)
    # query order: :150 +a :50 *ag -tt :100
    assert b.run_ql.tolist() == [100, 251, 302] and b.run_qr.tolist() == [249, 300, 401]
    assert b.run_rl.tolist() == [1302, 1152, 1099]


def test_alt_reader(product_lib, workdir):
    """aa_paf_read_alt: rows land behind their contig's own rows in contig coordinates; grouping by segment; errors."""
    import alignasm_b200 as aa
    main = os.path.join(workdir, "altr.paf")
    cs = "cs:Z::1000"
    with open(main, "w") as f:
        f.write(f"a\t9000\t0\t1000\t+\tchr1\t100000\t500\t1500\t1000\t1000\t60\ttp:A:P\t{cs}\n")
        f.write(f"b\t9000\t100\t1100\t+\tchr1\t100000\t500\t1500\t1000\t1000\t60\ttp:A:P\t{cs}\n")
    alt = os.path.join(workdir, "altr.altin.paf")
    seg = lambda q, qlen, s, e, n: f"{q}\t{qlen}\t{s}\t{e}\t-\tchr9\t100000\t700\t{700 + n}\t{n}\t{n}\t30\ttp:A:P\tcs:Z::{n}\n"
    with open(alt, "w") as f:
        f.write(seg("b:2001-4000", 2000, 0, 1500, 1500))    # above 0.5: taken
        f.write(seg("b:2001-4000", 2000, 1600, 1900, 300))  # below: dropped, the group already took a row
        f.write(seg("a:5001-7000", 2000, 100, 400, 300))    # group without a row above 0.5 ...
        f.write(seg("a:5001-7000", 2000, 500, 1300, 800))   # ... contributes its best row
        f.write(seg("zzz:11-2010", 2000, 0, 2000, 2000))    # unknown contig: bucket 0 (paf_map default, alignasm.cpp:258)
    pf = aa.read_paf(main, alt=alt)
    b = pf.batch
    assert b.ctg_off.tolist() == [0, 3, 5]
    assert b.qry_str.tolist() == [0, 5500, 10, 100, 2000] and b.qry_end.tolist() == [999, 6299, 2009, 1099, 3499]
    assert b.qry_total.tolist() == [9000] * 5
    assert b.aln_fwd.tolist() == [1, 0, 0, 1, 0] and b.ref_str[1] == 700 + 800 - 1 and b.ref_end[1] == 700
    assert aa.read_paf(main, alt=alt, alt_baseline=0.1).batch.n_blk == 7
    bad = os.path.join(workdir, "altbad.paf")
    with open(bad, "w") as f:
        f.write(seg("no_segment_name", 2000, 0, 1500, 1500))
    with pytest.raises(aa.AlignasmError):
        aa.read_paf(main, alt=bad)
    with pytest.raises(aa.AlignasmError):
        aa.read_paf(main, alt=os.path.join(workdir, "altr.txt"))
    empty = os.path.join(workdir, "altempty.paf")
    open(empty, "w").close()
    assert aa.read_paf(main, alt=empty).batch.n_blk == 2


def test_parallel_reader_equals_sequential(product_lib, workdir, monkeypatch):
    """The reader cuts the file into one slice per host thread: row numbers, target ids, contig buckets, runs and the
    first error in file order must be those of a sequential read; the writers give the same bytes at any thread count."""
    import numpy as np
    import alignasm_b200 as aa
    from oracle import oracle_py
    paf = pu.synth(os.path.join(workdir, "par.paf"), "--preset", "c1", "--scale", 0.08, "--seed", 9)
    assert os.path.getsize(paf) > (1 << 20)
    batches = {}
    for thr in ("1", "5", "16"):
        monkeypatch.setenv("AA_HOST_THREADS", thr)
        pf = aa.read_paf(paf)
        batches[thr] = pf.batch
        if thr != "1":
            for name, _ in aa.Batch.FIELDS:
                assert np.array_equal(getattr(batches["1"], name), getattr(batches[thr], name)), (thr, name)
        res = oracle_py.oracle_solve(pf.batch, threads=4, want_all=True)
        pf.write(res, os.path.join(workdir, "par_out" + thr))
    for ext in ("aln.paf", "aln.alt.paf", "aln.all.paf"):
        for thr in ("5", "16"):
            assert pu.files_equal(os.path.join(workdir, "par_out1." + ext), os.path.join(workdir, "par_out" + thr + "." + ext))
    # two malformed rows, one per half of the file: every thread count reports the first one
    lines = open(paf).read().split("\n")
    n = len(lines)
    lines[n // 4] = lines[n // 4].replace("cs:Z:", "cs:Z:~", 1)
    cols = lines[3 * n // 4].split("\t")
    cols[2] = "x"
    lines[3 * n // 4] = "\t".join(cols)
    bad = os.path.join(workdir, "par_bad.paf")
    open(bad, "w").write("\n".join(lines))
    msgs = set()
    for thr in ("1", "7"):
        monkeypatch.setenv("AA_HOST_THREADS", thr)
        with pytest.raises(aa.AlignasmError) as e:
            aa.read_paf(bad)
        msgs.add(str(e.value))
    assert len(msgs) == 1 and "Unsupported operation" in msgs.pop()


def test_integration_snippet_is_the_compiled_binding():
    """INTEGRATION.md §2 is not prose: the snippet is oracle/solve_batch_b200.cpp, which oracle/Makefile compiles against the
    reference's own headers and links into oracle/_ref/alignasm_ref_b200 (run on the GPU by tests/test_gpu_parity.py)."""
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    md = open(os.path.join(root, "INTEGRATION.md")).read()
    m = re.search(r"```cpp\n(// src/solve_batch_b200\.cpp.*?)```", md, re.S)
    assert m, "INTEGRATION.md lost its binding snippet"
    src = open(os.path.join(root, "oracle", "solve_batch_b200.cpp")).read()
    assert src.endswith(m.group(1)), "oracle/solve_batch_b200.cpp and the INTEGRATION.md snippet differ"
    exe = os.path.join(root, "oracle", "_ref", "alignasm_ref_b200")
    if os.path.exists(exe):  # built where /root/reference exists: it must link the product library, nothing of the oracle
        out = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
        assert "libalignasm_b200.so" in out and "liboracle" not in out
