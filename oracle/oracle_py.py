"""oracle/oracle_py.py — TEST INFRASTRUCTURE ONLY.

Python access to the two checkers:
  * the CPU restatement  oracle/_build/liboracle_port.so   (oracle_port.cpp; same C structs as the product ABI)
  * the reference itself oracle/_ref/alignasm_ref*          (reference sources compiled where they lie; built
                                                            only where /root/reference exists, travels as binaries)
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import json
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from alignasm_b200 import _abi  # noqa: E402  (struct layouts only)
from alignasm_b200 import Result, _opts  # noqa: E402

REF_DIR = os.path.join(_HERE, "_ref")
PORT_LIB = os.path.join(_HERE, "_build", "liboracle_port.so")
REFERENCE_SRC = "/root/reference/src"
_lib = None


def build(ref=True, quiet=True):
    """Compile the checkers (g++ only).  `ref` is skipped where the reference tree is absent."""
    targets = ["port"]
    if ref and os.path.isdir(REFERENCE_SRC):
        targets.append("ref")
    out = subprocess.run(["make", "-C", _HERE] + targets, capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if not quiet:
        print(out.stdout)


def port_lib():
    global _lib
    if _lib is None:
        if not os.path.exists(PORT_LIB):
            build(ref=False)
        lib = C.CDLL(PORT_LIB)
        lib.oracle_solve.argtypes = [C.POINTER(_abi.aa_batch), C.POINTER(_abi.aa_opts), C.POINTER(_abi.aa_result), C.c_int]
        lib.oracle_solve.restype = C.c_int
        lib.oracle_result_free.argtypes = [C.POINTER(_abi.aa_result)]
        lib.oracle_result_free.restype = None
        _lib = lib
    return _lib


def oracle_solve(batch, threads=1, **kw):
    """CPU restatement over a whole batch -> alignasm_b200.Result (same layout as the product's)."""
    lib = port_lib()
    res = _abi.aa_result()
    o = _opts(**kw)
    st = lib.oracle_solve(C.byref(batch.c_struct()), C.byref(o), C.byref(res), int(threads))
    if st != 0:
        raise RuntimeError(f"oracle_solve failed: {_abi.STATUS.get(st, st)}")
    return Result(res, batch.n_blk, lib.oracle_result_free)


def ref_binary(variant="canon"):
    name = {"canon": "alignasm_ref_canon", "glibc": "alignasm_ref", "dump": "alignasm_ref_dump",
            "dbg": "alignasm_ref_dbg", "b200": "alignasm_ref_b200"}[variant]
    p = os.path.join(REF_DIR, name)
    return p if os.path.exists(p) else None


def run_ref(paf_path, out_prefix, variant="canon", non_skip_linkable=False, threads=1, dump=None, no_write=False,
            limit_contigs=None, timeout=None, alt=None, alt_baseline=None):
    """Run the compiled reference on a PAF file; returns its timing JSON."""
    exe = ref_binary(variant)
    if exe is None:
        raise FileNotFoundError("oracle/_ref is not built (needs /root/reference)")
    cmd = [exe, paf_path, "--out-prefix", out_prefix, "-t", str(threads)]
    if non_skip_linkable:
        cmd.append("--non_skip_linkable")
    if no_write:
        cmd.append("--no-write")
    if dump:
        cmd += ["--dump", dump]
    if limit_contigs is not None:
        cmd += ["--limit-contigs", str(limit_contigs)]
    if alt is not None:
        cmd += ["--alt", alt]
    if alt_baseline is not None:
        cmd += ["--alt_baseline", repr(float(alt_baseline))]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    if out.returncode != 0:
        raise RuntimeError(f"{exe} failed ({out.returncode}): {out.stderr[-2000:]}")
    return json.loads(out.stdout.strip().splitlines()[-1])


def parse_dump(path):
    """Parse the hook dump (grammar in ref_dump_tu.cpp) into a list of per-contig dicts."""
    ctgs = []
    cur = None
    with open(path) as f:
        for line in f:
            t = line.split()
            if not t:
                continue
            k = t[0]
            if k == "C":
                cur = {"n": int(t[2]), "edges": [], "d": [], "walks": [], "recover": [], "anom_dis": None, "order": None}
                ctgs.append(cur)
            elif k == "E":
                cur["edges"].append(tuple(int(x) for x in t[1:]))
            elif k == "A":
                cur["anom_dis"] = int(t[1])
            elif k == "D":
                cur["d"].append(tuple(int(x) for x in t[1:]))
            elif k == "k":
                cur["walks"].append(tuple(int(x) for x in t[2:]))
            elif k == "O":
                cur["order"] = [int(x) for x in t[1:]]
            elif k == "W":
                cur["recover"].append((int(t[1]), [int(x) for x in t[3:]]))
    return ctgs
