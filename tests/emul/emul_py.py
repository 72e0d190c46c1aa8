"""tests/emul/emul_py.py — TEST INFRASTRUCTURE ONLY: build + load the host-loop instantiation of the product's
phase functions (see emul_backend.cpp).  Not reachable from the alignasm_b200 package."""
import ctypes as C
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
from alignasm_b200 import _abi, Result, _opts  # noqa: E402

LIB = os.path.join(_HERE, "_build", "libaa_emul.so")
_lib = None


def build():
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    src = os.path.join(_HERE, "emul_backend.cpp")
    deps = [src] + [os.path.join(_ROOT, "alignasm_b200", "csrc", f) for f in ("aa_core.cuh", "aa_pipeline.cuh")]
    if os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return
    subprocess.run([cxx, "-std=c++17", "-O2", "-Wall", "-Wno-unused-function", "-shared", "-fPIC", "-pthread", "-x", "c++",
                    "-o", LIB, src], check=True)


def emul_solve(batch, **kw):
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.emul_solve.argtypes = [C.POINTER(_abi.aa_batch), C.POINTER(_abi.aa_opts), C.POINTER(_abi.aa_result)]
        _lib.emul_solve.restype = C.c_int
        _lib.emul_result_free.argtypes = [C.POINTER(_abi.aa_result)]
        _lib.emul_result_free.restype = None
    res = _abi.aa_result()
    o = _opts(**kw)
    st = _lib.emul_solve(C.byref(batch.c_struct()), C.byref(o), C.byref(res))
    if st != 0:
        raise RuntimeError(f"emul_solve failed: {_abi.STATUS.get(st, st)}")
    return Result(res, batch.n_blk, _lib.emul_result_free)
