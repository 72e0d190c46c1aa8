"""Single-contig check of the walk enumeration (and everything before it): solve tools/heap_lab/data/<name>.paf on the GPU, compare the
walk lists and the result rows with a reference saved by an earlier (parity-tested) build, print the phase times.
    python tools/enum_check.py save|check name [name ...]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import alignasm_b200 as aa
D = os.path.join(ROOT, "tools", "heap_lab", "data")
mode, names = sys.argv[1], sys.argv[2:]
s = aa.Solver(0)
pn = s.phase_names()
rc = 0
for name in names:
    b = aa.read_paf(os.path.join(D, name + ".paf")).batch
    r = s.solve(b, want_all=False, keep_debug=True)
    got = {k: np.asarray(r.dbg[k]) for k in ("walk_off", "w_sum", "w_anom", "w_qnz", "w_qtot")}
    for which in ("out", "alt"):
        for k in ("ctg_index", "qry_str", "qry_end", "ref_str", "ref_end", "is_alt"):
            got[which + "_" + k] = np.asarray(getattr(r, which)[k])
    ref = os.path.join(D, name + ".enumref.npz")
    if mode == "save":
        np.savez_compressed(ref, **got)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        np.savez_compressed(os.path.join(ROOT, "gpurun_out", name + ".enumref.npz"), **got)
        print(name, "reference saved", {k: v.shape for k, v in got.items() if k.startswith("w_")})
    else:
        want = np.load(ref)
        bad = [k for k in got if not np.array_equal(got[k], want[k])]
        print(name, "MISMATCH " + str(bad) if bad else "matches the reference")
        rc |= bool(bad)
    dev = s.upload(b)
    for it in range(3):
        s.solve_device(dev, fetch=False)
    st = s.stats()
    print("  ", name, "dev %.2f ms" % st["ms_total"], {n: round(m, 2) for n, m in zip(pn, st["ms_phase"]) if m >= 0.05}, flush=True)
    dev.free()
sys.exit(rc)
