"""Where the time of a one-shot aa_solve (host buffers in, rows out) goes: wall clock vs the device phases, pageable vs page-locked inputs.
    python tools/e2e_times.py [c2|c5x8]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import alignasm_b200 as aa, parity_util as pu
which = sys.argv[1] if len(sys.argv) > 1 else "c2"
args = ("--preset", "c5", "--replicas", 8) if which == "c5x8" else ("--preset", which)
b = aa.read_paf(pu.synth("/tmp/e2e_%s.paf" % which, *args)).batch
s = aa.Solver(0); names = s.phase_names()
for label, batch in (("pageable", b), ("page-locked", b.pinned())):
    for _ in range(3):
        s.solve(batch, copy=False).close()
    t = time.perf_counter(); n = 5
    for _ in range(n):
        s.solve(batch, copy=False).close()
    wall = (time.perf_counter() - t) / n * 1e3
    st = s.stats()
    print("%s %s: wall %.1f ms, device events %.1f ms" % (which, label, wall, st["ms_total"]), {k: round(m, 2) for k, m in zip(names, st["ms_phase"]) if m >= 0.3}, flush=True)
dev = s.upload(b)
for _ in range(3):
    s.solve_device(dev, copy=False).close()
t = time.perf_counter()
for _ in range(5):
    s.solve_device(dev, copy=False).close()
print("%s resident: wall %.1f ms, device events %.1f ms" % (which, (time.perf_counter() - t) / 5 * 1e3, s.stats()["ms_total"]))
