import sys, time, os
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import parity_util as pu, alignasm_b200 as aa
from oracle import oracle_py as op
s = aa.Solver(0)
names = s.phase_names()
for n, chk in ((400, True), (845, True), (1645, False)):
    paf = pu.synth(f"/tmp/c4_{n}.paf", "--preset", "c4", "--n", n)
    b = aa.read_paf(paf).batch
    for nsl in (False, True):
        try:
            t=time.time(); r = s.solve(b, want_all=True, non_skip_linkable=nsl); dt=time.time()-t
        except Exception as e:
            print(n, nsl, "GPU failed:", e); continue
        st = r.stats
        print(n, "nsl", nsl, "gpu wall %.1f ms dev %.1f ms" % (dt*1e3, st["ms_total"]), {k: st[k] for k in ("n_pair","n_vtx","n_edge","n_heap","n_task")}, {nm: round(m,1) for nm,m in zip(names, st["ms_phase"]) if m > 1}, flush=True)
        if chk and (n <= 400 or not nsl or True):
            t=time.time(); wr = op.oracle_solve(b, threads=1, want_all=True, non_skip_linkable=nsl); dt=time.time()-t
            print("   oracle %.1fs equal:" % dt, pu.result_rows_equal(r, wr), flush=True)
