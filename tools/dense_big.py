import sys, time
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import parity_util as pu, alignasm_b200 as aa
s = aa.Solver(0); names = s.phase_names()
for n in (1645, 3290):
    b = aa.read_paf(pu.synth(f"/tmp/c4_{n}.paf", "--preset", "c4", "--n", n)).batch
    for nsl in (False, True):
        try:
            t=time.time(); r = s.solve(b, non_skip_linkable=nsl); dt=time.time()-t
            st = r.stats
            print(n, "nsl", nsl, "wall %.0f ms" % (dt*1e3), {k: st[k] for k in ("n_vtx","n_edge","n_heap")}, {nm: round(m,1) for nm,m in zip(names, st["ms_phase"]) if m > 5}, flush=True)
        except Exception as e:
            print(n, nsl, "failed:", str(e)[:200], flush=True)
