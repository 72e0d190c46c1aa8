import sys, os, time
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import alignasm_b200 as aa, parity_util as pu
s = aa.Solver(0); names = s.phase_names()
for tag in ("c2","c3"):
    b = aa.read_paf(pu.synth(f"/tmp/wa_{tag}.paf", "--preset", tag)).batch
    dev = s.upload(b)
    s.solve_device(dev, fetch=False)
    for wa in (False, True):
        t=time.perf_counter(); r = s.solve_device(dev, want_all=wa); dt=time.perf_counter()-t
        st = r.stats
        print(tag, "want_all", wa, "wall %.0f ms dev %.0f ms" % (dt*1e3, st["ms_total"]), "all rows", len(r.all["qry_str"]), "paths", int(r.all_path_off[-1]), {n: round(m,1) for n,m in zip(names, st["ms_phase"]) if m > 3}, flush=True)
        r.close()
