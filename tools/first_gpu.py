import sys, time, os, json
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import alignasm_b200 as aa, parity_util as pu
from oracle import oracle_py as op
s = aa.Solver(0)
names = s.phase_names()
os.makedirs('gpurun_out', exist_ok=True)
for tag, args in [("c1", ["--preset","c1"]), ("c2", ["--preset","c2"])]:
    paf = pu.synth(f"/tmp/{tag}.paf", *args)
    t=time.time(); pf = aa.read_paf(paf); b = pf.batch; tr=time.time()-t
    dev = s.upload(b)
    for it in range(3):
        t=time.time(); r = s.solve_device(dev); tt=time.time()-t
        st = r.stats
        print(tag, "iter", it, "blocks", st["n_blk"], "wall %.1f ms dev %.1f ms" % (tt*1e3, st["ms_total"]), {n: round(m,2) for n,m in zip(names, st["ms_phase"])}, flush=True)
    print({k: st[k] for k in ("n_ctg","n_pair","n_vtx","n_edge","n_heap","n_walk","n_task","n_launch")}, "read %.2fs" % tr, flush=True)
    if tag == "c1":
        t=time.time(); w = op.oracle_solve(b, threads=16); print("oracle port 16 threads %.2fs" % (time.time()-t), pu.result_rows_equal(r, w, check_all=False))
