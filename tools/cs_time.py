import sys, os, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import alignasm_b200 as aa, parity_util as pu
paf = pu.synth("/tmp/cst_c2.paf", "--preset", "c2")
s = aa.Solver(0)
for i in range(3):
    t=time.perf_counter(); pf = aa.read_paf(paf, solver=s); t1=time.perf_counter()-t
    t=time.perf_counter(); pfh = aa.read_paf(paf); t2=time.perf_counter()-t
    print("read device %.1f ms host %.1f ms" % (t1*1e3, t2*1e3), flush=True)
r = s.solve(pf.batch)
for i in range(3):
    t=time.perf_counter(); pf.write(r, "/tmp/cst_dev", solver=s); t1=time.perf_counter()-t
    t=time.perf_counter(); pf.write(r, "/tmp/cst_host"); t2=time.perf_counter()-t
    print("write device %.1f ms host %.1f ms" % (t1*1e3, t2*1e3), flush=True)
