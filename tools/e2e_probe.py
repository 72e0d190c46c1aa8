"""Where does the end-to-end time of aa_solve go? (upload / device solve / download)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import alignasm_b200 as aa, parity_util as pu
paf = pu.synth("/tmp/prof_c2.paf", "--preset", "c2")
b = aa.read_paf(paf).batch
s = aa.Solver(0)
for _ in range(2):
    s.solve(b).close()
for it in range(3):
    t0 = time.perf_counter(); dev = s.upload(b); t1 = time.perf_counter()
    r = s.solve_device(dev); t2 = time.perf_counter()
    st = r.stats
    r.close(); dev.free(); t3 = time.perf_counter()
    t4 = time.perf_counter(); r2 = s.solve(b); t5 = time.perf_counter(); r2.close()
    print(f"upload {1e3*(t1-t0):.1f} ms | solve_device+fetch {1e3*(t2-t1):.1f} ms (device {st['ms_total']:.1f}, d2h phase {st['ms_phase'][13]:.1f}, sort phase {st['ms_phase'][0]:.1f}) | free {1e3*(t3-t2):.1f} ms | aa_solve total {1e3*(t5-t4):.1f} ms")
