"""Two-GPU check of aa_solve_multi: same rows as one device, and the wall time of both."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import alignasm_b200 as aa, parity_util as pu
import torch
n = torch.cuda.device_count()
b = aa.read_paf(pu.synth("/tmp/mc3.paf", "--preset", "c3")).batch
s = aa.Solver(0)
s.solve(b).close()
t = time.perf_counter(); one = s.solve(b, want_all=True); t1 = time.perf_counter() - t
devs = list(range(n))
aa.solve_multi(b, devs).close()
t = time.perf_counter(); many = aa.solve_multi(b, devs, want_all=True); t2 = time.perf_counter() - t
print("devices", devs, "one device %.1f ms, sharded %.1f ms (incl. context creation), equal rows:" % (t1 * 1e3, t2 * 1e3), pu.result_rows_equal(one, many))
