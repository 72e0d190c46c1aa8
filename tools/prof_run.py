"""Small driver for ncu: stage one workload in HBM and run aa_solve_device N times (default: c2, 2 solves)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import alignasm_b200 as aa, parity_util as pu
preset = sys.argv[1] if len(sys.argv) > 1 else "c2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
paf = pu.synth(f"/tmp/prof_{preset}.paf", "--preset", preset)
s = aa.Solver(0)
dev = s.upload(aa.read_paf(paf).batch)
for _ in range(n):
    s.solve_device(dev, fetch=False)
st = s.stats()
print({k: round(v, 2) for k, v in zip(s.phase_names(), st["ms_phase"])}, st["ms_total"])
