"""Random shapes through the CUDA path against the oracle (rows + graph / d / best / order / walks), both modes.
python tools/fuzz_gpu.py [seconds] [first seed]      AA_SEG_GUESS=1|2 forces the sweep's redo path of the segmented relax."""
import os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import alignasm_b200 as aa, parity_util as pu
from oracle import oracle_py
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
s = aa.Solver(0)
t0, n, bad = time.time(), 0, 0
while time.time() - t0 < budget:
    rng = random.Random(seed0 + n)
    blocks = rng.choice([40, 300, 700, 1500, 3000, 6000])
    args = ["--contigs", rng.randint(1, 6), "--blocks", blocks, "--sd", blocks // rng.choice([3, 5, 20]), "--p_dup", rng.choice([0, 0, 0.05, 0.2]),
            "--p_trans", rng.choice([0.01, 0.1, 0.3]), "--p_inv", rng.choice([0.01, 0.1, 0.3]), "--p_ovl", rng.choice([0.2, 0.4, 0.6]),
            "--p_cont", rng.choice([0, 0.05, 0.15]), "--seed", seed0 + n]
    if rng.random() < 0.3:
        args += ["--lmin", 500, "--lmax", 5000, "--gap_max", rng.choice([50, 200, 2000])]
    paf = pu.synth(f"/tmp/fuzz_{n % 4}.paf", *args)
    b = aa.read_paf(paf).batch
    for nsl in (False, True):
        try:
            got = s.solve(b, non_skip_linkable=nsl, want_all=True, keep_debug=True)
            want = oracle_py.oracle_solve(b, threads=8, non_skip_linkable=nsl, want_all=True, keep_debug=True)
            msg = pu.result_rows_equal(got, want) or pu.debug_equal(got.dbg, want.dbg)
        except Exception as e:  # noqa: BLE001
            msg = "exception: " + str(e)[:200]
        if msg:
            bad += 1
            print("MISMATCH", args, "nsl", nsl, msg, flush=True)
    n += 1
print(f"fuzz: {n} inputs x 2 modes in {time.time() - t0:.0f} s, {bad} mismatches (AA_SEG_GUESS={os.environ.get('AA_SEG_GUESS', '0')})")
sys.exit(1 if bad else 0)
