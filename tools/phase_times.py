"""Per-phase device times of the presets (batch resident in HBM): python tools/phase_times.py c1 c2 c3"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import alignasm_b200 as aa, parity_util as pu
s = aa.Solver(0)
names = s.phase_names()
for tag in sys.argv[1:] or ["c1", "c2", "c3"]:
    args = ["--preset", "c5", "--replicas", 8] if tag == "c5x8" else ["--preset", tag]
    b = aa.read_paf(pu.synth(f"/tmp/pt_{tag}.paf", *args)).batch
    dev = s.upload(b)
    for it in range(3):
        s.solve_device(dev, fetch=False)
    st = s.stats()
    print(tag, "blocks", st["n_blk"], "contigs", st["n_ctg"], "dev %.1f ms" % st["ms_total"], "=> %.2f M blocks/s" % (st["n_blk"] / st["ms_total"] / 1e3),
          {n: round(m, 2) for n, m in zip(names, st["ms_phase"])}, {k: st[k] for k in ("n_edge", "n_heap", "n_walk", "n_task")}, flush=True)
    dev.free()
