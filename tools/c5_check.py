"""BASELINE config 5 (8 x cancer PAF, ~4.4 M blocks, 2080 contigs) on the visible GPUs: one device, then sharded."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import alignasm_b200 as aa, parity_util as pu
import torch
t = time.perf_counter(); paf = pu.synth("/tmp/c5.paf", "--preset", "c5", "--replicas", 8); tg = time.perf_counter() - t
t = time.perf_counter(); b = aa.read_paf(paf).batch; tr = time.perf_counter() - t
print("c5:", b.n_ctg, "contigs", b.n_blk, "blocks; generate %.1fs read+cs parse %.1fs" % (tg, tr), flush=True)
s = aa.Solver(0); names = s.phase_names()
s.solve(b).close()
s.solve(b).close()
t = time.perf_counter(); one = s.solve(b, copy=False); t1 = time.perf_counter() - t
st = one.stats
print("1 GPU: wall %.0f ms, device %.0f ms => %.2f M blocks/s e2e" % (t1 * 1e3, st["ms_total"], b.n_blk / t1 / 1e6), {n: round(m, 1) for n, m in zip(names, st["ms_phase"]) if m > 2}, flush=True)
n = torch.cuda.device_count()
if n > 1:
    devs = list(range(n))
    aa.solve_multi(b, devs).close()
    aa.solve_multi(b, devs).close()
    t = time.perf_counter(); many = aa.solve_multi(b, devs, copy=False); t2 = time.perf_counter() - t
    print(n, "GPUs (aa_solve_multi, warm contexts): wall %.0f ms => %.2f M blocks/s; rows equal:" % (t2 * 1e3, b.n_blk / t2 / 1e6), pu.result_rows_equal(one, many, check_all=False), flush=True)
    sh = aa.shard_contigs(b, n)
    print("shard sizes (blocks):", [int(np.diff(b.ctg_off)[sh == k].sum()) for k in range(n)])
