import sys, numpy as np
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import parity_util as pu, alignasm_b200 as aa
from oracle import oracle_py as op
from shapes import SMALL
args,_ = SMALL["c1_small"]
b = aa.read_paf(pu.synth("/tmp/c1s.paf", *args)).batch
s = aa.Solver(0)
got = s.solve(b, want_all=True, keep_debug=True)
want = op.oracle_solve(b, threads=8, want_all=True, keep_debug=True)
g, w = got.dbg, want.dbg
for k in ("vtx_off","edge_off","d_sum","best","order"):
    print(k, np.array_equal(g[k], w[k]))
print("stats", {k:(got.stats[k], want.stats[k]) for k in ("n_heap","n_walk","n_task")})
nw_g = np.diff(g["walk_off"]); nw_w = np.diff(w["walk_off"])
bad = np.nonzero(nw_g != nw_w)[0]
print("contigs with different walk counts", bad[:10], nw_g[bad[:10]], nw_w[bad[:10]])
c = int(bad[0]) if len(bad) else 0
a0, a1 = int(g["walk_off"][c]), int(w["walk_off"][c])
n = min(nw_g[c], nw_w[c])
for k in ("w_sum","w_anom","w_qnz","w_qtot"):
    d = np.nonzero(g[k][a0:a0+n] != w[k][a1:a1+n])[0]
    print(k, "first diff", d[:3])
print("contig", c, "V", int(g["vtx_off"][c+1]-g["vtx_off"][c]))
for i in range(18, 32):
    print(i, "gpu", [int(g[k][a0+i]) for k in ("w_sum","w_anom","w_qnz","w_qtot")], "ref", [int(w[k][a1+i]) for k in ("w_sum","w_anom","w_qnz","w_qtot")])
