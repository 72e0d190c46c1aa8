import sys, numpy as np
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import parity_util as pu, alignasm_b200 as aa
from oracle import oracle_py as op
from shapes import SMALL
name = sys.argv[1] if len(sys.argv) > 1 else "overlappy"
args,_ = SMALL[name]
b = aa.read_paf(pu.synth("/tmp/case.paf", *args)).batch
s = aa.Solver(0)
got = s.solve(b, want_all=True, keep_debug=True)
want = op.oracle_solve(b, threads=8, want_all=True, keep_debug=True)
g, w = got.dbg, want.dbg
print("stats", {k:(got.stats[k], want.stats[k]) for k in ("n_heap","n_walk","n_task")})
for c in range(b.n_ctg):
    a0, a1 = int(g["walk_off"][c]), int(g["walk_off"][c+1])
    b0, b1 = int(w["walk_off"][c]), int(w["walk_off"][c+1])
    n = min(a1-a0, b1-b0)
    d = np.nonzero((g["w_sum"][a0:a0+n] != w["w_sum"][b0:b0+n]) | (g["w_qnz"][a0:a0+n] != w["w_qnz"][b0:b0+n]))[0]
    if len(d) or (a1-a0) != (b1-b0):
        i = int(d[0]) if len(d) else n
        print("contig", c, "blocks", int(b.ctg_off[c+1]-b.ctg_off[c]), "V", int(g["vtx_off"][c+1]-g["vtx_off"][c]), "walks", a1-a0, b1-b0, "first diff at", i)
        for k in range(max(0,i-2), min(n, i+4)):
            print("  ", k, "gpu", [int(g[x][a0+k]) for x in ("w_sum","w_anom","w_qnz","w_qtot")], "ref", [int(w[x][b0+k]) for x in ("w_sum","w_anom","w_qnz","w_qtot")])
        break
