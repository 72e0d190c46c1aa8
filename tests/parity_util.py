"""Shared helpers for the parity tests: synthetic PAF generation, file comparison, dump comparison."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
SYNTH = os.path.join(ROOT, "alignasm_b200", "synth_paf")
HOSTCXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def synth_binary():
    if not os.path.exists(SYNTH):
        src = os.path.join(ROOT, "tools", "synth_paf.cpp")
        subprocess.run([HOSTCXX, "-O2", "-std=c++17", "-o", SYNTH, src], check=True)
    return SYNTH


def synth(out_path, *args):
    """Run the deterministic generator (tools/synth_paf.cpp)."""
    subprocess.run([synth_binary(), *[str(a) for a in args], "-o", out_path], check=True, capture_output=True)
    return out_path


def files_equal(a, b):
    with open(a, "rb") as fa, open(b, "rb") as fb:
        return fa.read() == fb.read()


def first_diff(a, b, limit=3):
    with open(a) as fa, open(b) as fb:
        la, lb = fa.readlines(), fb.readlines()
    out = []
    for i, (x, y) in enumerate(zip(la, lb)):
        if x != y:
            out.append(f"line {i}:\n  A {x[:200]}\n  B {y[:200]}")
            if len(out) >= limit:
                break
    if len(la) != len(lb):
        out.append(f"line counts differ: {len(la)} vs {len(lb)}")
    return "\n".join(out)


def result_rows_equal(ra, rb, check_all=True):
    """Compare two alignasm_b200.Result objects bit for bit; returns a message or None."""
    import numpy as np
    for name in ("out_off", "alt_off", "sorted_index"):
        if not np.array_equal(getattr(ra, name), getattr(rb, name)):
            bad = np.nonzero(getattr(ra, name)[: len(getattr(rb, name))] != getattr(rb, name)[: len(getattr(ra, name))])[0]
            return f"{name} differs (first at {bad[:5]})"
    for which in ("out", "alt"):
        for k in ("ctg_index", "qry_str", "qry_end", "ref_str", "ref_end", "is_alt"):
            if not np.array_equal(getattr(ra, which)[k], getattr(rb, which)[k]):
                bad = np.nonzero(getattr(ra, which)[k] != getattr(rb, which)[k])[0]
                return f"{which}.{k} differs at rows {bad[:5]}"
    if check_all:
        if not np.array_equal(ra.all_path_off, rb.all_path_off) or not np.array_equal(ra.all_row_off, rb.all_row_off):
            return "all offsets differ"
        for k in ("ctg_index", "qry_str", "qry_end", "ref_str", "ref_end", "is_alt"):
            if not np.array_equal(ra.all[k], rb.all[k]):
                return f"all.{k} differs"
    return None


def debug_equal(da, db):
    """Compare two Result.dbg dicts (graph / d / best / order / walks / anom)."""
    import numpy as np
    for k in ("vtx_off", "edge_off", "e_src", "e_dst", "e_qry", "e_ref", "e_anom", "e_qnz", "e_qtot", "anom_dis",
              "d_reach", "d_sum", "d_anom", "d_qnz", "d_qtot", "best", "order", "walk_off", "w_sum", "w_anom", "w_qnz",
              "w_qtot"):
        if not np.array_equal(da[k], db[k]):
            if len(da[k]) != len(db[k]):
                return f"dbg.{k}: length {len(da[k])} vs {len(db[k])}"
            bad = np.nonzero(da[k] != db[k])[0]
            return f"dbg.{k} differs at {bad[:5]}: {da[k][bad[:5]]} vs {db[k][bad[:5]]}"
    return None


def debug_vs_dump(dbg, dump):
    """Compare a Result.dbg (port or product) with the reference hook dump (oracle_py.parse_dump)."""
    import numpy as np
    ci = 0
    for c, ctg in enumerate(dump):
        if ctg["n"] == 1:
            continue  # singleton shortcut: the solver is never constructed (paf_data.cpp:235-239)
        e0, e1 = int(dbg["edge_off"][c]), int(dbg["edge_off"][c + 1])
        v0, v1 = int(dbg["vtx_off"][c]), int(dbg["vtx_off"][c + 1])
        w0, w1 = int(dbg["walk_off"][c]), int(dbg["walk_off"][c + 1])
        mine = list(zip(dbg["e_src"][e0:e1].tolist(), dbg["e_dst"][e0:e1].tolist(), dbg["e_qry"][e0:e1].tolist(),
                        dbg["e_ref"][e0:e1].tolist(), dbg["e_anom"][e0:e1].tolist(), dbg["e_qnz"][e0:e1].tolist(),
                        dbg["e_qtot"][e0:e1].tolist()))
        if mine != ctg["edges"]:
            for k, (a, b) in enumerate(zip(mine, ctg["edges"])):
                if a != b:
                    return f"contig {c}: edge {k} {a} vs ref {b}"
            return f"contig {c}: edge count {len(mine)} vs ref {len(ctg['edges'])}"
        if int(dbg["anom_dis"][c]) != ctg["anom_dis"]:
            return f"contig {c}: anom_dis {dbg['anom_dis'][c]} vs {ctg['anom_dis']}"
        if v1 - v0 != len(ctg["d"]):
            return f"contig {c}: V {v1 - v0} vs {len(ctg['d'])}"
        for v, (vid, reach, q, r, an, nz, tot, best) in enumerate(ctg["d"]):
            g = v0 + v
            if bool(dbg["d_reach"][g]) != bool(reach):
                return f"contig {c}: reach[{v}]"
            if reach and (int(dbg["d_sum"][g]), int(dbg["d_anom"][g]), int(dbg["d_qnz"][g]), int(dbg["d_qtot"][g])) != (q + r, an, nz, tot):
                return f"contig {c}: d[{v}]"
            if int(dbg["best"][g]) != best:
                return f"contig {c}: best[{v}] {dbg['best'][g]} vs {best}"
        order = np.empty(v1 - v0, dtype=np.int64)
        order[np.asarray(ctg["order"])] = np.arange(v1 - v0)
        if not np.array_equal(order, dbg["order"][v0:v1]):
            return f"contig {c}: forward topological order"
        walks = [(q + r, an, nz, tot) for (q, r, an, nz, tot) in ctg["walks"]]
        mine_w = list(zip(dbg["w_sum"][w0:w1].tolist(), dbg["w_anom"][w0:w1].tolist(), dbg["w_qnz"][w0:w1].tolist(),
                          dbg["w_qtot"][w0:w1].tolist()))
        if walks != mine_w:
            for k, (a, b) in enumerate(zip(mine_w, walks)):
                if a != b:
                    return f"contig {c}: walk {k} {a} vs ref {b}"
            return f"contig {c}: walk count {len(mine_w)} vs {len(walks)}"
        ci += 1
    return None
