"""BASELINE config 4 ladder: one dense contig of n blocks (every block overlaps the next 10-50), default mode and
--non_skip_linkable, on one B200.  Prints one JSON object per rung (sizes, device ms, phase times, or where it stops and why).
    python tools/c4_ladder.py [n[:nsl] ...]      e.g.  845 845:nsl 1645 1645:nsl 3290:nsl 3290 10000:nsl"""
import json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def one(n, nsl):
    import alignasm_b200 as aa, parity_util as pu
    paf = pu.synth(f"/tmp/c4_{n}.paf", "--preset", "c4", "--n", n)
    b = aa.read_paf(paf).batch
    s = aa.Solver(0)
    out = {"n": n, "mode": "non_skip_linkable" if nsl else "default"}
    try:
        s.solve(b, non_skip_linkable=nsl).close()  # warm-up (pool growth)
        t = time.perf_counter()
        r = s.solve(b, non_skip_linkable=nsl)
        out["wall_ms"] = round((time.perf_counter() - t) * 1e3, 1)
        st = r.stats
        out.update({"device_ms": round(st["ms_total"], 1), "V": st["n_vtx"], "E": st["n_edge"], "H": st["n_heap"], "walks": st["n_walk"],
                    "algo_bytes": int(st["algo_bytes"]), "rows": int(r.out_off[-1]),
                    "phases_ms": {k: round(v, 1) for k, v in zip(s.phase_names(), st["ms_phase"]) if v >= 1.0}})
    except Exception as e:  # the library fails loudly: record where and why
        out["stopped"] = str(e)[:300]
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] == "--one":
        n, _, m = sys.argv[2].partition(":")
        one(int(n), m == "nsl")
    else:
        for spec in sys.argv[1:] or ["845", "845:nsl", "1645", "1645:nsl", "3290:nsl", "3290", "10000:nsl", "30000:nsl", "100000:nsl"]:
            try:
                p = subprocess.run([sys.executable, __file__, "--one", spec], capture_output=True, text=True, timeout=600)
                line = [l for l in p.stdout.splitlines() if l.startswith("{")]
                print(line[-1] if line else json.dumps({"spec": spec, "stopped": "process died: " + p.stderr[-300:]}), flush=True)
            except subprocess.TimeoutExpired:
                print(json.dumps({"spec": spec, "stopped": "no result within 600 s"}), flush=True)
