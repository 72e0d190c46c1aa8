#!/usr/bin/env python
"""Pins the BENCH workloads themselves (C2 seed 2, C3 seed 3 at full size) against the CPU restatement (oracle/oracle_port.cpp).

Run in the CPU container (minutes: the 43 099-block contig of C2 takes ~4 min of one core):
    python tests/golden/make_fullsize.py [c2] [c3] [dense845] [dense845.nsl] [dense1645] [dense1645.nsl] [dense3290.nsl]
Writes tests/golden/fullsize_<tag>.json: sha256 of every result array of the port (primary rows, alt rows, sorted index, per-contig
offsets), sha256 of the walk-distance lists, per-contig (blocks, vertices, edges, walks) and the totals (n_pair, n_vtx, n_edge,
n_heap, n_walk, n_task).  tests/test_gpu_parity.py::test_fullsize_pins solves the same synthetic input on the GPU and compares.
The input is regenerated from the seed by tools/synth_paf.cpp on both sides, so only this small JSON is committed."""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_util as pu  # noqa: E402

from fullsize_util import PINS, STAT_KEYS, digest  # noqa: E402


def main():
    import alignasm_b200 as aa
    from oracle import oracle_py
    for tag in sys.argv[1:] or list(PINS):
        args, opts = PINS[tag]
        paf = pu.synth(os.path.join("/tmp", f"fullsize_{tag}.paf"), *args)
        pf = aa.read_paf(paf)
        t0 = time.time()
        res = oracle_py.oracle_solve(pf.batch, threads=os.cpu_count() or 1, want_all=False, keep_debug=True, **opts)
        dt = time.time() - t0
        d, per = digest(res)
        big = int(np.argmax(per[0]))
        out = {"workload": tag, "generator": "tools/synth_paf.cpp " + " ".join(str(x) for x in args), "options": opts, "oracle": "oracle/oracle_port.cpp",
               "oracle_seconds": round(dt, 1), "stats": {k: int(res.stats[k]) for k in STAT_KEYS},
               "largest_contig": {"index": big, "V": int(per[0][big]), "E": int(per[1][big]), "K": int(per[2][big])}, "sha256": d}
        with open(os.path.join(HERE, f"fullsize_{tag}.json"), "w") as f:
            json.dump(out, f, indent=1)
            f.write("\n")
        print(tag, "done in %.0f s" % dt, out["stats"], flush=True)


if __name__ == "__main__":
    main()
