"""Condense an ncu launch list (--metrics gpu__time_duration.sum --csv) into one line per launch + per-kernel shares.
python tools/launch_table.py gpurun_out/launches.csv > profiles/rNN_launches.txt"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = next(r for r in rows if r and r[0] == "ID")
ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
tot = collections.OrderedDict()
lines = []
for r in rows[rows.index(hdr) + 1:]:
    if len(r) != len(hdr):
        continue
    name = re.sub(r"^void ", "", r[ki])
    name = re.sub(r"aa::k_(warp_)?items<aa::(\w+)>.*", r"\2", name)
    name = re.sub(r"cub::(\w+)<.*", r"cub::\1", name)[:60]
    ns = float(r[vi].replace(",", ""))
    lines.append(f"{int(r[0]):4d} {name:40s} grid {r[gi]:>14s} block {r[bi]:>12s} {ns/1e3:12.1f} us")
    tot[name] = tot.get(name, 0.0) + ns
print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare shares, not absolutes)")
print("\n".join(lines))
s = sum(tot.values())
print(f"\n# per-kernel totals over the listed launches ({s/1e6:.2f} ms)")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{k:40s} {v/1e6:10.3f} ms {100*v/s:6.2f} %")
