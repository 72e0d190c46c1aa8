"""Summarise an .ncu-rep (read here, no GPU needed): per kernel duration, instructions, DRAM bytes, hit rates, stalls."""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "launch__shared_mem_per_block_dynamic", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
ki = hdr.index("Kernel Name")
for r in rows[2:]:
    print("==", r[ki][:90])
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            print(f"   {k:70s} {r[i]} {units[i]}")
# stall breakdown per kernel from the source page
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
cur = None; hd = None; agg = collections.defaultdict(lambda: collections.Counter()); inst = collections.Counter()
for r in csv.reader(io.StringIO(src)):
    if not r: continue
    if r[0] == "Kernel Name": cur = r[1][:60]; hd = None; continue
    if r[0] == "Address": hd = r; continue
    if hd and cur and len(r) == len(hd):
        for i, h in enumerate(hd):
            if h.startswith("stall_") and "Not Issued" not in h:
                try: agg[cur][h] += int(r[i])
                except ValueError: pass
        try: inst[cur] += int(r[hd.index("Instructions Executed")])
        except ValueError: pass
for k, c in agg.items():
    tot = sum(c.values())
    print("--", k, "warp-instr", inst[k], "samples", tot)
    print("   ", ", ".join(f"{n[6:]} {100*v/tot:.0f}%" for n, v in c.most_common(7)))
