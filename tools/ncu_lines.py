"""Per-CUDA-source-line hot spots of one kernel of an .ncu-rep: python tools/ncu_lines.py rep kernel-regex [top]"""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name-base", "demangled",
                      "-k", "regex:" + pat], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, cur_file, lines = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0]:
        lines.append((cur_file, r))
si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
def num(x):
    try:
        return int(x)
    except ValueError:
        return 0
tot = sum(num(r[si]) for _, r in lines) or 1
toti = sum(num(r[ii]) for _, r in lines) or 1
print(f"kernel {pat}: samples {tot}, warp instructions {toti}")
lines.sort(key=lambda fr: -num(fr[1][si]))
for f, r in lines[:top]:
    st = sorted(((num(r[i]), hdr[i][6:]) for i in stalls), reverse=True)[:2]
    print(f"{f[:14]:14s}:{r[0]:>5s} smp {100*num(r[si])/tot:5.1f}% ins {100*num(r[ii])/toti:5.1f}%  {st[0][1]}/{st[1][1]:12s} {r[1].strip()[:100]}")
