"""Synthetic alternative PAF (the reference CLI's `--alt` input, alignasm.cpp:186-332) for a given main PAF.

Rows are alignments of contig segments named `<contig>:<START>-<END>` (1-based START, coordinates relative to the
segment).  The generator covers what the reference's reader distinguishes: groups with one or several rows above the
baseline, groups with none (the best row is taken), a group that is interrupted and resumed (two groups), segments that
overlap existing blocks (pair vertices between main and alt blocks), both strands, new target names, and a segment of a
contig the main PAF does not hold (the reference appends it to bucket 0)."""
import random

TARGETS = [("chr1", 248956422), ("chr7", 159345973), ("chrAltOnly", 5000000)]


def _cs(rng, n, fwd):
    """Ops in query orientation consuming n query and n reference bases; returns (cs text, matches)."""
    ops, left = [], n
    while left > 0:
        m = min(left, rng.randint(20, 400))
        ops.append((":", m))
        left -= m
        if left > 1 and rng.random() < 0.7:
            ops.append(("*", 1))
            left -= 1
    seq = ops if fwd else ops[::-1]
    return "cs:Z:" + "".join(f":{k}" if t == ":" else "*ag" for t, k in seq), sum(k for t, k in ops if t == ":")


def make_alt(main_paf, out_path, seed=1, p_contig=0.7):
    rng = random.Random(seed)
    contigs, order = {}, []
    with open(main_paf) as f:
        for line in f:
            c = line.split("\t")
            if c[0] not in contigs:
                contigs[c[0]] = {"len": int(c[1]), "iv": []}
                order.append(c[0])
            contigs[c[0]]["iv"].append((int(c[2]), int(c[3])))
    rows = []

    def emit(name, s, e, a, b):
        """One alignment of segment [s, e) of contig `name` covering [a, b) of the segment."""
        fwd = rng.random() < 0.6
        tname, tlen = rng.choice(TARGETS)
        ts = rng.randint(0, tlen - (b - a) - 1)
        cs, nm = _cs(rng, b - a, fwd)
        rows.append("\t".join(str(x) for x in (f"{name}:{s + 1}-{e}", e - s, a, b, "+" if fwd else "-", tname, tlen, ts,
                                                ts + (b - a), nm, b - a, rng.choice([0, 1, 30, 60, 60]), "tp:A:P", cs)) + "\n")

    for name in order:
        if rng.random() > p_contig:
            continue
        info = contigs[name]
        iv = sorted(info["iv"])
        gaps, reach = [], 0
        for qs, qe in iv:
            if qs - reach >= 400:
                gaps.append((reach, qs))
            reach = max(reach, qe)
        if info["len"] - reach >= 400:
            gaps.append((reach, info["len"]))
        rng.shuffle(gaps)
        for lo, hi in gaps[:3]:
            kind = rng.choice("ABCDE")
            if kind == "E" and lo >= 300:
                lo -= rng.randint(100, 300)  # reaches into the blocks on its left
            s, e = lo, min(hi, lo + rng.randint(400, 6000))
            n = e - s
            if kind == "A":  # one row above the baseline, one below
                emit(name, s, e, 0, int(n * 0.8))
                emit(name, s, e, int(n * 0.7), int(n * 0.95))
            elif kind == "B":  # nothing above the baseline: the best row of the group is taken
                emit(name, s, e, 0, int(n * 0.2))
                emit(name, s, e, int(n * 0.3), int(n * 0.7))
                emit(name, s, e, int(n * 0.75), int(n * 0.9))
            elif kind == "C":  # two rows above the baseline, overlapping each other
                emit(name, s, e, 0, int(n * 0.6))
                emit(name, s, e, int(n * 0.35), n)
            elif kind == "D":  # a single short row
                emit(name, s, e, int(n * 0.1), int(n * 0.45))
            else:
                emit(name, s, e, 0, int(n * 0.9))
    if order:  # a group interrupted by another segment and resumed = two groups of that segment
        name = order[0]
        L = contigs[name]["len"]
        if L > 2000:
            emit(name, L - 900, L, 0, 300)
            emit(name, L - 1800, L - 1000, 0, 200)
            emit(name, L - 900, L, 400, 650)
    emit("ghost_contig", 0, 700, 0, 600)  # unknown contig -> bucket 0
    with open(out_path, "w") as f:
        f.writelines(rows)
    return out_path
