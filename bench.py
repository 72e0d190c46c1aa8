#!/usr/bin/env python
"""bench.py — PAF alignment blocks/sec through the alignasm hot path (graph build + k-walks + selection).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c1|c3|c5]

One step = one pass of solve_ctg_read over every contig of one synthetic PAF batch (BASELINE.json configs[1]:
human-scale diploid assembly, ~500k alignment blocks, tools/synth_paf.cpp --preset c2).  One process per GPU;
N>1 is launched by torchrun, every rank solves its own replica of the workload (seed 2+rank: contigs are
independent, there is no data-path collective => weak scaling) and the job value is N*blocks / max-rank time.

Prints ONE JSON line (rank 0):
  value     blocks/s with the batch already resident in HBM (aa_solve_device), max over ranks
  e2e       blocks/s through the public host-buffer call aa_solve (H2D of the batch + D2H of the rows inside)
  roofline  the dominant kernel (phase) of the step: algorithmic bytes / CUDA-event time vs measured HBM peak
  cli       the stages of the drop-in command line: host reader (all host threads), solve, host writers
  cpu_baseline  the reference's own solve_ctg_read (oracle/_ref, built from the reference sources) on this
            box's host cores, on a bounded sample of the same workload
--impl reference times that CPU arm alone (no GPU code on its path).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    "c1": (["--preset", "c1"], 1, "synthetic small PAF: 1k contigs x ~50 blocks (configs[0])"),
    "c2": (["--preset", "c2"], 2, "synthetic human-scale diploid assembly PAF, ~500k blocks, 260 contigs (configs[1])"),
    "c3": (["--preset", "c3"], 3, "synthetic cancer-karyotype PAF, ~500k blocks (configs[2])"),
}
SAMPLE_MAX_BLOCKS_PER_CONTIG = 4000   # the reference allocates 56*n^2 B per contig (paf_data.cpp:268-282)
SAMPLE_TARGET_BLOCKS = 60000


def synth_bin():
    p = os.path.join(ROOT, "alignasm_b200", "synth_paf")
    if not os.path.exists(p):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.run([cxx, "-O2", "-std=c++17", "-o", p, os.path.join(ROOT, "tools", "synth_paf.cpp")], check=True)
    return p


def make_paf(workload, seed_shift, tmp):
    args, seed, _ = WORKLOADS[workload]
    path = os.path.join(tmp, f"{workload}_s{seed + seed_shift}.paf")
    subprocess.run([synth_bin(), *[str(a) for a in args], "--seed", str(seed + seed_shift), "-o", path], check=True,
                   capture_output=True)
    return path


def make_sample(paf_path, out_path):
    """Bounded sample of the workload for the CPU arm: whole contigs in file order, skipping contigs the reference
    cannot hold (56*n^2 B tables), until ~SAMPLE_TARGET_BLOCKS blocks."""
    groups, cur, name = [], [], None
    with open(paf_path) as f:
        for line in f:
            q = line.split("\t", 1)[0]
            if q != name:
                if cur:
                    groups.append(cur)
                cur, name = [], q
            cur.append(line)
    if cur:
        groups.append(cur)
    total_ctg, total_blk = len(groups), sum(len(g) for g in groups)
    picked, nblk = [], 0
    for g in groups:
        if len(g) > SAMPLE_MAX_BLOCKS_PER_CONTIG:
            continue
        picked.append(g)
        nblk += len(g)
        if nblk >= SAMPLE_TARGET_BLOCKS:
            break
    with open(out_path, "w") as f:
        for g in picked:
            f.writelines(g)
    desc = (f"{len(picked)} of {total_ctg} contigs ({nblk} of {total_blk} blocks), file order, contigs with >"
            f"{SAMPLE_MAX_BLOCKS_PER_CONTIG} blocks skipped (the reference allocates 56*n^2 B per contig)")
    return nblk, desc


def cpu_reference(paf_path, tmp, steps=1, warmup=0):
    """Time the reference's own CPU implementation of the path on a bounded sample, all host threads."""
    from oracle import oracle_py
    sample = os.path.join(tmp, "sample.paf")
    nblk, desc = make_sample(paf_path, sample)
    cores = os.cpu_count() or 1
    times = []
    if oracle_py.ref_binary("glibc"):
        kind = "reference"
        for i in range(warmup + steps):
            js = oracle_py.run_ref(sample, os.path.join(tmp, "sample_ref"), variant="glibc", threads=cores, no_write=True)
            if i >= warmup:
                times.append(js["solve_s"])
    else:  # oracle/_ref was not built on a box with the reference sources: time the CPU restatement instead
        kind = "port"
        import alignasm_b200 as aa
        pf = aa.read_paf(sample)
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            oracle_py.oracle_solve(pf.batch, threads=cores)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return {"value": nblk / sec, "unit": "blocks/s", "cores": cores, "kind": kind, "sample": desc,
            "seconds_per_pass": sec, "sample_blocks": nblk}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except (ValueError, KeyError):
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    K, W = a.steps, a.warmup
    wl_desc = WORKLOADS[a.workload][2]
    config = {"workload": f"{a.workload}: {wl_desc}", "generator": "tools/synth_paf.cpp", "walks_per_contig": 10000,
              "per_rank": "one replica of the workload per GPU (seed + rank), no collective on the data path",
              "l2": "inputs + workspace (>1 GB) are larger than the 126 MB L2; nothing is cached between steps"}

    with tempfile.TemporaryDirectory(prefix="aa_bench_") as tmp:
        # ------------------------------------------------------------------ reference arm (CPU only)
        if a.impl == "reference":
            if rank != 0:
                return 0
            paf = make_paf(a.workload, 0, tmp)
            cb = cpu_reference(paf, tmp, steps=max(1, K), warmup=min(W, 1))
            line = {"impl": "reference", "metric": "paf_alignment_blocks_per_sec", "value": cb["value"], "unit": "blocks/s",
                    "n_gpus": a.gpus, "steps": max(1, K), "warmup": min(W, 1), "ms_per_step": cb["seconds_per_pass"] * 1e3,
                    "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
                    "config": config, "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                    "e2e": {"value": cb["value"], "unit": "blocks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                    "gpu_launches": 0}
            print(json.dumps(line))
            return 0

        # ------------------------------------------------------------------ our arm (one process per GPU)
        import torch
        import alignasm_b200 as aa
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device — alignasm_b200 has no CPU path")
        torch.cuda.set_device(local)
        dist = None
        if world > 1:
            # NCCL prints its version banner on stdout at the VERSION level: keep stdout to the one JSON line
            if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
                os.environ["NCCL_DEBUG"] = "WARN"
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))

        def barrier():
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()

        paf = make_paf(a.workload, rank, tmp)
        pf = aa.read_paf(paf)
        batch = pf.batch
        solver = aa.Solver(local)
        names = solver.phase_names()
        dev = solver.upload(batch)

        # ---- value: batch resident in HBM ----
        for _ in range(W):
            solver.solve_device(dev, fetch=False)
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        barrier()
        t0 = time.perf_counter()
        ev_ms, ph_ms, launches = 0.0, [0.0] * 16, 0
        for _ in range(K):
            solver.solve_device(dev, fetch=False)
            st = solver.stats()
            ev_ms += st["ms_total"]
            ph_ms = [x + y for x, y in zip(ph_ms, st["ms_phase"])]
            launches += st["n_launch"]
        barrier()
        wall = time.perf_counter() - t0
        clocks = sampler.stop() if rank == 0 else None
        st = solver.stats()

        # ---- e2e: host buffers in, rows out, through the public call ----
        for _ in range(min(W, 2)):
            solver.solve(batch, copy=False).close()
        barrier()
        t1 = time.perf_counter()
        for _ in range(K):
            r = solver.solve(batch, copy=False)  # the rows as aa_solve hands them to a C caller (no second copy in Python)
            d2h = sum(v.nbytes for v in r.out.values()) + sum(v.nbytes for v in r.alt.values()) + r.out_off.nbytes + r.alt_off.nbytes
            r.close()
        barrier()
        wall_e2e = time.perf_counter() - t1
        h2d = sum(getattr(batch, n).nbytes for n, _ in aa.Batch.FIELDS) + 4 * batch.n_blk

        # ---- drop-in CLI stages (parse + write reported separately): host reader, solve, host writers ----
        cli = None
        if rank == 0:
            t2 = time.perf_counter()
            pf2 = aa.read_paf(paf)
            t3 = time.perf_counter()
            r = solver.solve(pf2.batch)
            t4 = time.perf_counter()
            pf2.write(r, os.path.join(tmp, "cli_out"))
            t5 = time.perf_counter()
            out_bytes = sum(os.path.getsize(os.path.join(tmp, "cli_out" + e)) for e in (".aln.paf", ".aln.alt.paf", ".aln.all.paf"))
            r.close()
            pf2.close()
            cli = {"read_s": t3 - t2, "solve_s": t4 - t3, "write_s": t5 - t4, "blocks_per_s": batch.n_blk / (t5 - t2),
                   "paf_bytes": os.path.getsize(paf), "out_bytes": out_bytes, "host_threads": os.cpu_count(),
                   "note": "aa_paf_read + aa_solve + aa_paf_write as `alignasm --no_all` chains them (the .aln.all.paf of this workload, every tied "
                           "max-coverage walk of every contig, is 25.7 GB and is left out); process and CUDA start-up excluded"}

        # max over ranks
        tv = torch.tensor([wall, wall_e2e, ev_ms / 1e3], dtype=torch.float64, device="cuda")
        nb = torch.tensor([batch.n_blk], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(tv, op=dist.ReduceOp.MAX)
            dist.all_reduce(nb, op=dist.ReduceOp.SUM)
        wall_max, e2e_max, ev_max = tv.tolist()
        blocks = nb.item()
        if rank != 0:
            if dist is not None:
                dist.destroy_process_group()
            return 0

        peak, peak_src = hbm_peak()
        ph = [m / K for m in ph_ms]
        dom = max(range(len(names)), key=lambda i: ph[i])
        algo = st["algo_bytes_phase"][dom]
        achieved = algo / (ph[dom] * 1e-3) / 1e9 if ph[dom] > 0 else 0.0
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath) and a.workload == "c2":  # the ncu capture is of this workload
            try:
                traffic = json.load(open(tpath)).get(names[dom])
            except ValueError:
                traffic = None
        line = {
            "metric": "paf_alignment_blocks_per_sec", "value": blocks * K / wall_max, "unit": "blocks/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": wall_max / K * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int64", "data": "synthetic", "config": config,
            "device_ms_per_step": ev_max / K * 1e3,
            "e2e": {"value": blocks * K / e2e_max, "unit": "blocks/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": algo, "ms_per_launch": ph[dom],
                         "note": "latency-bound integer graph work: see DESIGN.md for the per-phase byte model"},
            "phases_ms": {n: round(m, 3) for n, m in zip(names, ph)},
            "sizes": {k: st[k] for k in ("n_ctg", "n_blk", "n_pair", "n_vtx", "n_edge", "n_heap", "n_walk", "n_task")},
            "cli": cli,
        }
        if not a.no_cpu_baseline and world == 1:
            cb = cpu_reference(paf, tmp)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line))
        if dist is not None:
            dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
