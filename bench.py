#!/usr/bin/env python
"""bench.py — PAF alignment blocks/sec through the alignasm hot path (graph build + k-walks + selection).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload auto|c1|c2|c3|c4|c5] [--n BLOCKS] [--nsl] [--no-cpu-baseline]

One step = one pass of solve_ctg_read over every contig of one synthetic PAF (tools/synth_paf.cpp, seeded).

Workloads (BASELINE.json configs): c1 1k small contigs, c2 human-scale diploid assembly (~500k blocks, 260 contigs),
c3 cancer karyotype, c4 one dense contig of --n blocks (--nsl: --non_skip_linkable), c5 the box-scaling input:
8 concatenated cancer PAFs, ~4.4 M blocks, 2 080 contigs.

--workload auto (the default, what the driver runs): the job is ONE c5 input, contig-sharded over the N ranks
(alignasm_b200/sharding.py: cost-balanced LPT, no data-path collective; every rank keeps its shard resident, solves it, and
publishes its row lists in a shared-memory segment that rank 0 maps, inside the timed region) => strong scaling, N = 1 is the whole input on one
GPU.  At N = 1 the line also carries the single-GPU figures of c2 (BASELINE.json configs[1], the workload the metric is
quoted on) as the sub-object "c2": value, e2e, roofline, phases, sizes, cli, same-work comparison and cpu_baseline.

Prints ONE JSON line (rank 0):
  value        blocks/s of the whole job with the shards resident in HBM (aa_solve_device + gather), max over ranks
  e2e          blocks/s through the public host-buffer call aa_solve (H2D of the shard from page-locked host arrays, aa_host_alloc,
               + D2H of the rows inside) + gather
  roofline     the dominant kernel (phase) of the step: algorithmic bytes / CUDA-event time vs the measured HBM peak
  per_rank     device ms and shard sizes of every rank (the imbalance is the largest contig's serial chain)
  cpu_baseline the reference's own solve_ctg_read (oracle/_ref, compiled from the reference sources) on this box's host
               cores, on a bounded sample of the same workload; plus one host thread on a smaller sample, with peak RSS
  same_work    the GPU timed on exactly the contigs of the CPU sample (resident and through aa_solve)
--impl reference times the CPU arm alone (no GPU code on its path) on the sample the same --steps/--warmup give our arm.
want_all is false throughout: the path's outputs are the .aln.paf / .aln.alt.paf rows (the .aln.all.paf list of c2, every tied
max-coverage walk of every contig, is 25.7 GB); the reference builds that list too (paf_data.cpp:1595-1611).
"""
import argparse
import json
import math
import os
import resource
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
    "c3": (["--preset", "c3"], 3, "synthetic cancer-karyotype PAF, ~550k blocks, 260 contigs (configs[2])"),
    "c4": (["--preset", "c4"], 4, "pathological dense contig: one contig, every block overlaps the next 10-50 (configs[3])"),
    "c5": (["--preset", "c5", "--replicas", "8"], 30, "whole-box scaling input: 8 concatenated cancer PAFs, ~4.4 M blocks, 2 080 contigs (configs[4])"),
}


def synth_bin():
    p = os.path.join(ROOT, "alignasm_b200", "synth_paf")
    if not os.path.exists(p):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.run([cxx, "-O2", "-std=c++17", "-o", p, os.path.join(ROOT, "tools", "synth_paf.cpp")], check=True)
    return p


def make_paf(workload, tmp, n=None):
    args, seed, _ = WORKLOADS[workload]
    extra = ["--n", str(n)] if (workload == "c4" and n) else []
    path = os.path.join(tmp, f"{workload}{'_' + str(n) if extra else ''}_s{seed}.paf")
    if not os.path.exists(path):
        subprocess.run([synth_bin(), *[str(a) for a in args], *extra, "--seed", str(seed), "-o", path], check=True, capture_output=True)
    return path


# ------------------------------------------------------------------------------------------------ CPU arm
def mem_available():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) * 1024
    except OSError:
        pass
    return 32 << 30


def read_groups(paf_path, limit_rows=None):
    """Contigs of a PAF in file order: list of lists of lines."""
    groups, cur, name, rows = [], [], None, 0
    with open(paf_path) as f:
        for line in f:
            q = line.split("\t", 1)[0]
            if q != name:
                if cur:
                    groups.append(cur)
                    if limit_rows and rows >= limit_rows:
                        cur = []
                        break
                cur, name = [], q
            cur.append(line)
            rows += 1
    if cur:
        groups.append(cur)
    return groups


def pick_sample(paf_path, out_path, threads, passes):
    """Bounded sample of the workload for the CPU arm: whole contigs in file order.
      * a contig is admitted up to the size the reference's four n x n tables (56 n^2 B, paf_data.cpp:268-282) allow on this box
        with every host thread holding one: n <= sqrt(MemAvailable / 2 / (56 threads)), capped at 6 000 (a 6 000-block contig alone is
        2 GB of tables and several seconds of one core);
      * contigs are taken until every host thread has at least 4 of them AND ~2 M / passes blocks (50 k .. 160 k) are reached, so that
        `passes` passes stay within a few minutes."""
    groups = read_groups(paf_path, limit_rows=1200000)  # (c5: the sample comes from the first replicas)
    total_ctg, total_blk = len(groups), sum(len(g) for g in groups)
    n_mem = int(math.sqrt(mem_available() / 2 / (56.0 * max(1, threads))))
    target = int(min(160000, max(50000, 2.0e6 / max(1, passes))))
    n_max = max(500, min(n_mem, 6000 if target >= 120000 else 4000))
    picked, nblk = [], 0
    for g in groups:
        if len(g) > n_max or len(g) < 2:
            continue
        picked.append(g)
        nblk += len(g)
        if nblk >= target and len(picked) >= 4 * threads:
            break
    with open(out_path, "w") as f:
        for g in picked:
            f.writelines(g)
    desc = (f"{len(picked)} whole contigs, {nblk} blocks, in file order out of the first {total_ctg} contigs / {total_blk} blocks; contigs above "
            f"{n_max} blocks skipped (the reference allocates 56 n^2 B per contig: MemAvailable allows {n_mem} with {threads} threads; cap "
            f"{6000 if target >= 120000 else 4000}); sized for {passes} pass(es): target {target} blocks and >= 4 contigs per host thread")
    return picked, nblk, desc


def run_cpu(sample_path, tmp, threads, steps=1, warmup=0):
    """The reference's own CPU implementation of the path (oracle/_ref/alignasm_ref: paf_data.cpp + headers compiled unmodified, a
    std::thread pool over contigs standing in for tbb::parallel_for, alignasm.cpp:351-359); falls back to the CPU restatement when
    oracle/_ref was not built on a box with the reference sources.  Returns seconds per pass (mean of `steps`), kind, peak RSS."""
    from oracle import oracle_py
    times = []
    if oracle_py.ref_binary("glibc"):
        kind = "reference"
        for i in range(warmup + steps):
            js = oracle_py.run_ref(sample_path, os.path.join(tmp, "sample_ref"), variant="glibc", threads=threads, no_write=True)
            if i >= warmup:
                times.append(js["solve_s"])
        rss = resource.getrusage(resource.RUSAGE_CHILDREN).ru_maxrss / 1024.0  # MB, the largest child so far
    else:
        kind = "port"
        import alignasm_b200 as aa
        pf = aa.read_paf(sample_path)
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            oracle_py.oracle_solve(pf.batch, threads=threads)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
        rss = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1024.0
    return sum(times) / len(times), kind, rss


def cpu_arm(paf_path, tmp, passes, steps=1, warmup=0, one_thread=True):
    """cpu_baseline object: all host threads on the sample, and one thread on its first ~12 k blocks."""
    cores = os.cpu_count() or 1
    sample = os.path.join(tmp, "sample.paf")
    picked, nblk, desc = pick_sample(paf_path, sample, cores, passes)
    out = {"unit": "blocks/s", "cores": cores, "sample": desc, "sample_blocks": nblk, "sample_path": sample}
    if one_thread:
        small = os.path.join(tmp, "sample_1t.paf")
        nb1 = 0
        with open(small, "w") as f:
            for g in picked:
                f.writelines(g)
                nb1 += len(g)
                if nb1 >= 12000:
                    break
        sec1, kind1, rss1 = run_cpu(small, tmp, 1)
        out["one_thread"] = {"value": nb1 / sec1, "unit": "blocks/s", "cores": 1, "sample": f"the first {nb1} blocks of the sample", "peak_rss_mb": round(rss1, 1)}
    sec, kind, rss = run_cpu(sample, tmp, cores, steps=steps, warmup=warmup)
    out.update({"value": nblk / sec, "kind": kind, "seconds_per_pass": sec, "peak_rss_mb": round(rss, 1)})
    return out


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


def roofline_of(names, ph, st, traffic_ok):
    """The dominant phase of a step against the HBM roofline (algorithmic bytes per phase from the run's own sizes)."""
    peak, peak_src = hbm_peak()
    dom = max(range(len(names)), key=lambda i: ph[i])
    algo = st["algo_bytes_phase"][dom]
    achieved = algo / (ph[dom] * 1e-3) / 1e9 if ph[dom] > 0 else 0.0
    traffic, tsrc = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if traffic_ok and os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(names[dom])
            tsrc = "static: dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this workload (profiles/traffic.json), not measured in this run"
        except ValueError:
            traffic = None
    return {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "traffic_source": tsrc, "peak_source": peak_src, "algorithmic_bytes_per_launch": algo,
            "ms_per_launch": ph[dom], "note": "latency-bound integer graph work: see DESIGN.md for the per-phase byte model"}


def main():
    # libraries (NCCL's version banner, torchrun notices) write to stdout: the one JSON line must be alone there
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto"] + sorted(WORKLOADS))
    ap.add_argument("--n", type=int, default=845, help="c4: blocks of the dense contig")
    ap.add_argument("--nsl", action="store_true", help="--non_skip_linkable")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    K, W = max(1, a.steps), max(0, a.warmup)
    auto = a.workload == "auto"
    wl = "c5" if auto else a.workload
    wl_desc = WORKLOADS[wl][2] + (f", n = {a.n}" if wl == "c4" else "") + (", --non_skip_linkable" if a.nsl else "")
    config = {"workload": f"{wl}: {wl_desc}", "generator": "tools/synth_paf.cpp", "walks_per_contig": 10000, "want_all": False,
              "non_skip_linkable": bool(a.nsl),
              "sharding": ("one input, contigs LPT-sharded over the ranks by estimated cost (alignasm_b200/sharding.py), shards resident in HBM, "
                           "no collective on the data path: every rank publishes the primary/alt rows of its shard in a POSIX shared-memory segment that "
                           "rank 0 maps (host-side merge, inside the timed region); NCCL carries only the barriers and the timing reductions" if wl != "c4" else "a single contig does not shard: replicas only (every rank solves the same contig)"),
              "cpu_sample_rule": "see cpu_baseline.sample",
              "l2": "inputs + workspace (> 1 GB) are larger than the 126 MB L2; nothing is cached between steps",
              "e2e_host_memory": "inputs in page-locked host arrays (aa_host_alloc), rows returned in the library's pinned result slab; "
                                 "pageable inputs go through one more staging pass (see DESIGN.md section 6)"}

    with tempfile.TemporaryDirectory(prefix="aa_bench_") as tmp:
        # ------------------------------------------------------------------ reference arm (CPU only, rank 0 only)
        if a.impl == "reference":
            if rank != 0:
                return 0
            paf = make_paf(wl, tmp, a.n)
            cb = cpu_arm(paf, tmp, passes=K + W, steps=K, warmup=W, one_thread=False)
            line = {"impl": "reference", "metric": "paf_alignment_blocks_per_sec", "value": cb["value"], "unit": "blocks/s",
                    "n_gpus": a.gpus, "steps": K, "warmup": W, "ms_per_step": cb["seconds_per_pass"] * 1e3,
                    "higher_is_better": True, "scaling": "strong" if wl != "c4" else "weak", "vs_baseline": None, "dtype": "int64",
                    "data": "synthetic", "config": config,
                    "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "peak_rss_mb")},
                    "e2e": {"value": cb["value"], "unit": "blocks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                    "gpu_launches": 0}
            emit(line)
            return 0

        # ------------------------------------------------------------------ our arm (one process per GPU)
        import numpy as np
        import torch
        import alignasm_b200 as aa
        from alignasm_b200 import sharding
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device — alignasm_b200 has no CPU path")
        torch.cuda.set_device(local)
        cuda = torch.device("cuda", local)
        dist = None
        if world > 1:
            # NCCL prints its version banner on stdout at the VERSION level: keep stdout to the one JSON line
            if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
                os.environ["NCCL_DEBUG"] = "WARN"
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=cuda)

        def barrier():
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()

        solver = aa.Solver(local)
        names = solver.phase_names()
        opts = {"non_skip_linkable": bool(a.nsl)}

        def timed(batch_full, shard_ids, sampler=None):
            """value / e2e of one job: `batch_full` sharded as `shard_ids` (None: this rank solves all of it and nothing is gathered)."""
            sub = batch_full.select(shard_ids[rank]) if shard_ids is not None else batch_full
            have = sub.n_ctg > 0
            dev = solver.upload(sub) if have else None
            # the e2e leg reads its inputs from page-locked host memory (the bench contract; aa_host_alloc in the C ABI)
            sub_host = sub.pinned() if have else sub
            gather = shard_ids is not None and world > 1
            shm = None
            if gather:  # one solve tells how many rows the shard has; the shared segment gets 25 % head room
                need = 64
                if have:
                    r = solver.solve_device(dev, copy=False, **opts)
                    need = sharding.packed_size(r)
                    r.close()
                shm = sharding.ShmRows(f"aa_bench_{os.environ.get('MASTER_PORT', '0')}", rank, world, need + need // 4 + 4096)
                barrier()
                if rank == 0:
                    shm.attach_all()

            def step(resident):
                nbytes = 0
                if have:
                    r = solver.solve_device(dev, copy=False, **opts) if resident else solver.solve(sub_host, copy=False, **opts)
                    nbytes = sum(v.nbytes for v in r.out.values()) + sum(v.nbytes for v in r.alt.values()) + r.out_off.nbytes + r.alt_off.nbytes
                    if gather:
                        shm.publish(r)
                    r.close()
                elif gather:
                    shm.publish_empty()
                if gather:
                    barrier()  # every rank's rows are in its segment
                    if rank == 0:  # the writer's view: input contig -> (shard, position); the rows stay where their rank put them
                        sharding.contig_index(batch_full.n_ctg, shard_ids)
                        assert sum(int(v[0].shape[0]) - 1 for v in shm.views()) == batch_full.n_ctg
                    barrier()  # (the segments are free to be overwritten by the next step)
                return nbytes

            for _ in range(W):
                step(True)
            if sampler is not None:
                sampler.start()
            barrier()
            t0 = time.perf_counter()
            ev_ms, ph_ms, launches = 0.0, [0.0] * 16, 0
            for _ in range(K):
                step(True)
                if have:
                    st = solver.stats()
                    ev_ms += st["ms_total"]
                    ph_ms = [x + y for x, y in zip(ph_ms, st["ms_phase"])]
                    launches += st["n_launch"]
            barrier()
            wall = time.perf_counter() - t0
            clocks = sampler.stop() if sampler is not None else None
            st = solver.stats() if have else None
            for _ in range(min(W, 2)):
                step(False)
            barrier()
            t1 = time.perf_counter()
            d2h = 0
            for _ in range(K):
                d2h = step(False)
            barrier()
            wall_e2e = time.perf_counter() - t1
            h2d = sum(getattr(sub, n).nbytes for n, _ in aa.Batch.FIELDS) + 4 * sub.n_blk if have else 0
            if dev is not None:
                dev.free()
            if shm is not None:
                barrier()
                shm.close()
            return {"wall": wall, "wall_e2e": wall_e2e, "ev_ms": ev_ms / K, "ph": [m / K for m in ph_ms], "launches": launches,
                    "st": st, "clocks": clocks, "h2d": int(h2d), "d2h": int(d2h), "blocks": sub.n_blk, "contigs": sub.n_ctg,
                    "largest": int(np.diff(sub.ctg_off).max()) if have else 0}

        def allmax(vals):
            t = torch.tensor(vals, dtype=torch.float64, device=cuda)
            if dist is not None:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return t.tolist()

        def allsum(vals):
            t = torch.tensor(vals, dtype=torch.float64, device=cuda)
            if dist is not None:
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
            return t.tolist()

        # ---- the job ----
        paf = make_paf(wl, tmp, a.n)
        pf = aa.read_paf(paf)
        batch = pf.batch
        if wl == "c4":
            shard_ids, job_blocks = None, batch.n_blk * world  # replicas only
        else:
            shard_ids = sharding.lpt_shards(sharding.contig_costs(batch), world)
            job_blocks = batch.n_blk
        sampler = ClockSampler(local) if rank == 0 else None
        m = timed(batch, shard_ids, sampler)
        wall_max, e2e_max, ev_max = allmax([m["wall"], m["wall_e2e"], m["ev_ms"]])
        launches, h2d, d2h = allsum([m["launches"], m["h2d"], m["d2h"]])
        per_rank = None
        if world > 1:
            t = torch.zeros(world, 4, dtype=torch.float64, device=cuda)
            t[rank] = torch.tensor([m["ev_ms"], m["blocks"], m["contigs"], m["largest"]], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            per_rank = [{"rank": r, "device_ms": round(v[0], 3), "blocks": int(v[1]), "contigs": int(v[2]), "largest_contig": int(v[3])}
                        for r, v in enumerate(t.tolist())]
        if rank != 0:
            if dist is not None:
                dist.destroy_process_group()
            return 0
        st = m["st"]
        line = {
            "metric": "paf_alignment_blocks_per_sec", "value": job_blocks * K / wall_max, "unit": "blocks/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": wall_max / K * 1e3, "higher_is_better": True,
            "scaling": "strong" if wl != "c4" else "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic", "config": config,
            "device_ms_per_step": ev_max,
            "e2e": {"value": job_blocks * K / e2e_max, "unit": "blocks/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches),
            "clocks": m["clocks"],
            "roofline": roofline_of(names, m["ph"], st, traffic_ok=False),
            "phases_ms": {n: round(x, 3) for n, x in zip(names, m["ph"])},
            "phases_note": "CUDA-event pairs on the main stream; on batches of many contigs the small contigs' heaps + enumeration run on a second "
                           "stream inside the `heaps` phase and `enum` is the large contigs' (DESIGN.md 3.5)",
            "sizes": {k: st[k] for k in ("n_ctg", "n_blk", "n_pair", "n_vtx", "n_edge", "n_heap", "n_walk", "n_task")},
            "per_rank": per_rank,
        }
        if world == 1:
            line["sizes_note"] = "rank 0 = the whole input"
        else:
            line["sizes_note"] = "sizes and phases_ms are rank 0's shard; per_rank has every rank's device time"

        # ---- N = 1: the CPU arm on a bounded sample, the GPU on the same contigs, and the c2 figures ----
        def same_work(cb):
            sb = aa.read_paf(cb["sample_path"]).batch
            sm = timed(sb, None)
            return {"sample": cb["sample"], "blocks": sb.n_blk, "contigs": sb.n_ctg,
                    "gpu_resident_blocks_per_s": sb.n_blk * K / sm["wall"], "gpu_e2e_blocks_per_s": sb.n_blk * K / sm["wall_e2e"],
                    "gpu_ms_per_pass": sm["wall"] / K * 1e3, "cpu_blocks_per_s": cb["value"], "cpu_cores": cb["cores"], "cpu_kind": cb["kind"],
                    "ratio_e2e": (sb.n_blk * K / sm["wall_e2e"]) / cb["value"]}

        def cpu_fields(cb):
            out = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "peak_rss_mb")}
            if "one_thread" in cb:
                out["one_thread"] = cb["one_thread"]
            return out

        if world == 1 and not a.no_cpu_baseline and wl != "c4":
            cb = cpu_arm(paf, tmp, passes=K + W)
            line["cpu_baseline"] = cpu_fields(cb)
            line["same_work"] = same_work(cb)
        if world == 1 and auto:
            paf2 = make_paf("c2", tmp)
            pf2 = aa.read_paf(paf2)
            m2 = timed(pf2.batch, None)
            st2 = m2["st"]
            c2 = {"workload": "c2: " + WORKLOADS["c2"][2], "value": pf2.batch.n_blk * K / m2["wall"], "unit": "blocks/s",
                  "ms_per_step": m2["wall"] / K * 1e3, "device_ms_per_step": m2["ev_ms"],
                  "e2e": {"value": pf2.batch.n_blk * K / m2["wall_e2e"], "unit": "blocks/s", "h2d_bytes_per_step": m2["h2d"], "d2h_bytes_per_step": m2["d2h"]},
                  "roofline": roofline_of(names, m2["ph"], st2, traffic_ok=True),
                  "phases_ms": {n: round(x, 3) for n, x in zip(names, m2["ph"])},
                  "sizes": {k: st2[k] for k in ("n_ctg", "n_blk", "n_pair", "n_vtx", "n_edge", "n_heap", "n_walk", "n_task")}}
            # the stages of the drop-in command line (parse + write reported separately)
            solver.solve(pf2.batch, copy=False).close()  # (warm-up: the first pageable aa_solve of a process allocates the staging buffer)
            t2 = time.perf_counter()
            pf3 = aa.read_paf(paf2)
            t3 = time.perf_counter()
            r = solver.solve(pf3.batch, copy=False)
            t4 = time.perf_counter()
            pf3.write(r, os.path.join(tmp, "cli_out"))
            t5 = time.perf_counter()
            out_bytes = sum(os.path.getsize(os.path.join(tmp, "cli_out" + e)) for e in (".aln.paf", ".aln.alt.paf", ".aln.all.paf"))
            r.close()
            pf3.close()
            c2["cli"] = {"read_s": t3 - t2, "solve_s": t4 - t3, "write_s": t5 - t4, "blocks_per_s": pf2.batch.n_blk / (t5 - t2),
                         "paf_bytes": os.path.getsize(paf2), "out_bytes": out_bytes, "host_threads": os.cpu_count(),
                         "note": "aa_paf_read + aa_solve + aa_paf_write as `alignasm --no_all` chains them; process and CUDA start-up and first-call buffer allocation excluded"}
            # the same with the cs:Z: codec on the device (aa_paf_read_device / aa_paf_write_device)
            aa.read_paf(paf2, solver=solver).close()  # (warm-up: first-touch of the device buffers)
            t2 = time.perf_counter()
            pf4 = aa.read_paf(paf2, solver=solver)
            t3 = time.perf_counter()
            r = solver.solve(pf4.batch, copy=False)
            t4 = time.perf_counter()
            pf4.write(r, os.path.join(tmp, "cli_out_dev"), solver=solver)
            t5 = time.perf_counter()
            same = all(open(os.path.join(tmp, "cli_out" + e), "rb").read() == open(os.path.join(tmp, "cli_out_dev" + e), "rb").read()
                       for e in (".aln.paf", ".aln.alt.paf", ".aln.all.paf"))
            r.close()
            pf4.close()
            c2["cli_cs_device"] = {"read_s": t3 - t2, "solve_s": t4 - t3, "write_s": t5 - t4, "blocks_per_s": pf2.batch.n_blk / (t5 - t2),
                                   "files_identical_to_host_codec": bool(same),
                                   "note": "`alignasm --cs_device --no_all`: parse_short_cs / get_overlap_range / get_edited_paf_data as CUDA kernels over the file image"}
            if not a.no_cpu_baseline:
                cb2 = cpu_arm(paf2, tmp, passes=K + W)
                c2["cpu_baseline"] = cpu_fields(cb2)
                c2["same_work"] = same_work(cb2)
            line["c2"] = c2
        emit(line)
        if dist is not None:
            dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
