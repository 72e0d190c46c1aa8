"""CPU: the N>1 host logic — cost-balanced contig sharding, per-rank solve, gather, merge — with world_size 2 over
gloo.  The per-rank solve is the emulated core (no GPU here); on a GPU box bench.py / the CLI use the CUDA solver."""
import os
import sys

import numpy as np
import pytest

import parity_util as pu


def test_lpt_balances_and_covers():
    from alignasm_b200 import sharding
    rng = np.random.default_rng(3)
    costs = rng.lognormal(0, 1.3, size=260)
    for n in (1, 2, 4, 8):
        shards = sharding.lpt_shards(costs, n)
        ids = np.sort(np.concatenate(shards))
        assert np.array_equal(ids, np.arange(260))
        load = np.array([costs[s].sum() for s in shards])
        assert load.max() <= costs.sum() / n + costs.max() + 1e-9


def test_select_roundtrip(product_lib):
    import alignasm_b200 as aa
    from alignasm_b200 import sharding
    import emul_py
    pf = aa.read_paf(os.path.join(pu.GOLDEN, "tiny.paf"))
    b = pf.batch
    full = sharding.rows_by_contig(emul_py.emul_solve(b, want_all=True))
    shards = sharding.lpt_shards(sharding.contig_costs(b), 3)
    rows = [sharding.rows_by_contig(emul_py.emul_solve(b.select(s), want_all=True)) for s in shards]
    assert sharding.merge_shards(b.n_ctg, shards, rows) == full


def _worker(rank, world, port, paf, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    for p in (pu.ROOT, os.path.join(pu.ROOT, "tests"), os.path.join(pu.ROOT, "tests", "emul")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import alignasm_b200 as aa
    from alignasm_b200 import sharding
    import emul_py
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b = aa.read_paf(paf).batch

    def gather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    merged = sharding.solve_sharded(b, emul_py.emul_solve, rank, world, gather, want_all=True)
    if rank == 0:
        q.put(merged)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gloo(product_lib):
    import torch.multiprocessing as mp
    import alignasm_b200 as aa
    from alignasm_b200 import sharding
    import emul_py
    paf = os.path.join(pu.GOLDEN, "tiny.paf")
    emul_py.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, paf, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    single = sharding.rows_by_contig(emul_py.emul_solve(aa.read_paf(paf).batch, want_all=True))
    assert merged == single


def _gather_worker(rank, world, port, paf, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    for p in (pu.ROOT, os.path.join(pu.ROOT, "tests"), os.path.join(pu.ROOT, "tests", "emul")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import alignasm_b200 as aa
    from alignasm_b200 import sharding
    import emul_py
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b = aa.read_paf(paf).batch
    shards = sharding.lpt_shards(sharding.contig_costs(b), world)
    res = emul_py.emul_solve(b.select(shards[rank]))
    got = sharding.gather_packed(sharding.pack_rows(res), dist, rank, world)  # what bench.py --gpus N does with NCCL
    if rank == 0:
        shard_of, local = sharding.contig_index(b.n_ctg, shards)
        rows = []
        parts = [sharding.unpack_rows(g) for g in got]
        for c in range(b.n_ctg):  # a writer's walk over the contigs in input order
            out_off, alt_off, out, alt = parts[shard_of[c]]
            l = int(local[c])
            rows.append((tuple(int(x) for x in out["qry_str"][out_off[l]:out_off[l + 1]]), tuple(int(x) for x in alt["ref_end"][alt_off[l]:alt_off[l + 1]]),
                         tuple(int(x) for x in out["is_alt"][out_off[l]:out_off[l + 1]])))
        q.put(rows)
    else:
        assert got is None
    dist.barrier()
    dist.destroy_process_group()


def test_packed_gather_two_ranks_gloo(product_lib):
    """pack_rows / gather_packed / contig_index: the one-process-per-GPU merge of bench.py (NCCL there, gloo here)."""
    import torch.multiprocessing as mp
    import alignasm_b200 as aa
    import emul_py
    paf = os.path.join(pu.GOLDEN, "ties.paf")
    emul_py.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31000 + os.getpid() % 2000
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, paf, q)) for r in range(2)]
    for p in procs:
        p.start()
    rows = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    one = emul_py.emul_solve(aa.read_paf(paf).batch)
    for c in range(one.n_ctg):
        a, b = int(one.out_off[c]), int(one.out_off[c + 1])
        x, y = int(one.alt_off[c]), int(one.alt_off[c + 1])
        assert rows[c] == (tuple(int(v) for v in one.out["qry_str"][a:b]), tuple(int(v) for v in one.alt["ref_end"][x:y]),
                           tuple(int(v) for v in one.out["is_alt"][a:b]))
