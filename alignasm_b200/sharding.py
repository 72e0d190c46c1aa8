"""Contig sharding across GPUs / ranks.

Contigs are independent (reference: tbb::parallel_for over contigs, src/alignasm.cpp:351-359), so the multi-GPU
path partitions them: a host-side cost estimate per contig, LPT (largest first) into one bin per rank, every rank
solves its own sub-batch on its own GPU, and the per-contig row lists are merged back in input order.
There is no collective on the data path; torch.distributed is only used to gather the (small) results.
"""
import numpy as np


def contig_costs(batch, walks=10000):
    """Estimated SM time per contig (one warp per contig in the serial phases), same model as aa_shard_contigs in
    csrc/aa_multi.cpp: ~1.3 us per enumerated walk whatever the contig's size, ~0.6 us per block for the heap / relax /
    walk-0 chains.  Largest first also spreads the longest serial chains, which bound a shard from below."""
    n = np.diff(batch.ctg_off).astype(np.float64)
    enum_cost = np.where(n > 1, 1.3 * min(walks, 10000), 0.0)
    return enum_cost + 0.6 * n


def lpt_shards(costs, n_shards):
    """Longest-processing-time-first assignment; returns a list of sorted contig-index arrays, one per shard."""
    costs = np.asarray(costs, dtype=np.float64)
    order = np.argsort(-costs, kind="stable")
    load = np.zeros(n_shards)
    bins = [[] for _ in range(n_shards)]
    for c in order:
        k = int(np.argmin(load))
        bins[k].append(int(c))
        load[k] += costs[c]
    return [np.array(sorted(b), dtype=np.int64) for b in bins]


def rows_by_contig(result):
    """Per-contig (out, alt, all) row lists of a Result, as python tuples (small; used for the merge)."""
    return [(result.rows_of("out", c), result.rows_of("alt", c), result.all_of(c)) for c in range(result.n_ctg)]


def merge_shards(n_ctg, shards, shard_rows):
    """shards[k] = contig ids of shard k, shard_rows[k] = rows_by_contig of its result -> list over all contigs."""
    merged = [None] * n_ctg
    for ids, rows in zip(shards, shard_rows):
        assert len(ids) == len(rows)
        for c, r in zip(ids.tolist(), rows):
            merged[c] = r
    assert all(m is not None for m in merged), "a contig was not assigned to any shard"
    return merged


def solve_sharded(batch, solve_fn, rank=0, world=1, gather=None, **opts):
    """Solve `batch` split over `world` ranks.  solve_fn(sub_batch, **opts) -> Result runs on this rank's device;
    gather(obj) -> list of every rank's obj (e.g. torch.distributed.all_gather_object); returns the merged
    per-contig rows on every rank."""
    shards = lpt_shards(contig_costs(batch), world)
    mine = shards[rank]
    rows = rows_by_contig(solve_fn(batch.select(mine), **opts)) if len(mine) else []
    all_rows = gather(rows) if gather is not None else [rows]
    return merge_shards(batch.n_ctg, shards, all_rows)
