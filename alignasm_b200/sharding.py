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


# ---- one process per GPU: the rows of every rank's shard brought to rank 0 (bench.py --gpus N, torchrun) -------------------
ROW_KEYS = (("ctg_index", np.int32), ("qry_str", np.int64), ("qry_end", np.int64), ("ref_str", np.int64), ("ref_end", np.int64),
            ("is_alt", np.uint8))


def packed_size(result):
    n_ctg, n_out, n_alt = result.n_ctg, int(result.out_off[-1]), int(result.alt_off[-1])
    per_row = sum((np.dtype(dt).itemsize) for _, dt in ROW_KEYS)
    return 32 + 16 * (n_ctg + 1) + (n_out + n_alt) * per_row + 8 * 2 * len(ROW_KEYS)


def pack_rows_into(result, buf):
    """The primary and alternative row lists of a Result (per-contig offsets + the six row arrays each) written into the byte
    array `buf`: header int64[4] = (contigs, primary rows, alt rows, 0), then out_off, alt_off, then the arrays (8-byte
    aligned).  Returns the bytes used."""
    n_ctg, n_out, n_alt = result.n_ctg, int(result.out_off[-1]), int(result.alt_off[-1])
    at = 0

    def put(a):
        nonlocal at
        v = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
        buf[at:at + len(v)] = v
        at += (len(v) + 7) // 8 * 8

    put(np.array([n_ctg, n_out, n_alt, 0], dtype=np.int64))
    put(np.asarray(result.out_off, dtype=np.int64))
    put(np.asarray(result.alt_off, dtype=np.int64))
    for which in (result.out, result.alt):
        for k, dt in ROW_KEYS:
            put(np.asarray(which[k], dtype=dt))
    return at


def pack_rows(result):
    buf = np.zeros(packed_size(result), dtype=np.uint8)
    n = pack_rows_into(result, buf)
    return buf[:n]


def unpack_rows(buf):
    """Inverse of pack_rows: (out_off, alt_off, out, alt) as views of `buf`."""
    n_ctg, n_out, n_alt, _ = np.frombuffer(buf, dtype=np.int64, count=4).tolist()
    at = 32
    out_off = np.frombuffer(buf, dtype=np.int64, count=n_ctg + 1, offset=at)
    at += 8 * (n_ctg + 1)
    alt_off = np.frombuffer(buf, dtype=np.int64, count=n_ctg + 1, offset=at)
    at += 8 * (n_ctg + 1)
    lists = []
    for n in (n_out, n_alt):
        rows = {}
        for k, dt in ROW_KEYS:
            nbytes = n * np.dtype(dt).itemsize
            rows[k] = np.frombuffer(buf, dtype=dt, count=n, offset=at)
            at += (nbytes + 7) // 8 * 8
        lists.append(rows)
    return out_off, alt_off, lists[0], lists[1]


def contig_index(n_ctg, shards):
    """Where the rows of every input contig are after a gather: (shard, position inside the shard), so that a writer walks the
    contigs in input order (alignasm.cpp:417-441) without the shards being copied into one array."""
    shard_of = np.full(n_ctg, -1, dtype=np.int32)
    local = np.zeros(n_ctg, dtype=np.int64)
    for k, ids in enumerate(shards):
        shard_of[ids] = k
        local[ids] = np.arange(len(ids))
    assert (shard_of >= 0).all(), "a contig was not assigned to any shard"
    return shard_of, local


def gather_packed(packed, dist, rank, world, device=None):
    """torch.distributed gather of one byte array per rank to rank 0 (NCCL: through device buffers; gloo: host tensors).
    Returns the list of arrays on rank 0, None elsewhere."""
    import torch
    if world == 1 or dist is None:
        return [packed]
    dev = device if device is not None else "cpu"
    sizes = torch.zeros(world, dtype=torch.int64, device=dev)
    mine = torch.tensor([len(packed)], dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(sizes, mine) if dev != "cpu" else dist.all_gather(list(sizes.split(1)), mine)
    sizes = sizes.tolist()
    cap = max(sizes)
    buf = torch.empty(cap, dtype=torch.uint8, device=dev)
    buf[:len(packed)].copy_(torch.from_numpy(packed), non_blocking=True)
    out = [torch.empty(cap, dtype=torch.uint8, device=dev) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, out, dst=0)
    if rank != 0:
        return None
    return [o[:n].cpu().numpy() for o, n in zip(out, sizes)]


class ShmRows:
    """Host-side gather for one process per GPU on one box: every rank publishes the rows of its shard in a POSIX shared-memory
    segment of its own, rank 0 maps all of them.  No collective moves rows; the caller brackets publish / read with a barrier."""

    def __init__(self, tag, rank, world, capacity):
        from multiprocessing import shared_memory
        self.rank, self.world, self.tag = rank, world, tag
        name = f"{tag}_{rank}"
        try:  # a stale segment of a crashed run
            old = shared_memory.SharedMemory(name=name)
            old.close()
            old.unlink()
        except FileNotFoundError:
            pass
        self.mine = shared_memory.SharedMemory(name=name, create=True, size=int(capacity))
        self.buf = np.ndarray((self.mine.size,), dtype=np.uint8, buffer=self.mine.buf)
        self.others = None

    def attach_all(self):
        """Rank 0, after every rank has created its segment (barrier in between)."""
        from multiprocessing import resource_tracker, shared_memory
        self.others = []
        for r in range(self.world):
            if r == self.rank:
                self.others.append(self.mine)
                continue
            m = shared_memory.SharedMemory(name=f"{self.tag}_{r}")
            try:  # the segment belongs to rank r, which unlinks it: keep this process's tracker out of it
                resource_tracker.unregister(m._name, "shared_memory")
            except Exception:
                pass
            self.others.append(m)

    def publish(self, result):
        if packed_size(result) > self.mine.size:
            raise RuntimeError("shard rows outgrew the shared segment")
        return pack_rows_into(result, self.buf)

    def publish_empty(self):
        self.buf[:32].view(np.int64)[:] = 0

    def views(self):
        """Rank 0: (out_off, alt_off, out, alt) of every rank, as views of the shared segments."""
        return [unpack_rows(np.ndarray((m.size,), dtype=np.uint8, buffer=m.buf)) for m in self.others]

    def close(self):
        self.buf = None
        if self.others:
            for m in self.others:
                if m is not self.mine:
                    m.close()
        self.others = None
        try:
            self.mine.close()
            self.mine.unlink()
        except (FileNotFoundError, BufferError):
            pass
