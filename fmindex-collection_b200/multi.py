"""Multi-GPU host logic: the index is replicated per GPU, queries are sharded contiguously, results are
gathered on the host; there is NO collective on the search path.  One process per GPU (torchrun): every rank searches its own
shard with its own index replica, rank 0 receives the located rows of all ranks.  The C++ counterpart for a single process
driving several GPUs is include/fmb200/multi.hpp (same shard_range)."""
import numpy as np


def shard_range(count, rank, world):
    """contiguous shard [begin, end) of `count` items for `rank` of `world`; sizes differ by at most one"""
    base, rest = divmod(count, world)
    begin = rank * base + min(rank, rest)
    return begin, begin + base + (1 if rank < rest else 0)


def shard_queries(symbols, offsets, rank, world):
    """the rank's slice of a flattened query batch: (symbols, offsets rebased to 0, first qidx)"""
    nq = len(offsets) - 1
    b, e = shard_range(nq, rank, world)
    off = np.asarray(offsets[b:e + 1], dtype=np.uint64)
    sym = np.asarray(symbols[int(off[0]):int(off[-1])], dtype=np.uint8)
    return sym, off - off[0], b


def gather_rows(rows, first_qidx, dist=None, dst=0):
    """Shift the rank-local qidx of a structured result array (hits or located rows) by the shard's first qidx and gather all
    ranks' rows on rank `dst` (torch.distributed gather_object: host-side concatenation, not a data-path collective).
    Without a process group the shifted rows are returned as they are."""
    rows = rows.copy()
    rows["qidx"] += first_qidx
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return rows
    parts = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(rows, parts, dst=dst)
    if dist.get_rank() != dst:
        return None
    return np.concatenate(parts)


def search_sharded(search_fn, symbols, offsets, dist=None):
    """run `search_fn(symbols, offsets) -> structured rows` on this rank's shard and gather the rows on rank 0"""
    rank = dist.get_rank() if dist is not None and dist.is_initialized() else 0
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    sym, off, first = shard_queries(symbols, offsets, rank, world)
    return gather_rows(search_fn(sym, off), first, dist)
