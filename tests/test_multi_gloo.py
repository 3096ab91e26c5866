"""N > 1 host logic on CPU: world_size-2 gloo process group, contiguous query shards, gather of the rows on rank 0.
The per-rank "engine" here is the oracle (the checker standing in for a GPU replica -- no GPU in this test); what is tested
is the sharding / qidx rebasing / gather code that bench.py and fmb200.multi use around the device calls."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import fmb200  # noqa: F401
    from fmb200 import multi, schemes, synth
    from oracle.pyoracle import Oracle, sort_hits, sort_locs
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    text = synth.multi_text([20000, 3000], 5, 1)            # every rank builds the same replica
    o = Oracle.build(text, 5, 8, True)
    reads, _ = synth.reads_from_text(text[:20001], 101, 40, 2)        # 101 reads: shards of unequal size
    reads[50:] = synth.plant_errors(reads[50:], 5, 1, True, 3)
    sym, off = synth.flatten(reads)
    sch = schemes.optimum(0, 1)
    part = schemes.uniform_partition(2, 40)

    def engine(s, o_):
        return o.locate(o.search_ng26(s, o_, sch, part, True))

    rows = multi.search_sharded(engine, sym, off, dist)
    hits = multi.search_sharded(lambda s, o_: o.search_exact(s, o_), sym, off, dist)
    if rank == 0:
        exp = o.locate(o.search_ng26(sym, off, sch, part, True))
        exp_hits = o.search_exact(sym, off)
        ok = np.array_equal(sort_locs(rows), sort_locs(exp)) and np.array_equal(sort_hits(hits), sort_hits(exp_hits))
        with open(out_path, "w") as f:
            f.write("ok" if ok and len(exp) > 50 else "mismatch")
    else:
        assert rows is None and hits is None
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_everything():
    sys.path.insert(0, ROOT)
    import fmb200  # noqa: F401
    from fmb200.multi import shard_range
    for count in (0, 1, 7, 8, 9, 1000003):
        for world in (1, 2, 3, 4, 8):
            edges = [shard_range(count, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == count
            assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
            sizes = [e - b for b, e in edges]
            assert max(sizes) - min(sizes) <= 1


def test_two_ranks_gloo(tmp_path):
    import torch.multiprocessing as mp
    out = tmp_path / "result.txt"
    mp.spawn(_worker, args=(2, _free_port(), str(out)), nprocs=2, join=True)
    assert out.read_text() == "ok"
