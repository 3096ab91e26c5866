"""GPU tests of the one-call end-to-end path (csrc/fmb_engine.cu): fmb_search_and_locate with several chunks in flight (rows come
back in chunk order), its 2-bit packed variant, fmb_queries_upload_packed, and the multi-replica call (one replica on a 1-GPU box;
two replicas -- fmb_index_replicate over peer memory -- when the box has two GPUs)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import hits_equal, locs_equal, make_index_pair

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def dna(gpu):
    from fmb200 import synth
    text = synth.multi_text([60000, 9000, 300], 5, 41)
    o, g = make_index_pair(gpu, text, 5, 8)
    reads, _ = synth.reads_from_text(text[:60001], 3000, 40, 6)
    reads[1500:] = synth.plant_errors(reads[1500:], 5, 1, True, 8)
    sym, off = synth.flatten(reads)
    return text, o, g, sym, off


def test_packed_upload_equals_byte_upload(gpu, dna):
    from fmb200 import capi, schemes
    text, o, g, sym, off = dna
    sym = sym.copy()
    sym[[5, 777, 40 * 100 + 3]] = [0, 0, 0]              # symbols without 2-bit code: exception list, flagged queries take the byte path
    packed = capi.pack_queries(sym, 5)
    assert packed[1].size == 3
    qb, qp = g.upload(sym, off), g.upload(None, off, packed=packed)
    assert hits_equal(g.search_exact(qp).hits(), g.search_exact(qb).hits())
    assert hits_equal(g.search_exact(qp).hits(), o.search_exact(sym, off))
    bad = sym.copy()
    bad[40 * 7 + 1] = 9                                  # >= sigma: such a query matches nothing (exact search)
    assert hits_equal(g.search_exact(g.upload(None, off, packed=capi.pack_queries(bad, 5))).hits(), g.search_exact(g.upload(bad, off)).hits())
    sch, part = schemes.optimum(0, 1), schemes.uniform_partition(2, 40)
    assert hits_equal(g.search_scheme(qp, sch, part, True).hits(), o.search_ng26(sym, off, sch, part, True))
    # a slice of the batch that starts inside a packed word
    sl = slice(7, 1234)
    qs = g.upload(None, off[sl.start: sl.stop + 1], packed=packed)
    exp = o.search_exact(sym[int(off[sl.start]): int(off[sl.stop])], off[sl.start: sl.stop + 1] - off[sl.start])
    assert hits_equal(g.search_exact(qs).hits(), exp)


def test_packed_one_call_path(gpu, dna):
    from fmb200 import capi, schemes
    text, o, g, sym, off = dna
    packed = capi.pack_queries(sym, 5)
    sch, part = schemes.optimum(0, 1), schemes.uniform_partition(2, 40)
    for scheme, edit in ((None, False), (sch, True), (sch, False)):
        a, _ = g.search_and_locate(sym, off, scheme=scheme, partition=part, edit=edit)
        b, _ = g.search_and_locate(None, off, scheme=scheme, partition=part, edit=edit, packed=packed)
        exp = o.locate(o.search_exact(sym, off) if scheme is None else o.search_ng26(sym, off, scheme, part, edit))
        assert locs_equal(a, exp) and locs_equal(b, exp)


def test_packed_queries_need_a_dna_index(gpu):
    from fmb200 import capi, synth
    text = synth.multi_text([2000], 21, 3)
    o, g = make_index_pair(gpu, text, 21, 8)
    with pytest.raises(gpu.FmbError) as e:
        g.upload(None, np.array([0, 4], dtype=np.uint64), packed=capi.pack_queries(np.array([1, 2, 3, 4], dtype=np.uint8)))
    assert e.value.code == -5


def test_rows_come_back_in_chunk_order():
    """many small chunks on six streams (fresh process: the chunk size is read once): qidx never decreases across the output except
    inside a chunk of 2^6 queries, results equal the oracle's"""
    code = r'''
import os, sys
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import numpy as np
import fmb200 as fmb
from fmb200 import schemes, synth
from helpers import locs_equal, make_index_pair
text = synth.multi_text([40000, 5000], 5, 13)
o, g = make_index_pair(fmb, text, 5, 8)
reads, _ = synth.reads_from_text(text[:40001], 2000, 36, 3)
reads[1000:] = synth.plant_errors(reads[1000:], 5, 1, True, 4)
sym, off = synth.flatten(reads)
sch, part = schemes.optimum(0, 1), schemes.uniform_partition(2, 36)
for scheme, edit in ((None, False), (sch, True)):
    for rep in range(3):
        locs, st = g.search_and_locate(sym, off, scheme=scheme, partition=part, edit=edit)
        chunk = locs["qidx"] >> 6
        assert np.all(np.diff(chunk.astype(np.int64)) >= 0), "rows are not grouped by ascending chunks of qidx"
        exp = o.locate(o.search_exact(sym, off) if scheme is None else o.search_ng26(sym, off, scheme, part, edit))
        assert locs_equal(locs, exp)
# capacity too small: FMB_EOVERFLOW and the size that is needed
try:
    g.search_and_locate(sym, off, capacity=10)
    raise SystemExit("no overflow error")
except fmb.FmbError as e:
    assert e.code == -6 and "rows found" in str(e), str(e)
print("ok")
''' % (ROOT, ROOT)
    env = dict(os.environ, FMB_E2E_CHUNK_LOG2="6")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


def _check_multi(gpu, replicas, o, sym, off):
    from fmb200 import capi, schemes
    sch, part = schemes.optimum(0, 1), schemes.uniform_partition(2, 40)
    nq = off.size - 1
    G = len(replicas)
    per = (nq + G - 1) // G
    for scheme, edit in ((None, False), (sch, True)):
        parts, st = capi.search_and_locate_multi(replicas, sym, off, scheme=scheme, partition=part, edit=edit)
        assert len(parts) == G
        for g, rows in enumerate(parts):            # shard g holds exactly the queries [g * per, (g + 1) * per)
            assert rows.size == 0 or (rows["qidx"].min() >= g * per and rows["qidx"].max() < min(nq, (g + 1) * per))
        exp = o.locate(o.search_exact(sym, off) if scheme is None else o.search_ng26(sym, off, scheme, part, edit))
        assert locs_equal(np.concatenate(parts), exp)
    with pytest.raises(gpu.FmbError) as e:           # a shard that does not fit: the call says so
        capi.search_and_locate_multi(replicas, sym, off, shard_capacity=3)
    assert e.value.code == -6


def test_multi_call_with_one_replica(gpu, dna):
    text, o, g, sym, off = dna
    _check_multi(gpu, [g], o, sym, off)


def test_replica_on_the_same_device_is_refused(gpu, dna):
    text, o, g, sym, off = dna
    with pytest.raises(gpu.FmbError) as e:
        g.replicate(0)
    assert e.value.code == -1


def test_replicated_index_and_sharded_call_on_two_gpus(gpu, dna):
    if gpu.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    text, o, g, sym, off = dna
    r = g.replicate(1)
    assert r.info.device == 1 and r.info.tables == g.info.tables and r.info.device_bytes == g.info.device_bytes
    a, b = g.export(), r.export()
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    assert hits_equal(r.search_exact(r.upload(sym, off)).hits(), o.search_exact(sym, off))
    _check_multi(gpu, [g, r], o, sym, off)


def test_collection_split_into_parts_returns_the_rows_of_one_index(gpu):
    """fmb_search_and_locate_parts (the capacity path for n >= 2^32 rows): sequences split into three indices, the whole batch searched in
    every part, sequence numbers of the whole collection -- the located rows equal those of one index over everything"""
    from fmb200 import capi, schemes, synth
    lens = [9000, 400, 7000, 30, 5000, 6000, 1200]
    text = synth.multi_text(lens, 5, 77)
    o, whole = make_index_pair(gpu, text, 5, 8)
    ends = np.flatnonzero(text == 0) + 1
    cuts = [0, int(ends[1]), int(ends[4]), int(ends[6])]          # parts of 2, 3 and 2 sequences
    seq_base = [0, 2, 5]
    parts = [gpu.Index.build(5, text[a:b], sampling_rate=8, device=0) for a, b in zip(cuts[:-1], cuts[1:])]
    reads, _ = synth.reads_from_text(text, 1500, 40, 9)
    reads = [r for r in reads if 0 not in r][:1200]
    reads[600:] = synth.plant_errors(np.array(reads[600:], dtype=np.uint8), 5, 1, True, 4)
    sym, off = synth.flatten(reads)
    sch, part = schemes.optimum(0, 1), schemes.uniform_partition(2, 40)
    for scheme, edit in ((None, False), (sch, True), (sch, False)):
        rows, st = capi.search_and_locate_parts(parts, seq_base, sym, off, scheme=scheme, partition=part, edit=edit)
        assert len(rows) == 3 and all(r.size for r in rows)
        for p, r in enumerate(rows):                    # part p only reports its own sequences
            hi = seq_base[p + 1] if p + 1 < 3 else len(lens)
            assert r["seq"].min() >= seq_base[p] and r["seq"].max() < hi
        exp = o.locate(o.search_exact(sym, off) if scheme is None else o.search_ng26(sym, off, scheme, part, edit))
        assert locs_equal(np.concatenate(rows), exp)
        one, _ = whole.search_and_locate(sym, off, scheme=scheme, partition=part, edit=edit)
        assert locs_equal(np.concatenate(rows), one)
    with pytest.raises(gpu.FmbError) as e:
        capi.search_and_locate_parts(parts, seq_base, sym, off, part_capacity=5)
    assert e.value.code == -6
    with pytest.raises(gpu.FmbError):
        capi.search_and_locate_parts([parts[0], parts[0]], [0, 2], sym, off)
