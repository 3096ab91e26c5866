"""GPU parity for the generic occurrence-table layout (5 < sigma <= 32): protein alphabet sigma = 21 (BASELINE configs[4])
and a few other alphabet sizes, against the oracle."""
import numpy as np
import pytest

from helpers import hits_equal, locs_equal, make_index_pair

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def prot(gpu):
    from fmb200 import synth
    text = synth.multi_text([30000, 5000, 64, 1], 21, 17)
    o, g = make_index_pair(gpu, text, 21, 8)
    return text, o, g


def test_info_and_C(prot):
    text, o, g = prot
    i = g.info
    assert i.sigma == 21 and i.n == text.size and i.occ_block_bytes == 128 and i.n_delims == 4
    assert np.array_equal(g.C, o.C)


@pytest.mark.parametrize("dir", [0, 1])
def test_string_ops(prot, dir):
    text, o, g = prot
    rng = np.random.default_rng(3)
    idx = np.concatenate([np.arange(0, 300), rng.integers(0, text.size + 1, size=3000), [text.size]]).astype(np.uint64)
    symb = rng.integers(0, 21, size=idx.size).astype(np.uint8)
    assert np.array_equal(g.rank(idx, symb, dir), np.array([o.rank(i, s, dir) for i, s in zip(idx, symb)], dtype=np.uint64))
    symb2 = rng.integers(0, 22, size=idx.size).astype(np.uint8)          # prefix_rank accepts symb == sigma
    assert np.array_equal(g.prefix_rank(idx, symb2, dir), np.array([o.prefix_rank(i, s, dir) for i, s in zip(idx, symb2)], dtype=np.uint64))
    rows = idx[idx < text.size]
    assert np.array_equal(g.symbol(rows, dir), np.array([o.symbol(i, dir) for i in rows], dtype=np.uint8))
    rs, prs = g.all_ranks(idx[:200], dir)
    for k, i in enumerate(idx[:200]):
        ers, eprs = o.all_ranks_and_prefix_ranks(i, dir)
        assert np.array_equal(rs[k], ers) and np.array_equal(prs[k], eprs)


def test_export_roundtrip(prot):
    text, o, g = prot
    bwt, rev, bm, sq, sp = g.export()
    assert np.array_equal(bwt, o.bwt) and np.array_equal(rev, o.bwt_rev)


def test_exact_and_locate(prot):
    from fmb200 import synth
    text, o, g = prot
    rng = np.random.default_rng(4)
    reads = []
    for L in (1, 2, 3, 10, 50):
        for _ in range(100):
            p = int(rng.integers(0, text.size - L))
            reads.append(text[p:p + L].copy())
        for _ in range(20):
            reads.append(rng.integers(0, 21, size=L).astype(np.uint8))
    sym, off = synth.flatten(reads)
    res = g.search_exact(g.upload(sym, off))
    exp = o.search_exact(sym, off)
    assert hits_equal(res.hits(), exp)
    assert locs_equal(g.locate(res).locs(), o.locate(exp))


@pytest.mark.parametrize("edit", [False, True])
@pytest.mark.parametrize("k", [1, 2])
def test_scheme_search(prot, k, edit):
    from fmb200 import schemes, synth
    text, o, g = prot
    reads, _ = synth.reads_from_text(text[:30001], 300, 50, 5)
    reads[100:200] = synth.plant_errors(reads[100:200], 21, 1, edit, 6)
    reads[200:] = synth.plant_errors(reads[200:], 21, 2, edit, 7)
    sym, off = synth.flatten(reads)
    sch = schemes.optimum(0, k)
    part = schemes.uniform_partition(sch[0].shape[1], 50)
    res = g.search_scheme(g.upload(sym, off), sch, part, edit)
    exp = o.search_ng26(sym, off, sch, part, edit)
    assert hits_equal(res.hits(), exp)
    assert locs_equal(g.locate(res).locs(), o.locate(exp))
    assert len(exp) >= 100


def test_backtracking(prot):
    from fmb200 import synth
    text, o, g = prot
    reads, _ = synth.reads_from_text(text[:30001], 60, 12, 8)
    sym, off = synth.flatten(reads)
    res = g.search_backtracking(g.upload(sym, off), 1)
    assert hits_equal(res.hits(), o.search_backtracking(sym, off, 1))


@pytest.mark.parametrize("sigma", [6, 8, 21, 32])
def test_gpu_build_other_alphabets(gpu, sigma):
    from fmb200 import synth
    from oracle.pyoracle import Oracle
    text = synth.multi_text([4000, 333, 2], sigma, 9)
    g = gpu.Index.build(sigma, text, sampling_rate=4, bidirectional=True)
    o = Oracle.build(text, sigma, 4, True)
    bwt, rev, bm, sq, sp = g.export()
    assert np.array_equal(bwt, o.bwt) and np.array_equal(rev, o.bwt_rev)
    obm, osq, osp = o.samples
    assert np.array_equal(bm, obm) and np.array_equal(sq, osq) and np.array_equal(sp, osp)
    assert np.array_equal(g.C, o.C)
    reads, _ = synth.reads_from_text(text[:4001], 200, 15, 10)
    sym, off = synth.flatten(reads)
    res = g.search_exact(g.upload(sym, off))
    exp = o.search_exact(sym, off)
    assert hits_equal(res.hits(), exp)
    assert locs_equal(g.locate(res).locs(), o.locate(exp))


def test_sigma_out_of_range(gpu):
    with pytest.raises(gpu.FmbError):
        gpu.Index.build(33, np.array([1, 2, 0], dtype=np.uint8))
