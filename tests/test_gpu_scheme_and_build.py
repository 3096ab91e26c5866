"""GPU parity: k-error search-scheme search (Hamming + Edit), backtracking, GPU index construction."""
import numpy as np
import pytest

from helpers import hits_equal, locs_equal, make_index_pair

pytestmark = pytest.mark.gpu


def _remap(s):
    return np.array([{"A": 1, "B": 2, "C": 3, "D": 4}[c] for c in s], dtype=np.uint8)


def _located(g, res):
    return sorted((int(r["qidx"]), int(r["seq"]), int(r["pos"])) for r in g.locate(res).locs())


@pytest.fixture(scope="module")
def golden_pair(gpu):
    """the fixture of the reference's search tests (search/checkSearches.cpp:14-21): two sequences over {A,B,C},
    sampling rate 1; the alphabet is remapped order-preservingly to 1..4 (positions do not depend on it)"""
    text = np.concatenate([_remap("AAACAAABAAA"), [0], _remap("AAABAAACAAA"), [0]]).astype(np.uint8)
    return make_index_pair(gpu, text, 5, 1)


def test_golden_checkSearches(golden_pair):
    from fmb200 import schemes, synth
    o, g = golden_pair
    q1 = synth.flatten([_remap("CC"), _remap("BB")])
    q2 = synth.flatten([_remap("CD"), _remap("DB")])
    pigeon_opt = (np.array([[0, 1], [1, 0]]), np.array([[0, 0], [0, 1]]), np.array([[0, 1], [0, 1]]))  # generator/pigeon.h, k=1
    part = schemes.uniform_partition(2, 2)
    ham = [(0, 0, 2), (0, 0, 3), (0, 1, 6), (0, 1, 7), (1, 0, 6), (1, 0, 7), (1, 1, 2), (1, 1, 3)]
    # search/checkSearches.cpp:1093-1120  ng26, edit, queries CD / DB
    assert _located(g, g.search_scheme(g.upload(*q2), pigeon_opt, part, edit=True)) == [(0, 0, 3), (0, 1, 7), (1, 0, 7), (1, 1, 3)]
    # :1174-1199 hamming, queries CC / BB
    assert _located(g, g.search_scheme(g.upload(*q1), pigeon_opt, part, edit=False)) == ham
    # :23-102 backtracking k=1
    assert _located(g, g.search_backtracking(g.upload(*q1), 1)) == ham
    # :104-117 no errors: nothing
    assert _located(g, g.search_exact(g.upload(*q1))) == []
    # :1422-1444 facade fmc::search<true>(index, queries CC/BB, 1): h2 scheme, short form for length-2 queries
    sch, p = schemes.facade_scheme(True, 1, 2)
    exp = [(0, 0, 3), (0, 0, 3), (0, 1, 7), (0, 1, 7), (1, 0, 7), (1, 0, 7), (1, 1, 3), (1, 1, 3)]
    assert _located(g, g.search_scheme(g.upload(*q1), sch, p, edit=True)) == exp
    # :1468-1490 facade hamming
    sch, p = schemes.facade_scheme(False, 1, 2)
    assert _located(g, g.search_scheme(g.upload(*q1), sch, p, edit=False)) == ham


@pytest.fixture(scope="module")
def pair(gpu):
    from fmb200 import synth
    text = synth.multi_text([6000, 3000, 700, 5], 5, 21)
    o, g = make_index_pair(gpu, text, 5, 8)
    return text, o, g


@pytest.mark.parametrize("k", [1, 2])
@pytest.mark.parametrize("edit", [False, True])
@pytest.mark.parametrize("planted", [0, 1, 2])
def test_scheme_search_matches_oracle(pair, k, edit, planted):
    from fmb200 import schemes, synth
    text, o, g = pair
    L = 40
    reads, _ = synth.reads_from_text(text[:6000], 300, L, 100 + planted)
    if planted:
        reads = synth.plant_errors(reads, 5, planted, edit, 7 + k)
    sym, off = synth.flatten(reads)
    sch = schemes.optimum(0, k)
    part = schemes.uniform_partition(sch[0].shape[1], L)
    res = g.search_scheme(g.upload(sym, off), sch, part, edit)
    exp = o.search_ng26(sym, off, sch, part, edit)
    assert len(exp) > 0
    assert hits_equal(res.hits(), exp)
    assert locs_equal(g.locate(res).locs(), o.locate(exp))


@pytest.mark.parametrize("edit", [False, True])
def test_scheme_search_repetitive_text(gpu, edit):
    """intervals stay wider than one row for a long time: exercises search_next_dir (all-symbol extension)"""
    from fmb200 import schemes, synth
    rng = np.random.default_rng(3)
    unit = rng.integers(1, 5, size=50).astype(np.uint8)
    seqs = []
    for _ in range(40):
        u = unit.copy()
        for p in rng.integers(0, 50, size=2):
            u[p] = rng.integers(1, 5)
        seqs.append(u)
    text = np.concatenate([np.concatenate([s, [0]]) for s in seqs]).astype(np.uint8)
    o, g = make_index_pair(gpu, text, 5, 4)
    reads = np.array([seqs[i % 40][5:35] for i in range(60)], dtype=np.uint8)
    reads = synth.plant_errors(reads, 5, 1, edit, 4)
    sym, off = synth.flatten(reads)
    for k in (1, 2):
        sch = schemes.optimum(0, k)
        part = schemes.uniform_partition(sch[0].shape[1], 30)
        res = g.search_scheme(g.upload(sym, off), sch, part, edit)
        exp = o.search_ng26(sym, off, sch, part, edit)
        assert hits_equal(res.hits(), exp)
        assert locs_equal(g.locate(res).locs(), o.locate(exp))


def test_facade_h2_schemes(pair):
    """fmc::search<Edit>(index, queries, k, cb) = ng26 with the h2 scheme (search/CachedSearchScheme.h:26)"""
    from fmb200 import schemes, synth
    text, o, g = pair
    L = 36
    reads, _ = synth.reads_from_text(text[:6000], 200, L, 5)
    reads = synth.plant_errors(reads, 5, 1, True, 9)
    sym, off = synth.flatten(reads)
    for k in (1, 2, 3):
        for edit in (False, True):
            sch, part = schemes.facade_scheme(edit, k, L)
            res = g.search_scheme(g.upload(sym, off), sch, part, edit)
            assert hits_equal(res.hits(), o.search_ng26(sym, off, sch, part, edit)), (k, edit)


@pytest.mark.parametrize("bidirectional", [True, False])
def test_backtracking(gpu, bidirectional):
    from fmb200 import synth
    text = synth.multi_text([4000, 1000], 5, 5)
    o, g = make_index_pair(gpu, text, 5, 4, bidirectional=bidirectional)
    reads, _ = synth.reads_from_text(text[:4000], 150, 14, 1)
    reads = synth.plant_errors(reads, 5, 1, False, 2)
    sym, off = synth.flatten(reads)
    for k in (0, 1, 2):
        res = g.search_backtracking(g.upload(sym, off), k)
        exp = o.search_backtracking(sym, off, k)
        assert hits_equal(res.hits(), exp), k
        assert locs_equal(g.locate(res).locs(), o.locate(exp))


def test_scheme_argument_errors(pair):
    from fmb200 import FmbError, schemes, synth
    text, o, g = pair
    sym, off = synth.flatten(np.ones((4, 10), dtype=np.uint8))
    q = g.upload(sym, off)
    sch = schemes.optimum(0, 1)
    with pytest.raises(FmbError):     # partition does not sum to the query length
        g.search_scheme(q, sch, schemes.uniform_partition(2, 12), False)
    bad = (np.array([[0, 0]]), sch[1][:1], sch[2][:1])
    with pytest.raises(FmbError):     # pi not a permutation
        g.search_scheme(q, bad, schemes.uniform_partition(2, 10), False)


@pytest.mark.parametrize("lengths,rate", [([5000, 300, 1, 64], 4), ([20000], 16), ([7], 1), ([100] * 30, 3)])
@pytest.mark.parametrize("bidirectional", [True, False])
def test_gpu_index_build_matches_oracle(gpu, lengths, rate, bidirectional):
    """suffix sort / BWT / reverse BWT / text-space sampling on the GPU == the oracle's CPU construction
    (which follows utils.h:97-163 and BiFMIndex.h:82-135 and is pinned against the reference build)"""
    from fmb200 import synth
    from oracle.pyoracle import Oracle
    text = synth.multi_text(lengths, 5, 17)
    o = Oracle.build(text, 5, rate, bidirectional)
    g = gpu.Index.build(5, text, sampling_rate=rate, bidirectional=bidirectional)
    bwt, rev, bm, sq, sp = g.export()
    obm, osq, osp = o.samples
    assert np.array_equal(bwt, o.bwt)
    if bidirectional:
        assert np.array_equal(rev, o.bwt_rev)
    assert np.array_equal(bm, obm) and np.array_equal(sq, osq) and np.array_equal(sp, osp)
    assert np.array_equal(g.C, o.C)


def test_gpu_index_build_repetitive(gpu):
    """many prefix-doubling rounds: a text of few distinct long repeats"""
    from oracle.pyoracle import Oracle
    rng = np.random.default_rng(1)
    unit = rng.integers(1, 5, size=300).astype(np.uint8)
    text = np.concatenate([np.tile(unit, 20), [0], np.ones(2000, dtype=np.uint8), [0]]).astype(np.uint8)
    o = Oracle.build(text, 5, 8, True)
    g = gpu.Index.build(5, text, sampling_rate=8)
    bwt, rev, bm, sq, sp = g.export()
    assert np.array_equal(bwt, o.bwt) and np.array_equal(rev, o.bwt_rev)
    obm, osq, osp = o.samples
    assert np.array_equal(bm, obm) and np.array_equal(sq, osq) and np.array_equal(sp, osp)


def test_search_and_locate_one_call(pair):
    from fmb200 import schemes, synth
    text, o, g = pair
    reads, _ = synth.reads_from_text(text[:6000], 500, 40, 77)
    sym, off = synth.flatten(reads)
    locs, st = g.search_and_locate(sym, off)
    exp = o.locate(o.search_exact(sym, off))
    got = np.zeros(len(locs), dtype=exp.dtype)
    for f in ("qidx", "seq", "pos", "e"):
        got[f] = locs[f]
    assert locs_equal(got, exp)
    sch = schemes.optimum(0, 1)
    part = schemes.uniform_partition(2, 40)
    locs, st = g.search_and_locate(sym, off, sch, part, edit=True, capacity=100000)
    exp = o.locate(o.search_ng26(sym, off, sch, part, True))
    got = np.zeros(len(locs), dtype=exp.dtype)
    for f in ("qidx", "seq", "pos", "e"):
        got[f] = locs[f]
    assert locs_equal(got, exp)


@pytest.mark.parametrize("edit", [False, True])
def test_scheme_work_counters_match_oracle(gpu, edit):
    """fmb_stats.extensions of the scheme kernel = the oracle's count of cursor extensions of search_ng26 (the algorithmic work
    of SURVEY.md §8d), although the kernel covers sixteen of them with one LF^16 jump; occ_lookups agrees up to the rows whose
    interval end falls into the next block (a jump counts one lookup per step)."""
    from fmb200 import schemes, synth
    from oracle.pyoracle import Counters
    text = synth.text(300000, 5, 4)
    o, g = make_index_pair(gpu, text, 5, 16)
    reads, _ = synth.reads_from_text(text, 2000, 100, 5)
    reads[1000:] = synth.plant_errors(reads[1000:], 5, 1, edit, 6)
    sym, off = synth.flatten(reads)
    sch = schemes.optimum(0, 2)
    part = schemes.uniform_partition(4, 100)
    res = g.search_scheme(g.upload(sym, off), sch, part, edit)
    ctr = Counters()
    exp = o.search_ng26(sym, off, sch, part, edit, counters=ctr)
    assert hits_equal(res.hits(), exp)
    st = res.stats
    assert st.extensions == ctr.extensions
    assert abs(int(st.occ_lookups) - int(ctr.occ_lookups)) <= 0.05 * ctr.occ_lookups
    assert 0 < st.line_requests < st.occ_lookups          # the jumps save physical fetches


@pytest.mark.parametrize("L", [170, 400, 700])
def test_long_reads_cross_the_text_window(gpu, L):
    """reads longer than the text kernel's window (144 symbols) and query words (160 symbols): a direction run is then cut into several
    items (hand-overs of kind "continue", straight to the next text list); hits and extension counters still equal the oracle's"""
    from fmb200 import schemes, synth
    from oracle.pyoracle import Counters
    text = synth.multi_text([200000, 3000], 5, 14)
    o, g = make_index_pair(gpu, text, 5, 16)
    reads, _ = synth.reads_from_text(text[:200001], 400, L, 8)
    reads[100:250] = synth.plant_errors(reads[100:250], 5, 1, True, 3)
    reads[250:] = synth.plant_errors(reads[250:], 5, 2, True, 4)
    sym, off = synth.flatten(reads)
    q = g.upload(sym, off)
    for k in (1, 2):
        sch = schemes.optimum(0, k)
        part = schemes.uniform_partition(sch[0].shape[1], L)
        for edit in (False, True):
            ctr = Counters()
            exp = o.search_ng26(sym, off, sch, part, edit, counters=ctr)
            res = g.search_scheme(q, sch, part, edit)
            assert hits_equal(res.hits(), exp), (L, k, edit)
            assert res.stats.extensions == ctr.extensions, (L, k, edit)
            assert len(exp) >= 100
