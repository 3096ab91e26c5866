"""GPU parity: device occurrence table (String_c), cursor steps, exact search and locate against the oracle."""
import numpy as np
import pytest

from helpers import hits_equal, locs_equal, make_index_pair, strip_lb_rev

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pair(gpu):
    from fmb200 import synth
    text = synth.multi_text([3000, 2000, 500, 1, 64, 63], 5, 7)
    o, g = make_index_pair(gpu, text, 5, 4)
    return text, o, g


def test_index_info_and_C(pair):
    text, o, g = pair
    info = g.info
    assert info.n == text.size and info.sigma == 5 and info.bidirectional == 1
    assert info.n_delims == 6 and info.occ_block_bytes == 32 and info.occ_block_rows == 64
    assert np.array_equal(g.C, o.C)


def test_export_roundtrip(pair):
    text, o, g = pair
    bwt, rev, bm, sq, sp = g.export()
    obm, osq, osp = o.samples
    assert np.array_equal(bwt, o.bwt) and np.array_equal(rev, o.bwt_rev)
    assert np.array_equal(bm, obm) and np.array_equal(sq, osq) and np.array_equal(sp, osp)


@pytest.mark.parametrize("dir", [0, 1])
def test_string_ops_every_row(pair, dir):
    text, o, g = pair
    n = o.n
    rows = np.arange(n, dtype=np.uint64)
    assert np.array_equal(g.symbol(rows, dir), np.array([o.symbol(i, dir) for i in range(n)], dtype=np.uint8))
    rows1 = np.arange(n + 1, dtype=np.uint64)
    for s in range(5):
        symb = np.full(n + 1, s, dtype=np.uint8)
        assert np.array_equal(g.rank(rows1, symb, dir), np.array([o.rank(i, s, dir) for i in range(n + 1)], dtype=np.uint64)), s
    for s in range(6):   # prefix_rank accepts symb == sigma (string/FlattenedBitvectors2L.h:226-228)
        symb = np.full(n + 1, s, dtype=np.uint8)
        assert np.array_equal(g.prefix_rank(rows1, symb, dir), np.array([o.prefix_rank(i, s, dir) for i in range(n + 1)], dtype=np.uint64)), s
    sel = rows1[:: 37]
    rs, prs = g.all_ranks(sel, dir)
    for k, i in enumerate(sel):
        ors, oprs = o.all_ranks_and_prefix_ranks(int(i), dir)
        assert np.array_equal(rs[k], ors) and np.array_equal(prs[k], oprs)


def test_cursor_extend(pair):
    text, o, g = pair
    rng = np.random.default_rng(5)
    # cursors reached by real searches: start from the root and extend randomly, keep the non-empty ones
    curs = [np.array([0, 0, o.n, 0], dtype=np.uint64)]
    for _ in range(400):
        c = curs[rng.integers(0, len(curs))]
        nc = o.extend(c, int(rng.integers(0, 5)), bool(rng.integers(0, 2)))
        if nc[2] > 0:
            curs.append(nc)
    curs = np.array(curs, dtype=np.uint64)
    for right in (0, 1):
        for s in range(5):
            got = g.extend(curs, np.full(len(curs), s, dtype=np.uint8), right)
            exp = np.array([o.extend(c, s, right) for c in curs], dtype=np.uint64)
            assert np.array_equal(got, exp), (right, s)
        got = g.extend_all(curs, right)
        exp = np.array([o.extend_all(c, right) for c in curs], dtype=np.uint64)
        assert np.array_equal(got, exp)


def test_exact_search_and_locate(pair):
    from fmb200 import synth
    from oracle.pyoracle import Counters
    text, o, g = pair
    rng = np.random.default_rng(11)
    reads = []
    for L in (1, 2, 5, 16, 17, 31, 32, 33, 50, 100):
        for _ in range(40):
            p = int(rng.integers(0, text.size - L))
            reads.append(text[p:p + L].copy())            # may contain delimiters: must behave like the reference
        for _ in range(10):
            reads.append(rng.integers(1, 5, size=L).astype(np.uint8))
    sym, off = synth.flatten(reads)
    q = g.upload(sym, off)
    res = g.search_exact(q)
    exp = o.search_exact(sym, off)
    got = res.hits()
    assert hits_equal(got, exp)
    assert np.all(np.diff(got["qidx"].astype(np.int64)) > 0)     # compaction keeps query order
    loc = g.locate(res)
    assert locs_equal(loc.locs(), o.locate(exp))
    assert np.array_equal(loc.locs32()["pos"].astype(np.uint64), loc.locs()["pos"])
    # the algorithmic work counters come from the one-symbol kernel; both kernels give the same hits
    assert res.stats.line_requests > 0
    g.set_exact_mode(1)
    res1 = g.search_exact(q)
    g.set_exact_mode(0)
    assert hits_equal(res1.hits(), exp)
    ctr = Counters()
    o.search_exact(sym, off, ctr)
    st = res1.stats
    assert st.extensions == ctr.extensions and st.occ_lookups == ctr.occ_lookups
    ctr = Counters()
    o.locate(exp, ctr)
    assert loc.stats.lf_steps == ctr.lf_steps


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("lengths", [[40000], [3000, 1, 2, 700, 5, 129, 128, 127], [300] * 50])
def test_exact_modes_multi_sequence(gpu, mode, lengths):
    """one-symbol and two-symbol (128-byte pair table) kernels on collections with many delimiters; reads that
    contain or straddle delimiters take the one-symbol path inside the two-symbol kernel"""
    from fmb200 import synth
    text = synth.multi_text(lengths, 5, 21)
    o, g = make_index_pair(gpu, text, 5, 4)
    g.set_exact_mode(mode)
    rng = np.random.default_rng(5)
    reads = []
    for L in (1, 2, 3, 4, 7, 8, 20, 21, 64):
        for _ in range(60):
            p = int(rng.integers(0, text.size - L))
            reads.append(text[p:p + L].copy())
        for _ in range(10):
            reads.append(rng.integers(0, 5, size=L).astype(np.uint8))      # symbol 0 inside a query
    # all 2-mers and 3-mers over {0..4}: every pair code, every special-row correction
    for a in range(5):
        for b in range(5):
            reads.append(np.array([a, b], dtype=np.uint8))
            for c in range(5):
                reads.append(np.array([a, b, c], dtype=np.uint8))
    sym, off = synth.flatten(reads)
    res = g.search_exact(g.upload(sym, off))
    exp = o.search_exact(sym, off)
    assert hits_equal(res.hits(), exp)
    assert locs_equal(g.locate(res).locs(), o.locate(exp))


def test_empty_inputs(pair):
    text, o, g = pair
    q = g.upload(np.zeros(0, dtype=np.uint8), np.zeros(1, dtype=np.uint64))
    res = g.search_exact(q)
    assert len(res) == 0 and len(g.locate(res)) == 0
    # one empty query: the single-query form returns the whole index (search/SearchNoErrors.h:13-26)
    q = g.upload(np.zeros(0, dtype=np.uint8), np.zeros(2, dtype=np.uint64))
    h = g.search_exact(q).hits()
    assert len(h) == 1 and h[0]["lb"] == 0 and h[0]["len"] == o.n and h[0]["steps"] == 0


def test_unidirectional_index(gpu):
    from fmb200 import synth
    text = synth.multi_text([5000], 5, 3)
    o, g = make_index_pair(gpu, text, 5, 16, bidirectional=False)
    assert g.info.bidirectional == 0
    reads, _ = synth.reads_from_text(text, 300, 25, 9)
    sym, off = synth.flatten(reads)
    res = g.search_exact(g.upload(sym, off))
    exp = o.search_exact(sym, off)
    assert hits_equal(res.hits(), exp)
    assert locs_equal(g.locate(res).locs(), o.locate(exp))


def test_larger_single_sequence(gpu):
    """1 Mbp single sequence, rate 16: every read is found, every located position equals its source offset"""
    from fmb200 import synth
    text = synth.text(1 << 20, 5, 1)
    o, g = make_index_pair(gpu, text, 5, 16)
    reads, src = synth.reads_from_text(text, 20000, 100, 2)
    sym, off = synth.flatten(reads)
    res = g.search_exact(g.upload(sym, off))
    exp = o.search_exact(sym, off)
    assert hits_equal(res.hits(), exp)
    loc = g.locate(res).locs()
    assert locs_equal(loc, o.locate(exp))
    first = {}
    for r in loc:
        first.setdefault(int(r["qidx"]), set()).add(int(r["pos"]))
    assert all(int(src[q]) in first[q] for q in range(len(src)))


def test_locate_modes_and_irregular_sampling(gpu):
    """locate shortcut table vs LF walk, and a sparse array whose walks are too long for the shortcut word (table dropped)"""
    from fmb200 import synth
    from oracle.pyoracle import Oracle
    text = synth.multi_text([6000, 500], 5, 31)
    o, g = make_index_pair(gpu, text, 5, 16)
    reads, _ = synth.reads_from_text(text[:6001], 400, 12, 4)
    sym, off = synth.flatten(reads)
    res = g.search_exact(g.upload(sym, off))
    exp = o.locate(o.search_exact(sym, off))
    fast = g.locate(res)
    g.set_locate_mode(1)
    walk = g.locate(res)
    g.set_locate_mode(0)
    assert locs_equal(fast.locs(), exp) and locs_equal(walk.locs(), exp)
    assert fast.stats.lf_steps == walk.stats.lf_steps > 0
    rows = np.arange(0, text.size, 7, dtype=np.uint64)
    seq, pos, steps = g.locate_rows(rows)
    for r, s, p, k in zip(rows[:200], seq, pos, steps):
        assert o.locate_row(int(r)) == (int(s), int(p), int(k))
    # SA-space sampling of every 97th row only: walks of up to ~100 steps, the sample of a row is no longer "pos % rate == 0"
    bwt, rev = o.bwt, o.bwt_rev
    sa = o.sa
    keep = np.zeros(text.size, dtype=bool)
    keep[::97] = True
    delim = np.flatnonzero(text == 0)
    seq_of = np.searchsorted(delim, sa[keep], side="left").astype(np.uint32)
    start = np.concatenate([[0], delim + 1])
    pos_of = (sa[keep] - start[seq_of]).astype(np.uint32)
    bm = np.zeros((text.size + 63) // 64, dtype=np.uint64)
    idx = np.flatnonzero(keep)
    np.bitwise_or.at(bm, idx // 64, np.uint64(1) << (idx % 64).astype(np.uint64))
    g2 = gpu.Index.from_bwt(5, bwt, rev, bm, seq_of, pos_of)
    o2 = Oracle.from_bwt(5, bwt, rev, bm, seq_of, pos_of)
    res2 = g2.search_exact(g2.upload(sym, off))
    exp2 = o2.locate(o2.search_exact(sym, off))
    assert locs_equal(g2.locate(res2).locs(), exp2)
    # (with SA-space sampling a walk may cross a delimiter, where the reference's pos + steps arithmetic leaves the sequence: the
    #  device reproduces the reference, the positions are not comparable with the text-order sampling above)


def test_reverse_complement_doubling_on_the_device(gpu):
    """fmb_queries_upload_revcomp: the device batch = every read followed by its reverse complement (example/utils.h:62-74),
    compared with the same doubling done on the host; ragged lengths, an empty read, a read holding a delimiter"""
    from fmb200 import schemes, synth
    text = synth.multi_text([5000, 800], 5, 13)
    o, g = make_index_pair(gpu, text, 5, 4)
    comp = np.array([0, 4, 3, 2, 1], dtype=np.uint8)
    rng = np.random.default_rng(5)
    reads = []
    for i in range(400):
        L = int(rng.integers(1, 60)) if i % 50 else 0
        p = int(rng.integers(0, 5000 - 60))
        r = text[p:p + L].copy()
        if i % 3 == 0 and L:
            r = comp[r[::-1]]                     # a read from the other strand: its reverse complement is the one that matches
        reads.append(r)
    reads[7] = np.array([1, 2, 0, 3, 4, 4, 1], dtype=np.uint8)
    doubled = []
    for r in reads:
        doubled += [r, comp[r[::-1]]]
    sym, off = synth.flatten(reads)
    dsym, doff = synth.flatten(doubled)
    q = g.upload(sym, off, complement=comp)
    assert len(q) == 2 * len(reads)
    exp = o.search_exact(dsym, doff)
    got = g.search_exact(q).hits()
    assert hits_equal(strip_lb_rev(got), strip_lb_rev(exp)) and len(exp) > 400
    assert hits_equal(strip_lb_rev(g.search_exact(g.upload(dsym, doff)).hits()), strip_lb_rev(exp))
    # equal-length batch through a scheme search
    eq = [text[p:p + 30].copy() for p in rng.integers(0, 4900, 100)]
    eq = [comp[r[::-1]] if i % 2 else r for i, r in enumerate(eq)]
    eq = list(synth.plant_errors(np.array(eq, dtype=np.uint8), 5, 1, True, 3))
    d2 = []
    for r in eq:
        d2 += [r, comp[r[::-1]]]
    sch = schemes.optimum(0, 1)
    part = schemes.uniform_partition(2, 30)
    es, eo = synth.flatten(d2)
    assert hits_equal(g.search_scheme(g.upload(*synth.flatten(eq), complement=comp), sch, part, True).hits(), o.search_ng26(es, eo, sch, part, True))
