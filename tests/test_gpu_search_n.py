"""GPU parity: hit-limited search (search_n: search_ng26::search(..., n), SearchNg26.h:408-433; fmc::search_n, search/search.h:37-45).

With a limit the reference reports, per query, the first n ROWS in the order its depth-first search finds them, clipping the
cursor that crosses the limit.  fmb_search_scheme_n must return exactly that list -- hits are compared IN ORDER, not as sets:
the oracle (and the reference) emit them by ascending qidx, then discovery order."""
import numpy as np
import pytest

from helpers import make_index_pair

pytestmark = pytest.mark.gpu


def _remap(s):
    return np.array([{"A": 1, "B": 2, "C": 3, "D": 4}[c] for c in s], dtype=np.uint8)


def _located(g, res):
    return sorted((int(r["qidx"]), int(r["seq"]), int(r["pos"])) for r in g.locate(res).locs())


def _same_list(got, exp):
    got, exp = np.asarray(got), np.asarray(exp)
    if got.shape != exp.shape:
        return False
    return all(np.array_equal(got[f].astype(np.uint64), exp[f].astype(np.uint64)) for f in ("qidx", "lb", "lb_rev", "len", "steps", "e"))


def test_golden_search_n(gpu):
    """search/checkSearches.cpp:1148-1172 (ng26, n = 3) and :1446-1466 (fmc::search_n<true>, n = 3)"""
    from fmb200 import schemes, synth
    text = np.concatenate([_remap("AAACAAABAAA"), [0], _remap("AAABAAACAAA"), [0]]).astype(np.uint8)
    o, g = make_index_pair(gpu, text, 5, 1)
    q = synth.flatten([_remap("CC"), _remap("BB")])
    pigeon_opt = (np.array([[0, 1], [1, 0]]), np.array([[0, 0], [0, 1]]), np.array([[0, 1], [0, 1]]))
    part = schemes.uniform_partition(2, 2)
    exp_n3 = [(0, 0, 3), (0, 1, 7), (0, 1, 7), (1, 0, 7), (1, 0, 7), (1, 1, 3)]
    assert _located(g, g.search_scheme(g.upload(*q), pigeon_opt, part, True, n=3)) == exp_n3
    sch, p = schemes.facade_scheme(True, 1, 2)
    assert _located(g, g.search_scheme(g.upload(*q), sch, p, True, n=3)) == exp_n3
    assert len(g.search_scheme(g.upload(*q), sch, p, True, n=0)) == 0                       # SearchNg26.h:410


def test_search_n_against_reference_output(gpu):
    """the reference's own hit-limited output, in order (tests/golden/ref_vectors.json, produced by tests/golden/make_golden.py from
    the compiled reference): the device must return exactly these lists"""
    import json
    import os
    from fmb200 import schemes, synth
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_vectors.json")) as f:
        golden = json.load(f)
    checked = 0
    for c in list(golden["cases"]) + [golden["hit_limit_case"]]:
        o, g = make_index_pair(gpu, np.array(c["text"], dtype=np.uint8), 5, c["rate"])
        sym, off = synth.flatten(np.array(c["queries"], dtype=np.uint8))
        q = g.upload(sym, off)
        for key, val in c["searches"].items():
            if "hits_in_order" not in val:
                continue
            _, _, kk, tag, nn = key.split("_")
            sch = schemes.optimum(0, int(kk[1:]))
            part = schemes.uniform_partition(sch[0].shape[1], c["L"])
            got = g.search_scheme(q, sch, part, tag == "edit", n=int(nn[1:])).hits()
            exp = np.array([tuple(r) for r in val["hits_in_order"]], dtype=got.dtype) if val["hits_in_order"] else got[:0]
            assert _same_list(got, exp), (c["name"], key)
            checked += len(exp)
    assert checked > 2000


@pytest.fixture(scope="module")
def pair(gpu):
    from fmb200 import synth
    text = synth.multi_text([6000, 3000, 700, 5], 5, 21)
    o, g = make_index_pair(gpu, text, 5, 8)
    return text, o, g


@pytest.mark.parametrize("k", [1, 2])
@pytest.mark.parametrize("edit", [False, True])
def test_search_n_random_text(pair, k, edit):
    from fmb200 import schemes, synth
    text, o, g = pair
    L = 40
    reads, _ = synth.reads_from_text(text[:6000], 300, L, 77)
    reads = synth.plant_errors(reads, 5, 1, edit, 7 + k)
    sym, off = synth.flatten(reads)
    q = g.upload(sym, off)
    for sch in (schemes.optimum(0, k), schemes.facade_scheme(edit, k, L)[0]):
        part = schemes.uniform_partition(sch[0].shape[1], L)
        for n in (1, 2, 3, 7, 10**9):
            exp = o.search_ng26(sym, off, sch, part, edit, max_hits=n)
            got = g.search_scheme(q, sch, part, edit, n=n).hits()
            assert _same_list(got, exp), (k, edit, n)


@pytest.mark.parametrize("edit", [False, True])
def test_search_n_repetitive_text(gpu, edit):
    """wide intervals: the order of the children of search_next_dir (match, deletion(c) / substitution(c) by symbol, insertion) and
    clipping of cursors with many rows"""
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
    q = g.upload(sym, off)
    for k in (1, 2, 3):
        sch = schemes.optimum(0, k) if k < 3 else schemes.h2(5, 0, 3)
        if not edit:
            sch = schemes.limit_to_hamming(sch)
        part = schemes.uniform_partition(sch[0].shape[1], 30)
        for n in (1, 4, 25, 26, 1000, 10**7):
            exp = o.search_ng26(sym, off, sch, part, edit, max_hits=n)
            got = g.search_scheme(q, sch, part, edit, n=n).hits()
            assert _same_list(got, exp), (k, edit, n)
            assert int(got["len"].astype(np.int64).sum()) <= n * 60
        assert len(o.search_ng26(sym, off, sch, part, edit)) > len(o.search_ng26(sym, off, sch, part, edit, max_hits=4))   # the limit cuts


def test_search_n_protein(gpu):
    """generic layout (sigma = 21): byte-symbol jump tables, 42 + 2 child ranks per node"""
    from fmb200 import schemes, synth
    text = synth.multi_text([5000, 1200], 21, 9)
    o, g = make_index_pair(gpu, text, 21, 4)
    L = 20
    reads, _ = synth.reads_from_text(text[:5000], 200, L, 3)
    reads = synth.plant_errors(reads, 21, 1, True, 4)
    sym, off = synth.flatten(reads)
    q = g.upload(sym, off)
    for k, edit in ((1, False), (1, True), (2, True)):
        sch = schemes.optimum(0, k)
        part = schemes.uniform_partition(sch[0].shape[1], L)
        for n in (1, 2, 5):
            exp = o.search_ng26(sym, off, sch, part, edit, max_hits=n)
            assert _same_list(g.search_scheme(q, sch, part, edit, n=n).hits(), exp), (k, edit, n)


def test_search_n_refuses_keys_that_do_not_fit(pair):
    from fmb200 import schemes, synth
    text, o, g = pair
    L = 40
    reads, _ = synth.reads_from_text(text[:6000], 4, L, 1)
    sym, off = synth.flatten(reads)
    sch = schemes.backtracking(1, 0, 6)              # 6 errors x 11 bits > 56
    with pytest.raises(gpu_error()):
        g.search_scheme(g.upload(sym, off), sch, [L], True, n=5)


def gpu_error():
    import fmb200
    return fmb200.FmbError


def test_search_n_with_spilled_frontier(gpu):
    """a 40-item stack per warp (FMB_SCHEME_CAP, read once per process) makes the frontier spill to the global overflow list, whose
    items carry their keys to the next launch: the repetitive-text cases again, in a fresh process"""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, FMB_SCHEME_CAP="40")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_search_n.py"), "-q", "-x", "-k", "repetitive or random_text"],
                       env=env, capture_output=True, text=True, timeout=900, cwd=root)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("edit", [False, True])
def test_first_hit_limit_bounds_the_work(gpu, edit):
    """n = 1: the kernel keeps the smallest discovery-order key per query and drops subtrees that cannot beat it (and starts the later
    searches of a query after the earlier ones), so the first hit costs fewer extensions than enumerating everything -- with the same
    result as the oracle's search_n"""
    from fmb200 import schemes, synth
    text = synth.multi_text([120000], 5, 29)
    o, g = make_index_pair(gpu, text, 5, 8)
    # (enough roots that the searches of a query do not all start in the first wave: 3 x 200 k against ~150 k resident lanes)
    reads, _ = synth.reads_from_text(text[:120001], 200000, 48, 7)
    reads[150000:] = synth.plant_errors(reads[150000:], 5, 1, edit, 9)
    sym, off = synth.flatten(reads)
    q = g.upload(sym, off)
    sch, part = schemes.optimum(0, 2), schemes.uniform_partition(4, 48)
    first = g.search_scheme(q, sch, part, edit, n=1)
    exp = o.search_ng26(sym, off, sch, part, edit, max_hits=1)
    assert _same_list(first.hits(), exp)
    everything = g.search_scheme(q, sch, part, edit, n=10**9)
    assert first.stats.extensions < 0.9 * everything.stats.extensions, (first.stats.extensions, everything.stats.extensions)
    # small limits n <= 8: the n smallest keys per query are kept, the n-th is the bound; n = 9 bounds the output only
    few = g.search_scheme(q, sch, part, edit, n=3)
    assert _same_list(few.hits(), o.search_ng26(sym, off, sch, part, edit, max_hits=3))
    assert few.stats.extensions <= everything.stats.extensions          # (most reads have fewer than three rows: little to cut)
    nine = g.search_scheme(q, sch, part, edit, n=9)
    assert _same_list(nine.hits(), o.search_ng26(sym, off, sch, part, edit, max_hits=9))
    assert nine.stats.extensions == everything.stats.extensions


def test_small_hit_limits_on_repeats_count_rows_not_hits(gpu):
    """a repetitive text: single hits cover many rows, so one hit can fill the limit (a hit of len rows counts min(len, n) times in the
    per-query bound) -- the clipped lists still equal the oracle's for every small n, Hamming and edit distance"""
    from fmb200 import schemes, synth
    unit = synth.text(600, 5, 3)[:-1]
    rng = np.random.default_rng(12)
    copies = []
    for _ in range(40):
        c = unit.copy()
        for p in rng.integers(0, c.size, 6):
            c[p] = 1 + (c[p] + int(rng.integers(0, 3))) % 4
        copies.append(c)
    text = np.concatenate(copies + [np.zeros(1, dtype=np.uint8)]).astype(np.uint8)
    o, g = make_index_pair(gpu, text, 5, 8)
    reads, _ = synth.reads_from_text(text, 3000, 36, 5)
    sym, off = synth.flatten(reads)
    q = g.upload(sym, off)
    sch, part = schemes.optimum(0, 2), schemes.uniform_partition(4, 36)
    for edit in (False, True):
        for n in (1, 2, 4, 8):
            got = g.search_scheme(q, sch, part, edit, n=n)
            assert _same_list(got.hits(), o.search_ng26(sym, off, sch, part, edit, max_hits=n)), (edit, n)
