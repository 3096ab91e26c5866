"""GPU tests of the limits the k-error kernels state loudly (alphabet size, error bounds) and of the measurement aid
fmb_measure_gather (in-run random-request ceiling)."""
import numpy as np
import pytest

from helpers import hits_equal, make_index_pair

pytestmark = pytest.mark.gpu

FMB_EUNSUPPORTED = -5


def _reads(text, sigma, n, L, seed, errors, edit):
    from fmb200 import synth
    reads, _ = synth.reads_from_text(text, n, L, seed)
    reads = synth.plant_errors(reads, sigma, errors, edit, seed + 1)
    return synth.flatten(reads)


@pytest.mark.parametrize("sigma", [28])
def test_scheme_and_backtracking_at_the_largest_supported_alphabet(gpu, sigma):
    """the child mask of the scheme kernel keeps deletions at bits 8.. and substitutions at bits 36..: sigma = 28 is the largest
    alphabet whose symbols fit both ranges -- every symbol, the largest included, must produce its children"""
    from fmb200 import schemes, synth
    text = synth.multi_text([6000, 500], sigma, 23)
    o, g = make_index_pair(gpu, text, sigma, 8)
    sym, off = _reads(text[:6000], sigma, 300, 24, 5, 2, True)
    q = g.upload(sym, off)
    for k, edit in ((1, False), (1, True), (2, True)):
        sch = schemes.optimum(0, k)
        part = schemes.uniform_partition(sch[0].shape[1], 24)
        assert hits_equal(g.search_scheme(q, sch, part, edit).hits(), o.search_ng26(sym, off, sch, part, edit))
    sym, off = _reads(text[:6000], sigma, 100, 12, 9, 1, False)
    assert hits_equal(g.search_backtracking(g.upload(sym, off), 1).hits(), o.search_backtracking(sym, off, 1))


@pytest.mark.parametrize("sigma", [29, 32])
def test_larger_alphabets_are_refused_for_k_error_search_but_serve_exact_search(gpu, sigma):
    from fmb200 import schemes, synth
    text = synth.multi_text([3000, 200], sigma, 31)
    o, g = make_index_pair(gpu, text, sigma, 8)
    sym, off = _reads(text[:3000], sigma, 50, 20, 3, 0, False)
    q = g.upload(sym, off)
    assert hits_equal(g.search_exact(q).hits(), o.search_exact(sym, off))
    sch = schemes.optimum(0, 1)
    with pytest.raises(gpu.FmbError) as e:
        g.search_scheme(q, sch, schemes.uniform_partition(2, 20), True)
    assert e.value.code == FMB_EUNSUPPORTED and "sigma" in str(e.value)
    with pytest.raises(gpu.FmbError) as e:
        g.search_backtracking(q, 1)
    assert e.value.code == FMB_EUNSUPPORTED


def test_error_bounds_above_fifteen_are_refused(gpu):
    from fmb200 import synth
    text = synth.multi_text([2000], 5, 2)
    o, g = make_index_pair(gpu, text, 5, 8)
    sym, off = _reads(text[:2000], 5, 10, 40, 3, 0, False)
    q = g.upload(sym, off)
    pi = np.array([[0, 1]], dtype=np.uint32)
    l = np.array([[0, 0]], dtype=np.uint32)
    u = np.array([[0, 16]], dtype=np.uint32)
    with pytest.raises(gpu.FmbError) as e:
        g.search_scheme(q, (pi, l, u), [20, 20], True)
    assert e.value.code == FMB_EUNSUPPORTED
    with pytest.raises(gpu.FmbError) as e:
        g.search_backtracking(q, 16)
    assert e.value.code == FMB_EUNSUPPORTED


def test_measure_gather_reports_a_plausible_request_rate(gpu):
    from fmb200 import capi, synth
    text = synth.multi_text([400000], 5, 7)
    g = gpu.Index.build(5, text, sampling_rate=16, device=0)
    for table, req_bytes in ((0, 128), (1, 32)):
        rps, tb, rb = capi.measure_gather(g, table, 1 << 22)
        assert rb == req_bytes and tb >= text.size // 4 and 1e8 < rps < 1e12
    rps, tb, rb = capi.measure_gather(g, 2, 1 << 22)
    assert rb in (8, 16) and tb == text.size * rb


def test_image_budget_drops_tables_but_not_results(gpu):
    """fmb_set_image_budget: with a budget the optional tables are admitted in priority order; results never depend on them"""
    from fmb200 import capi, schemes, synth
    text = synth.multi_text([300000, 20000], 5, 11)
    sym, off = _reads(text[:300000], 5, 2000, 60, 3, 1, True)
    sch, part = schemes.optimum(0, 1), schemes.uniform_partition(2, 60)
    full = gpu.Index.build(5, text, sampling_rate=8, device=0)
    q = full.upload(sym, off)
    exact, edit = full.search_exact(q).hits(), full.search_scheme(q, sch, part, True).hits()
    locs = full.locate(full.search_exact(q)).locs()
    n = text.size
    try:
        seen = set()
        for budget in (2 * n, 5 * n, 14 * n, 24 * n, 40 * n, 80 * n):
            capi.set_image_budget(budget)
            g = gpu.Index.build(5, text, sampling_rate=8, device=0)
            i = g.info
            assert i.device_bytes <= budget * 1.02 + (1 << 20) or i.tables == 0
            seen.add(i.tables)
            qq = g.upload(sym, off)
            assert hits_equal(g.search_exact(qq).hits(), exact)
            assert hits_equal(g.search_scheme(qq, sch, part, True).hits(), edit)
            assert np.array_equal(np.sort(g.locate(g.search_exact(qq)).locs(), order=["qidx", "seq", "pos", "e"]), np.sort(locs, order=["qidx", "seq", "pos", "e"]))
        assert len(seen) >= 4 and full.info.tables in seen
    finally:
        capi.set_image_budget(0)
