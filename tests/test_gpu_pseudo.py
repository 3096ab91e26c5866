"""GPU parity: search_pseudo (search/SearchPseudo.h) -- the plain recursive search over an expanded scheme.  The device runs it on the
part form of the scheme (fmb_search_scheme_pseudo); for edit distance that is the scheme kernel WITHOUT the redundancy filter of
search_ng26, so every alignment and every duplicate the reference reports must appear."""
import numpy as np
import pytest

from helpers import hits_equal, make_index_pair

pytestmark = pytest.mark.gpu


def _remap(s):
    return np.array([{"A": 1, "B": 2, "C": 3, "D": 4}[c] for c in s], dtype=np.uint8)


def test_golden_pseudo(gpu):
    """search/checkSearches.cpp:143-176 (edit) and :215-241 (Hamming): expand(pigeon_opt(0, 1), 2), queries CC / BB"""
    from fmb200 import schemes, synth
    text = np.concatenate([_remap("AAACAAABAAA"), [0], _remap("AAABAAACAAA"), [0]]).astype(np.uint8)
    o, g = make_index_pair(gpu, text, 5, 1)
    q = g.upload(*synth.flatten([_remap("CC"), _remap("BB")]))
    pigeon_opt = (np.array([[0, 1], [1, 0]]), np.array([[0, 0], [0, 1]]), np.array([[0, 1], [0, 1]]))
    part = schemes.uniform_partition(2, 2)

    def located(res):
        return sorted((int(r["qidx"]), int(r["seq"]), int(r["pos"])) for r in g.locate(res).locs())

    exp_edit = [(0, 0, 2), (0, 0, 3), (0, 0, 3), (0, 0, 3), (0, 1, 6), (0, 1, 7), (0, 1, 7), (0, 1, 7),
                (1, 0, 6), (1, 0, 7), (1, 0, 7), (1, 0, 7), (1, 1, 2), (1, 1, 3), (1, 1, 3), (1, 1, 3)]
    ham = [(0, 0, 2), (0, 0, 3), (0, 1, 6), (0, 1, 7), (1, 0, 6), (1, 0, 7), (1, 1, 2), (1, 1, 3)]
    assert located(g.search_scheme_pseudo(q, pigeon_opt, part, True)) == exp_edit
    assert located(g.search_scheme_pseudo(q, pigeon_opt, part, False)) == ham


@pytest.mark.parametrize("kind", ["random", "repeats", "protein"])
def test_pseudo_matches_oracle(gpu, kind):
    from fmb200 import schemes, synth
    rng = np.random.default_rng(17)
    if kind == "random":
        sigma, L, nq = 5, 30, 200
        text = synth.multi_text([5000, 1500, 40], sigma, 31)
        body = text[:5000]
    elif kind == "repeats":
        sigma, L, nq = 5, 18, 60
        unit = rng.integers(1, 5, 45).astype(np.uint8)
        chunks = []
        for _ in range(30):
            u = unit.copy()
            u[rng.integers(0, 45, 2)] = rng.integers(1, 5, 2)
            chunks.append(np.concatenate([u, [0]]))
        text = np.concatenate(chunks).astype(np.uint8)
        body = None
    else:
        sigma, L, nq = 21, 16, 150
        text = synth.multi_text([4000, 900], sigma, 33)
        body = text[:4000]
    o, g = make_index_pair(gpu, text, sigma, 4)
    if body is not None:
        reads, _ = synth.reads_from_text(body, nq, L, 3)
    else:
        reads = np.array([text[46 * (i % 30) + 3: 46 * (i % 30) + 3 + L] for i in range(nq)], dtype=np.uint8)
    reads = synth.plant_errors(reads, sigma, 1, True, 4)
    sym, off = synth.flatten(reads)
    q = g.upload(sym, off)
    total = 0
    for k in (1, 2):
        for sch in (schemes.optimum(0, k), schemes.h2(k + 2, 0, k), schemes.backtracking(k + 1, 0, k)):
            part = schemes.uniform_partition(sch[0].shape[1], L)
            expanded = schemes.expand(sch, part)
            exp = o.search_pseudo(sym, off, expanded, True)
            got = g.search_scheme_pseudo(q, sch, part, True).hits()
            assert hits_equal(got, exp), (kind, k)
            total += len(exp)
            # more alignments than the filtered search of search_ng26 reports, and a superset of its cursors
            filtered = o.search_ng26(sym, off, sch, part, True)
            assert len(exp) >= len(filtered)
    assert total > 1000
