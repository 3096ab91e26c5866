"""shared helpers of the parity tests (oracle = checker, libfmb200 = the thing under test)"""
import numpy as np

from oracle.pyoracle import Oracle, sort_hits, sort_locs  # noqa: F401


def make_index_pair(fmb, text, sigma, rate, bidirectional=True, device=0):
    """oracle index built on the CPU from `text`; device index built from exactly the same BWT bytes and samples
    through fmb_index_create (the BiFMIndex(bwt, bwtRev, SparseArray) seam, fmindex/BiFMIndex.h:40-51)."""
    o = Oracle.build(text, sigma, rate, bidirectional)
    bm, sq, sp = o.samples
    g = fmb.Index.from_bwt(sigma, o.bwt, o.bwt_rev if bidirectional else None, bm, sq, sp, device=device)
    return o, g


def hits_equal(a, b):
    a, b = sort_hits(np.asarray(a)), sort_hits(np.asarray(b))
    return a.shape == b.shape and np.array_equal(a, b)


def locs_equal(a, b):
    a, b = sort_locs(np.asarray(a)), sort_locs(np.asarray(b))
    return a.shape == b.shape and np.array_equal(a, b)


def strip_lb_rev(h):
    h = h.copy()
    h["lb_rev"] = 0
    return h
