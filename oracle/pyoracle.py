"""ctypes wrappers of the oracle libraries.  TEST INFRASTRUCTURE ONLY: imported by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs -- never by the product package.

  Oracle  -> oracle/libfmoracle.so      (plain-C port, fm_oracle.c; always buildable: gcc only)
  Ref     -> oracle/_ref/libfmref.so    (the reference's own headers, ref_shim.cpp; prebuilt where
                                         /root/reference exists, travels to the GPU box as a binary)
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_LIB = os.path.join(HERE, "libfmoracle.so")
REF_LIB = os.path.join(HERE, "_ref", "libfmref.so")

HIT_DTYPE = np.dtype([("qidx", "<u8"), ("lb", "<u8"), ("lb_rev", "<u8"), ("len", "<u8"), ("steps", "<u8"), ("e", "<u8")])
LOC_DTYPE = np.dtype([("qidx", "<u8"), ("seq", "<u8"), ("pos", "<u8"), ("e", "<u8")])
UINT64_MAX = 0xFFFFFFFFFFFFFFFF


class Counters(C.Structure):
    _fields_ = [("extensions", C.c_uint64), ("occ_lookups", C.c_uint64), ("lf_steps", C.c_uint64), ("locate_lookups", C.c_uint64)]


def build_oracle():
    src = os.path.join(HERE, "fm_oracle.c")
    if not os.path.exists(ORACLE_LIB) or os.path.getmtime(ORACLE_LIB) < os.path.getmtime(src):
        tmp = ORACLE_LIB + ".tmp.%d" % os.getpid()          # linked under a temporary name: concurrent processes never map a partial file
        subprocess.run(["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-o", tmp, src], check=True, cwd=HERE)
        os.replace(tmp, ORACLE_LIB)
    return ORACLE_LIB


def build_ref():
    """(re)build oracle/_ref/libfmref.so when the reference sources are present; returns the path or None"""
    if os.path.isdir("/root/reference/src/fmindex-collection"):
        shim = os.path.join(HERE, "ref_shim.cpp")
        if not os.path.exists(REF_LIB) or os.path.getmtime(REF_LIB) < os.path.getmtime(shim):
            subprocess.run(["bash", os.path.join(HERE, "build_ref.sh")], check=True)
    return REF_LIB if os.path.exists(REF_LIB) else None


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def _take(lib_free, ptr, count, dtype):
    if count == 0:
        if ptr:
            lib_free(ptr)
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (count * dtype.itemsize)).from_address(ptr)
    out = np.frombuffer(buf, dtype=dtype, count=count).copy()
    lib_free(ptr)
    return out


def _scheme_args(scheme, partition):
    pi, l, u = (np.ascontiguousarray(a, dtype=np.uint32) for a in scheme)
    part = _u32(partition)
    return pi, l, u, part


class Oracle:
    """C port of the reference search path (fm_oracle.h)."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(build_oracle())
            L.fmo_index_build.restype = C.c_void_p
            L.fmo_index_from_bwt.restype = C.c_void_p
            for f in ("fmo_size", "fmo_n_samples", "fmo_symbol", "fmo_rank", "fmo_prefix_rank", "fmo_search_exact",
                      "fmo_search_ng26", "fmo_search_backtracking", "fmo_locate"):
                getattr(L, f).restype = C.c_uint64
            for f in ("fmo_bwt", "fmo_bwt_rev", "fmo_sa", "fmo_C", "fmo_sample_bitmap", "fmo_sample_seq", "fmo_sample_pos"):
                getattr(L, f).restype = C.c_void_p
            L.fmo_free.argtypes = [C.c_void_p]
            L.fmo_index_free.argtypes = [C.c_void_p]
            cls._lib = L
        return cls._lib

    def __init__(self, handle, sigma, bidirectional):
        self.h = C.c_void_p(handle)
        self.sigma = sigma
        self.bidirectional = bidirectional

    @classmethod
    def build(cls, text, sigma, rate, bidirectional=True):
        text = _u8(text)
        h = cls.lib().fmo_index_build(_p(text), C.c_uint64(text.size), C.c_uint32(sigma), C.c_uint32(rate), C.c_int(int(bidirectional)))
        return cls(h, sigma, bidirectional)

    @classmethod
    def from_bwt(cls, sigma, bwt, bwt_rev, bitmap, seq, pos):
        bwt = _u8(bwt)
        rev = None if bwt_rev is None else _u8(bwt_rev)
        bitmap, seq, pos = _u64(bitmap), _u32(seq), _u32(pos)
        h = cls.lib().fmo_index_from_bwt(C.c_uint32(sigma), C.c_uint64(bwt.size), _p(bwt), _p(rev), _p(bitmap), _p(seq), _p(pos), C.c_uint64(seq.size))
        return cls(h, sigma, rev is not None)

    def __del__(self):
        try:
            if self.h:
                self.lib().fmo_index_free(self.h)
                self.h = None
        except Exception:
            pass

    @property
    def n(self):
        return self.lib().fmo_size(self.h)

    def _arr(self, fn, count, dtype):
        ptr = getattr(self.lib(), fn)(self.h)
        if not ptr or count == 0:
            return None if not ptr else np.zeros(0, dtype=dtype)
        buf = (C.c_char * (count * np.dtype(dtype).itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype, count=count).copy()

    @property
    def bwt(self):
        return self._arr("fmo_bwt", self.n, np.uint8)

    @property
    def bwt_rev(self):
        return self._arr("fmo_bwt_rev", self.n, np.uint8) if self.bidirectional else None

    @property
    def sa(self):
        return self._arr("fmo_sa", self.n, np.uint64)

    @property
    def C(self):
        return self._arr("fmo_C", self.sigma + 1, np.uint64)

    @property
    def samples(self):
        ns = self.lib().fmo_n_samples(self.h)
        return (self._arr("fmo_sample_bitmap", (self.n + 63) // 64, np.uint64), self._arr("fmo_sample_seq", ns, np.uint32),
                self._arr("fmo_sample_pos", ns, np.uint32))

    def symbol(self, idx, dir=0):
        return self.lib().fmo_symbol(self.h, C.c_int(dir), C.c_uint64(int(idx)))

    def rank(self, idx, symb, dir=0):
        return self.lib().fmo_rank(self.h, C.c_int(dir), C.c_uint64(int(idx)), C.c_uint64(int(symb)))

    def prefix_rank(self, idx, symb, dir=0):
        return self.lib().fmo_prefix_rank(self.h, C.c_int(dir), C.c_uint64(int(idx)), C.c_uint64(int(symb)))

    def all_ranks_and_prefix_ranks(self, idx, dir=0):
        rs = np.zeros(self.sigma, dtype=np.uint64)
        prs = np.zeros(self.sigma, dtype=np.uint64)
        self.lib().fmo_all_ranks_and_prefix_ranks(self.h, C.c_int(dir), C.c_uint64(int(idx)), _p(rs), _p(prs))
        return rs, prs

    def extend(self, cur, symb, right):
        cur = _u64(cur)
        out = np.zeros(4, dtype=np.uint64)
        (self.lib().fmo_extend_right if right else self.lib().fmo_extend_left)(self.h, _p(cur), C.c_uint64(int(symb)), _p(out))
        return out

    def extend_all(self, cur, right):
        cur = _u64(cur)
        out = np.zeros((self.sigma, 4), dtype=np.uint64)
        (self.lib().fmo_extend_right_all if right else self.lib().fmo_extend_left_all)(self.h, _p(cur), _p(out))
        return out

    def search_exact(self, symbols, offsets, counters=None):
        symbols, offsets = _u8(symbols), _u64(offsets)
        out = C.c_void_p()
        n = self.lib().fmo_search_exact(self.h, _p(symbols), _p(offsets), C.c_uint64(offsets.size - 1), C.byref(out),
                                        C.byref(counters) if counters is not None else None)
        return _take(self.lib().fmo_free, out.value, n, HIT_DTYPE)

    def search_ng26(self, symbols, offsets, scheme, partition, edit, max_hits=UINT64_MAX, counters=None):
        symbols, offsets = _u8(symbols), _u64(offsets)
        pi, l, u, part = _scheme_args(scheme, partition)
        out = C.c_void_p()
        n = self.lib().fmo_search_ng26(self.h, _p(symbols), _p(offsets), C.c_uint64(offsets.size - 1), C.c_int(int(edit)),
                                       C.c_uint32(pi.shape[0]), C.c_uint32(pi.shape[1]), _p(pi), _p(l), _p(u), _p(part),
                                       C.c_uint64(max_hits), C.byref(out), C.byref(counters) if counters is not None else None)
        return _take(self.lib().fmo_free, out.value, n, HIT_DTYPE)

    def search_ng26_keys(self, symbols, offsets, scheme, partition, edit, max_hits=UINT64_MAX):
        """hits in the order the depth-first search reports them + their discovery-order keys (fm_oracle.c ng26_key_edge)"""
        symbols, offsets = _u8(symbols), _u64(offsets)
        pi, l, u, part = _scheme_args(scheme, partition)
        out, keys = C.c_void_p(), C.c_void_p()
        f = self.lib().fmo_search_ng26_keys
        f.restype = C.c_uint64
        n = f(self.h, _p(symbols), _p(offsets), C.c_uint64(offsets.size - 1), C.c_int(int(edit)),
              C.c_uint32(pi.shape[0]), C.c_uint32(pi.shape[1]), _p(pi), _p(l), _p(u), _p(part),
              C.c_uint64(max_hits), C.byref(out), C.byref(keys), None)
        return _take(self.lib().fmo_free, out.value, n, HIT_DTYPE), _take(self.lib().fmo_free, keys.value, n, np.dtype(np.uint64))

    def search_pseudo(self, symbols, offsets, expanded, edit):
        """search_pseudo::search<Edit> with an expanded scheme = (pi, l, u) arrays of shape n_searches x L"""
        symbols, offsets = _u8(symbols), _u64(offsets)
        pi, l, u = (np.ascontiguousarray(a, dtype=np.uint32) for a in expanded)
        out = C.c_void_p()
        f = self.lib().fmo_search_pseudo
        f.restype = C.c_uint64
        n = f(self.h, _p(symbols), _p(offsets), C.c_uint64(offsets.size - 1), C.c_int(int(edit)), C.c_uint32(pi.shape[0]), C.c_uint32(pi.shape[1]),
              _p(pi), _p(l), _p(u), C.byref(out), None)
        return _take(self.lib().fmo_free, out.value, n, HIT_DTYPE)

    def search_backtracking(self, symbols, offsets, max_errors, counters=None):
        symbols, offsets = _u8(symbols), _u64(offsets)
        out = C.c_void_p()
        n = self.lib().fmo_search_backtracking(self.h, _p(symbols), _p(offsets), C.c_uint64(offsets.size - 1), C.c_uint32(max_errors),
                                               C.byref(out), C.byref(counters) if counters is not None else None)
        return _take(self.lib().fmo_free, out.value, n, HIT_DTYPE)

    def locate(self, hits, counters=None):
        hits = np.ascontiguousarray(hits, dtype=HIT_DTYPE)
        out = C.c_void_p()
        n = self.lib().fmo_locate(self.h, _p(hits), C.c_uint64(hits.size), C.byref(out), C.byref(counters) if counters is not None else None)
        return _take(self.lib().fmo_free, out.value, n, LOC_DTYPE)

    def locate_row(self, row):
        out = np.zeros(3, dtype=np.uint64)
        self.lib().fmo_locate_row(self.h, C.c_uint64(int(row)), _p(out))
        return tuple(int(x) for x in out)

    def single_locate_step(self, row):
        out = np.zeros(2, dtype=np.uint64)
        ok = self.lib().fmo_single_locate_step(self.h, C.c_uint64(int(row)), _p(out))
        return (int(out[0]), int(out[1])) if ok else None


class Ref:
    """The reference's own implementation (oracle/_ref/libfmref.so).  Ref.available() is False on a box where
    the library was not prebuilt."""

    _lib = None

    @classmethod
    def available(cls):
        return os.path.exists(REF_LIB)

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(REF_LIB)
            L.fmr_index_from_bwt.restype = C.c_void_p
            L.fmr_index_build.restype = C.c_void_p
            for f in ("fmr_size", "fmr_C", "fmr_symbol", "fmr_rank", "fmr_prefix_rank", "fmr_search_exact", "fmr_search_ng26",
                      "fmr_search_facade", "fmr_search_backtracking", "fmr_locate"):
                getattr(L, f).restype = C.c_uint64
            L.fmr_scheme_generate.restype = C.c_uint32
            L.fmr_free.argtypes = [C.c_void_p]
            L.fmr_index_free.argtypes = [C.c_void_p]
            cls._lib = L
        return cls._lib

    def __init__(self, handle, sigma, bidirectional):
        if not handle:
            raise RuntimeError("reference index construction failed (sigma must be 5 or 21)")
        self.h = C.c_void_p(handle)
        self.sigma = sigma
        self.bidirectional = bidirectional
        self.last_seconds = 0.0

    @classmethod
    def build(cls, text, sigma, rate, bidirectional=True):
        text = _u8(text)
        return cls(cls.lib().fmr_index_build(_p(text), C.c_uint64(text.size), C.c_uint32(sigma), C.c_uint32(rate), C.c_int(int(bidirectional))),
                   sigma, bidirectional)

    @classmethod
    def from_bwt(cls, sigma, bwt, bwt_rev, bitmap, seq, pos):
        bwt = _u8(bwt)
        rev = None if bwt_rev is None else _u8(bwt_rev)
        bitmap, seq, pos = _u64(bitmap), _u32(seq), _u32(pos)
        return cls(cls.lib().fmr_index_from_bwt(C.c_uint32(sigma), C.c_uint64(bwt.size), _p(bwt), _p(rev), _p(bitmap), _p(seq), _p(pos),
                                                C.c_uint64(seq.size)), sigma, rev is not None)

    def __del__(self):
        try:
            if self.h:
                self.lib().fmr_index_free(self.h)
                self.h = None
        except Exception:
            pass

    @property
    def n(self):
        return self.lib().fmr_size(self.h)

    @property
    def C(self):
        return np.array([self.lib().fmr_C(self.h, C.c_uint32(s)) for s in range(self.sigma + 1)], dtype=np.uint64)

    def symbol(self, idx, dir=0):
        return self.lib().fmr_symbol(self.h, C.c_int(dir), C.c_uint64(int(idx)))

    def rank(self, idx, symb, dir=0):
        return self.lib().fmr_rank(self.h, C.c_int(dir), C.c_uint64(int(idx)), C.c_uint64(int(symb)))

    def prefix_rank(self, idx, symb, dir=0):
        return self.lib().fmr_prefix_rank(self.h, C.c_int(dir), C.c_uint64(int(idx)), C.c_uint64(int(symb)))

    def all_ranks_and_prefix_ranks(self, idx, dir=0):
        rs = np.zeros(self.sigma, dtype=np.uint64)
        prs = np.zeros(self.sigma, dtype=np.uint64)
        self.lib().fmr_all_ranks_and_prefix_ranks(self.h, C.c_int(dir), C.c_uint64(int(idx)), _p(rs), _p(prs))
        return rs, prs

    def extend(self, cur, symb, right):
        cur = _u64(cur)
        out = np.zeros(4, dtype=np.uint64)
        self.lib().fmr_extend(self.h, C.c_int(int(right)), _p(cur), C.c_uint64(int(symb)), _p(out))
        return out

    def locate_row(self, row):
        out = np.zeros(3, dtype=np.uint64)
        self.lib().fmr_locate_row(self.h, C.c_uint64(int(row)), _p(out))
        return tuple(int(x) for x in out)

    def _secs(self):
        self._s = C.c_double(0)
        return C.byref(self._s)

    def search_exact(self, symbols, offsets, threads=1):
        symbols, offsets = _u8(symbols), _u64(offsets)
        out = C.c_void_p()
        n = self.lib().fmr_search_exact(self.h, _p(symbols), _p(offsets), C.c_uint64(offsets.size - 1), C.byref(out), C.c_int(threads), self._secs())
        self.last_seconds = self._s.value
        return _take(self.lib().fmr_free, out.value, n, HIT_DTYPE)

    def search_ng26(self, symbols, offsets, scheme, partition, edit, max_hits=UINT64_MAX, threads=1):
        symbols, offsets = _u8(symbols), _u64(offsets)
        pi, l, u, part = _scheme_args(scheme, partition)
        out = C.c_void_p()
        n = self.lib().fmr_search_ng26(self.h, _p(symbols), _p(offsets), C.c_uint64(offsets.size - 1), C.c_int(int(edit)),
                                       C.c_uint32(pi.shape[0]), C.c_uint32(pi.shape[1]), _p(pi), _p(l), _p(u), _p(part),
                                       C.c_uint64(max_hits), C.byref(out), C.c_int(threads), self._secs())
        self.last_seconds = self._s.value
        return _take(self.lib().fmr_free, out.value, n, HIT_DTYPE)

    def search_facade(self, symbols, offsets, edit, errors, threads=1):
        symbols, offsets = _u8(symbols), _u64(offsets)
        out = C.c_void_p()
        n = self.lib().fmr_search_facade(self.h, _p(symbols), _p(offsets), C.c_uint64(offsets.size - 1), C.c_int(int(edit)),
                                         C.c_uint32(errors), C.byref(out), C.c_int(threads), self._secs())
        self.last_seconds = self._s.value
        return _take(self.lib().fmr_free, out.value, n, HIT_DTYPE)

    def search_pseudo(self, symbols, offsets, expanded, edit, threads=1):
        symbols, offsets = _u8(symbols), _u64(offsets)
        pi, l, u = (np.ascontiguousarray(a, dtype=np.uint32) for a in expanded)
        out = C.c_void_p()
        f = self.lib().fmr_search_pseudo
        f.restype = C.c_uint64
        n = f(self.h, _p(symbols), _p(offsets), C.c_uint64(offsets.size - 1), C.c_int(int(edit)), C.c_uint32(pi.shape[0]), C.c_uint32(pi.shape[1]),
              _p(pi), _p(l), _p(u), C.byref(out), C.c_int(threads), self._secs())
        return _take(self.lib().fmr_free, out.value, n, HIT_DTYPE)

    def search_backtracking(self, symbols, offsets, max_errors, threads=1):
        symbols, offsets = _u8(symbols), _u64(offsets)
        out = C.c_void_p()
        n = self.lib().fmr_search_backtracking(self.h, _p(symbols), _p(offsets), C.c_uint64(offsets.size - 1), C.c_uint32(max_errors),
                                               C.byref(out), C.c_int(threads), self._secs())
        self.last_seconds = self._s.value
        return _take(self.lib().fmr_free, out.value, n, HIT_DTYPE)

    def locate(self, hits, threads=1):
        hits = np.ascontiguousarray(hits, dtype=HIT_DTYPE)
        out = C.c_void_p()
        n = self.lib().fmr_locate(self.h, _p(hits), C.c_uint64(hits.size), C.byref(out), C.c_int(threads), self._secs())
        self.last_seconds = self._s.value
        return _take(self.lib().fmr_free, out.value, n, LOC_DTYPE)

    @classmethod
    def scheme(cls, name, min_k, max_k):
        cap = 4096
        pi = np.zeros(cap, dtype=np.uint32)
        l = np.zeros(cap, dtype=np.uint32)
        u = np.zeros(cap, dtype=np.uint32)
        npart = C.c_uint32(0)
        ns = cls.lib().fmr_scheme_generate(name.encode(), C.c_uint32(min_k), C.c_uint32(max_k), C.c_uint32(cap), C.byref(npart), _p(pi), _p(l), _p(u))
        if ns == 0:
            raise ValueError(f"reference generator {name}({min_k},{max_k}) unavailable")
        k = ns * npart.value
        shape = (ns, npart.value)
        return pi[:k].reshape(shape).copy(), l[:k].reshape(shape).copy(), u[:k].reshape(shape).copy()


def sort_hits(h):
    """canonical multiset order of hit records"""
    return np.sort(h, order=["qidx", "lb", "lb_rev", "len", "steps", "e"])


def sort_locs(l):
    return np.sort(l, order=["qidx", "seq", "pos", "e"])
