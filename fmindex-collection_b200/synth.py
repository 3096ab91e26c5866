"""Synthetic workloads.  Everything derives from splitmix64 so that the host (numpy) and the
device (synth_text_kernel in csrc/fmb_kernels.cuh) produce identical bytes."""
import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x):
    """vectorised splitmix64 finaliser over uint64 arrays (wrap-around arithmetic)"""
    with np.errstate(over="ignore"):
        z = (x.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def text(n, sigma, seed, start=0, count=None):
    """T[i] = 1 + (splitmix64(seed + i) >> 32) % (sigma - 1), T[n-1] = 0; returns T[start:start+count]."""
    if count is None:
        count = n - start
    i = np.arange(start, start + count, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = splitmix64(i + np.uint64(seed))
    t = (1 + ((z >> np.uint64(32)) % np.uint64(sigma - 1))).astype(np.uint8)
    if start + count == n and count > 0:
        t[-1] = 0
    return t


def multi_text(lengths, sigma, seed):
    """several sequences, each followed by delimiter 0 (utils.h:413-464 createSequences)"""
    parts = []
    for k, L in enumerate(lengths):
        s = text(L + 1, sigma, seed + 1000003 * (k + 1))
        parts.append(s)  # last symbol already 0
    return np.concatenate(parts) if parts else np.zeros(0, dtype=np.uint8)


def sample_offsets(n, nq, length, seed):
    """uniform offsets in [0, n - 1 - length] (the trailing delimiter is never covered)"""
    i = np.arange(nq, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = splitmix64(i * np.uint64(0x632BE59BD9B4E019) + np.uint64(seed))
    return (z % np.uint64(n - 1 - length + 1)).astype(np.int64)


def reads_from_text(t, nq, length, seed):
    """nq reads of `length` copied from random offsets of the host text t (single sequence ending in 0)"""
    off = sample_offsets(t.size, nq, length, seed)
    idx = off[:, None] + np.arange(length, dtype=np.int64)[None, :]
    return t[idx], off


def plant_errors(reads, sigma, n_err, edit, seed):
    """plant exactly n_err errors per read: substitutions only (edit=False) or a mix of substitutions,
    insertions and deletions that keeps the read length (edit=True), in the spirit of
    search/benchmark_bifmindex_searches.cpp:38-82 of the reference's test suite."""
    rng = np.random.default_rng(seed)
    out = reads.copy()
    nq, L = out.shape
    for q in range(nq):
        r = list(out[q])
        for _ in range(n_err):
            kind = rng.integers(0, 3) if edit else 0
            p = int(rng.integers(1, L - 1))
            if kind == 0:
                r[p] = 1 + (r[p] - 1 + int(rng.integers(1, sigma - 1))) % (sigma - 1)
            elif kind == 1:   # insertion into the read (drop last to keep the length)
                r.insert(p, int(rng.integers(1, sigma)))
                r.pop()
            else:             # deletion from the read (append a random symbol)
                r.pop(p)
                r.append(int(rng.integers(1, sigma)))
        out[q] = r
    return out


def flatten(reads):
    """(nq, L) array or list of 1-D arrays -> (symbols, offsets[nq+1])"""
    if isinstance(reads, np.ndarray) and reads.ndim == 2:
        nq, L = reads.shape
        return np.ascontiguousarray(reads, dtype=np.uint8).reshape(-1), np.arange(nq + 1, dtype=np.uint64) * np.uint64(L)
    lens = np.array([len(r) for r in reads], dtype=np.uint64)
    off = np.zeros(len(reads) + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens)
    sym = np.concatenate([np.asarray(r, dtype=np.uint8) for r in reads]) if len(reads) else np.zeros(0, dtype=np.uint8)
    return sym, off
