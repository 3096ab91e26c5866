"""Search-scheme inputs of the k-error search (host side, tiny).

A scheme is a tuple (pi, l, u) of integer arrays of shape (n_searches, n_parts): the flattened form of the
reference's `search_scheme::Scheme = std::vector<Search{pi, l, u}>` (search_scheme/Search.h:19-28, Scheme.h:13),
pi zero based.  The generators restate search_scheme/generator/{optimum,backtracking,h2}.h and the helpers of
search_scheme/expand.h; tests pin them against tables produced by the reference's own generators
(tests/golden/schemes.json).
"""
import numpy as np

# search_scheme/generator/optimum.h:11-75 (Kianfar et al. optimal schemes), keyed by (minK, K)
_OPTIMUM = {
    (0, 0): [([0], [0], [0])],
    (0, 1): [([0, 1], [0, 0], [0, 1]), ([1, 0], [0, 1], [0, 1])],
    (1, 1): [([0, 1], [0, 1], [0, 1]), ([1, 0], [0, 1], [0, 1])],
    (0, 2): [([0, 1, 2, 3], [0, 0, 1, 1], [0, 0, 2, 2]), ([2, 1, 0, 3], [0, 0, 0, 0], [0, 1, 1, 2]),
             ([3, 2, 1, 0], [0, 0, 0, 2], [0, 1, 2, 2])],
    (1, 2): [([0, 1, 2, 3], [0, 0, 0, 1], [0, 0, 2, 2]), ([2, 1, 0, 3], [0, 0, 1, 1], [0, 1, 1, 2]),
             ([3, 2, 1, 0], [0, 0, 0, 2], [0, 1, 2, 2])],
}


def _pack(searches):
    pi = np.array([s[0] for s in searches], dtype=np.uint32)
    l = np.array([s[1] for s in searches], dtype=np.uint32)
    u = np.array([s[2] for s in searches], dtype=np.uint32)
    return pi, l, u


def optimum(min_k, k):
    """search_scheme/generator/optimum.h:11 (only the table entries needed by BASELINE configs + neighbours)."""
    if (min_k, k) not in _OPTIMUM:
        raise ValueError(f"optimum({min_k},{k}) not tabulated")
    return _pack(_OPTIMUM[(min_k, k)])


def backtracking(n_parts, min_k, k):
    """search_scheme/generator/backtracking.h:15-22"""
    pi = list(range(n_parts))
    l = [0] * n_parts
    u = [k] * n_parts
    l[-1] = min_k
    return _pack([(pi, l, u)])


def uniform_partition(parts, total):
    """createUniformPartition, search_scheme/expand.h:324-336"""
    assert parts > 0 and total >= parts
    base, rest = divmod(total, parts)
    return np.array([base + (1 if i < rest else 0) for i in range(parts)], dtype=np.uint32)


def expand(scheme, counts):
    """expand(scheme, counts), search_scheme/expand.h:67-180: one pi / l / u entry per query symbol (the input of search_pseudo).
    A part is walked forwards when the next part lies to its right (the first part: like the second one); every symbol of a part
    carries the part's upper bound, the last one its lower bound, the others the lower bound of the part before.  (The reference
    drops searches that are not valid afterwards; the generators used here never produce such.)"""
    pi, l, u = scheme
    counts = [int(c) for c in counts]
    starts = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(int)
    out_pi, out_l, out_u = [], [], []
    for s in range(pi.shape[0]):
        rp, rl, ru = [], [], []
        P = pi.shape[1]
        for i in range(P):
            forward = (P == 1 or pi[s, 1] > pi[s, 0]) if i == 0 else pi[s, i] > pi[s, i - 1]
            b, c = starts[pi[s, i]], counts[pi[s, i]]
            rp += [b + j if forward else b + c - 1 - j for j in range(c)]
            ru += [int(u[s, i])] * c
            rl += [int(l[s, i - 1]) if i > 0 else 0] * (c - 1) + ([int(l[s, i])] if c > 0 else [])
        out_pi.append(rp); out_l.append(rl); out_u.append(ru)
    return (np.array(out_pi, dtype=np.uint32), np.array(out_l, dtype=np.uint32), np.array(out_u, dtype=np.uint32))


def limit_to_hamming(scheme):
    """limitToHamming, search_scheme/expand.h:301-319"""
    pi, l, u = (a.copy() for a in scheme)
    for s in range(pi.shape[0]):
        n = pi.shape[1]
        for i in range(n - 1, 0, -1):
            if l[s, i] == 0:
                break
            l[s, i - 1] = max(l[s, i - 1], l[s, i] - 1)
        for i in range(1, n):
            u[s, i] = min(u[s, i], u[s, i - 1] + 1)
    return pi, l, u


# ---- h2 (search_scheme/generator/h2.h) -------------------------------------------------------------------
def _h2_pi(row, n, N, K, mod):
    row = K - row
    shift = mod * row
    n = n + shift
    if n < N - row:
        return n + row
    return N + shift - n - 1


def _h2_diff_matrix(N, K):
    d = [[0] * N for _ in range(K + 1)]
    for i in range(K, N):
        for row in range(K + 1):
            d[row][i] = K - row
    for i in range(K):
        for row in range(K):
            d[row][i] = (row - i + K) % K
        d[K][i] = K
    return d


def _h2_optimized_diff_matrix(N, K):
    mat = _h2_diff_matrix(N, K)

    def is_valid(row, n, v):
        if row == n:
            return False
        if row > n:
            return all(mat[row][i] >= v for i in range(n))
        return all(mat[row][i] <= v for i in range(row + 1, n))

    for i in range(N):
        for j in range(K + 1):
            if i == j or mat[j][i] == 0:
                continue
            if not is_valid(j, i, mat[j][i]):
                index = None
                for k in range(j + 1, K + 1):
                    if is_valid(j, i, mat[k][i]) and is_valid(k, i, mat[j][i]):
                        index = k
                        break
                assert index is not None
                mat[index][i], mat[j][i] = mat[j][i], mat[index][i]
    return mat


def h2(N, min_k, K):
    """search_scheme/generator/h2.h:128-149: K+1 searches over N parts."""
    assert N > 0 and min_k <= K <= N
    pieces = [[_h2_pi(row, i, N, K, 0) for i in range(N)] for row in range(K + 1)]
    lower = [[0] * N for _ in range(K + 1)]
    for i in range(K + 1):
        for j in range(K - i + 1):
            lower[i][N - j - 1] = i
    diffs = _h2_optimized_diff_matrix(N, K)
    upper = [[0] * N for _ in range(K + 1)]
    for i in range(1, N):
        for row in range(K, -1, -1):
            j = pieces[row][i]
            upper[row][i] = max(upper[row][i - 1], lower[row][i - 1] + diffs[K - row][j])
    searches = []
    for i in range(K + 1):
        l = list(lower[i])
        l[-1] = max(l[-1], min_k)
        searches.append((pieces[i], l, upper[i]))
    return _pack(searches)


def facade_scheme(edit, max_errors, length):
    """Scheme + partition that fmc::search<Edit>(index, queries, errors, cb) selects for a query of `length`
    (search/SearchNg26.h:437-444, search/CachedSearchScheme.h:15-36,61-71)."""
    short = length == 2
    sch = h2(max_errors + (1 if short else 2), 0, max_errors)
    if not edit:
        sch = limit_to_hamming(sch)
    return sch, uniform_partition(sch[0].shape[1], length)
