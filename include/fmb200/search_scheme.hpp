// fmb200/search_scheme.hpp -- host-side search-scheme inputs of the k-error search (tiny, pure std).
//
// Mirrors the types and helpers of the reference that feed search_ng26 (paths relative to
// /root/reference/src/fmindex-collection/):
//   search_scheme/Search.h:19-28     struct Search{pi, l, u}
//   search_scheme/Scheme.h:13        using Scheme = std::vector<Search>
//   search_scheme/generator/optimum.h:11-75, backtracking.h:15-22, h2.h:128-149
//   search_scheme/expand.h:37-180    expandCount, expand;  :301-343 limitToHamming, createUniformPartition;  isValid.h:55-93 isValid
//   search/CachedSearchScheme.h:15-36,61-71   the scheme fmc::search<Edit>(index, queries, k, cb) selects
// The generators are pinned against tables produced by the reference's own generators
// (tests/golden/ref_vectors.json, tests/cpp/shim_test.cpp).
#pragma once
#include <algorithm>
#include <optional>
#include <cstddef>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

namespace fmb200::search_scheme {

struct Search {
    std::vector<size_t> pi;   // order in which the parts are searched, zero based
    std::vector<size_t> l;    // minimum number of errors after each part
    std::vector<size_t> u;    // maximum number of errors after each part
    bool operator==(Search const&) const = default;
};
using Scheme = std::vector<Search>;

// expand.h:301-319: with substitutions only the error count grows by at most one per part
inline Search limitToHamming(Search s) {
    size_t const len = s.pi.size();
    for (size_t i = len; i-- > 1;) {
        if (s.l[i] == 0) break;
        s.l[i - 1] = std::max(s.l[i - 1], s.l[i] - 1);
    }
    for (size_t i = 1; i < len; ++i) s.u[i] = std::min(s.u[i], s.u[i - 1] + 1);
    return s;
}
inline Scheme limitToHamming(Scheme ss) {
    for (auto& s : ss) s = limitToHamming(std::move(s));
    return ss;
}

// expand.h:324-343
inline std::vector<size_t> createUniformPartition(size_t parts, size_t totalSum) {
    if (parts == 0 || totalSum < parts) throw std::invalid_argument("createUniformPartition: need 0 < parts <= totalSum");
    std::vector<size_t> counts(parts, totalSum / parts);
    for (size_t i = 0; i < totalSum % parts; ++i) counts[i] += 1;
    return counts;
}
inline std::vector<size_t> createUniformPartition(Scheme const& ss, size_t totalSum) {
    if (ss.empty()) throw std::invalid_argument("createUniformPartition: empty scheme");
    return createUniformPartition(ss[0].pi.size(), totalSum);
}

// isValid.h:55-93: pi connected and covering 0, l and u non-decreasing, l <= u
inline bool isValid(Search const& s) {
    if (s.pi.empty() || s.pi.size() != s.u.size() || s.pi.size() != s.l.size()) return false;
    size_t lo = s.pi.front(), hi = s.pi.front();
    for (size_t i = 1; i < s.pi.size(); ++i) {
        if (s.pi[i] == hi + 1) hi = s.pi[i];
        else if (s.pi[i] + 1 == lo) lo = s.pi[i];
        else return false;
    }
    if (lo != 0) return false;
    for (size_t i = 1; i < s.pi.size(); ++i)
        if (s.l[i - 1] > s.l[i] || s.u[i - 1] > s.u[i]) return false;
    for (size_t i = 0; i < s.pi.size(); ++i)
        if (s.l[i] > s.u[i]) return false;
    return true;
}
inline bool isValid(Scheme const& ss) {
    for (auto const& s : ss)
        if (!isValid(s) || s.pi.size() != ss.front().pi.size()) return false;
    return true;
}

// expand.h:37-48: `newLen` symbols spread over `oldLen` parts, the first newLen % oldLen parts one longer
inline std::vector<size_t> expandCount(size_t oldLen, size_t newLen) {
    if (oldLen == 0 || newLen == 0) throw std::invalid_argument("expandCount: empty scheme or query");
    std::vector<size_t> counts(oldLen, newLen / oldLen);
    for (size_t i = 0; i < newLen % oldLen; ++i) counts[i] += 1;
    return counts;
}
// expand.h:146-180: one pi / l / u entry per query symbol (the input of search_pseudo).  A part is walked forwards when the next
// part lies to its right (the first part: like the second one, :21-28); every symbol of a part carries the part's upper bound, the
// last one its lower bound and the others the lower bound of the part before (:105-121).  Searches that are not valid afterwards
// are dropped, as in the reference.
inline std::optional<Search> expand(Search const& s, std::vector<size_t> const& counts) {
    size_t const P = s.pi.size();
    if (counts.size() != P) throw std::invalid_argument("expand: one count per part is needed");
    std::vector<size_t> starts(P, 0);
    for (size_t i = 1; i < P; ++i) starts[i] = starts[i - 1] + counts[i - 1];
    Search r;
    for (size_t i = 0; i < P; ++i) {
        bool const forward = i == 0 ? (P == 1 || s.pi[1] > s.pi[0]) : s.pi[i] > s.pi[i - 1];
        size_t const b = starts[s.pi[i]], c = counts[s.pi[i]];
        for (size_t j = 0; j < c; ++j) r.pi.push_back(forward ? b + j : b + c - 1 - j);
        for (size_t j = 0; j < c; ++j) r.u.push_back(s.u[i]);
        for (size_t j = 0; j + 1 < c; ++j) r.l.push_back(i > 0 ? s.l[i - 1] : 0);
        if (c > 0) r.l.push_back(s.l[i]);
        else if (!r.l.empty()) r.l.back() = s.l[i];
    }
    if (!isValid(r)) return std::nullopt;
    return r;
}
inline std::optional<Search> expand(Search const& s, size_t newLen) { return expand(s, expandCount(s.pi.size(), newLen)); }
inline Scheme expand(Scheme const& ss, std::vector<size_t> const& counts) {
    Scheme r;
    for (auto const& s : ss)
        if (auto o = expand(s, counts)) r.push_back(std::move(*o));
    return r;
}
inline Scheme expand(Scheme const& ss, size_t newLen) {
    Scheme r;
    for (auto const& s : ss)
        if (auto o = expand(s, newLen)) r.push_back(std::move(*o));
    return r;
}

namespace generator {

// generator/optimum.h:11-75 -- the optimal schemes of Kianfar et al. for K <= 2 (the BASELINE configs); pi zero based
inline Scheme optimum(size_t minK, size_t K) {
    auto S = [](std::vector<size_t> pi, std::vector<size_t> l, std::vector<size_t> u) { return Search{std::move(pi), std::move(l), std::move(u)}; };
    if (minK == 0 && K == 0) return {S({0}, {0}, {0})};
    if (minK == 0 && K == 1) return {S({0, 1}, {0, 0}, {0, 1}), S({1, 0}, {0, 1}, {0, 1})};
    if (minK == 1 && K == 1) return {S({0, 1}, {0, 1}, {0, 1}), S({1, 0}, {0, 1}, {0, 1})};
    if (minK == 0 && K == 2)
        return {S({0, 1, 2, 3}, {0, 0, 1, 1}, {0, 0, 2, 2}), S({2, 1, 0, 3}, {0, 0, 0, 0}, {0, 1, 1, 2}), S({3, 2, 1, 0}, {0, 0, 0, 2}, {0, 1, 2, 2})};
    if (minK == 1 && K == 2)
        return {S({0, 1, 2, 3}, {0, 0, 0, 1}, {0, 0, 2, 2}), S({2, 1, 0, 3}, {0, 0, 1, 1}, {0, 1, 1, 2}), S({3, 2, 1, 0}, {0, 0, 0, 2}, {0, 1, 2, 2})};
    throw std::invalid_argument("optimum(" + std::to_string(minK) + "," + std::to_string(K) + ") is not tabulated");
}

// generator/backtracking.h:15-22: one search over N parts, left to right
inline Scheme backtracking(size_t N, size_t minK, size_t K) {
    if (N == 0) throw std::invalid_argument("backtracking: N must be positive");
    Search s;
    s.pi.resize(N);
    for (size_t i = 0; i < N; ++i) s.pi[i] = i;
    s.l.assign(N, 0);
    s.u.assign(N, K);
    s.l.back() = minK;
    return {s};
}

// generator/h2.h:128-149: K+1 searches over N parts
inline Scheme h2(size_t N, size_t minK, size_t K) {
    // the reference asserts N >= K (h2.h:129-131) but its lower-bound table indexes part N-K-1: N == K is out of range
    if (N == 0 || minK > K || K >= N) throw std::invalid_argument("h2: need N > 0 and minK <= K < N");
    auto pi_of = [&](size_t row, size_t n) {
        row = K - row;
        return n < N - row ? n + row : N - n - 1;
    };
    std::vector<std::vector<size_t>> diff(K + 1, std::vector<size_t>(N, 0));
    for (size_t i = K; i < N; ++i)
        for (size_t row = 0; row <= K; ++row) diff[row][i] = K - row;
    for (size_t i = 0; i < K; ++i) {
        for (size_t row = 0; row < K; ++row) diff[row][i] = (row + K - i) % K;
        diff[K][i] = K;
    }
    auto valid = [&](size_t row, size_t n, size_t v) {
        if (row == n) return false;
        if (row > n) {
            for (size_t i = 0; i < n; ++i)
                if (diff[row][i] < v) return false;
            return true;
        }
        for (size_t i = row + 1; i < n; ++i)
            if (diff[row][i] > v) return false;
        return true;
    };
    for (size_t i = 0; i < N; ++i) {
        for (size_t j = 0; j <= K; ++j) {
            if (i == j || diff[j][i] == 0 || valid(j, i, diff[j][i])) continue;
            size_t k = j + 1;
            for (; k <= K; ++k)
                if (valid(j, i, diff[k][i]) && valid(k, i, diff[j][i])) break;
            if (k > K) throw std::logic_error("h2: no valid exchange");
            std::swap(diff[k][i], diff[j][i]);
        }
    }
    std::vector<std::vector<size_t>> pieces(K + 1, std::vector<size_t>(N)), lower(K + 1, std::vector<size_t>(N, 0)), upper(K + 1, std::vector<size_t>(N, 0));
    for (size_t row = 0; row <= K; ++row)
        for (size_t i = 0; i < N; ++i) pieces[row][i] = pi_of(row, i);
    for (size_t i = 0; i <= K; ++i)
        for (size_t j = 0; j <= K - i; ++j) lower[i][N - j - 1] = i;
    for (size_t i = 1; i < N; ++i)
        for (size_t row = K + 1; row-- > 0;) {
            size_t j = pieces[row][i];
            upper[row][i] = std::max(upper[row][i - 1], lower[row][i - 1] + diff[K - row][j]);
        }
    Scheme ss;
    for (size_t i = 0; i <= K; ++i) {
        Search s{pieces[i], lower[i], upper[i]};
        s.l.back() = std::max(s.l.back(), minK);
        ss.push_back(std::move(s));
    }
    return ss;
}

}  // namespace generator

// The (scheme, partition) pair that fmc::search<Edit>(index, queries, maxErrors, cb) selects for a query of `length`
// (search/SearchNg26.h:437-444 -> getCachedSearchScheme / getCachedPartition)
template <bool Edit>
inline std::tuple<Scheme, std::vector<size_t>> facadeScheme(size_t maxErrors, size_t length) {
    bool const shortLen = length == 2;
    Scheme ss = generator::h2(maxErrors + (shortLen ? 1 : 2), 0, maxErrors);
    if constexpr (!Edit) ss = limitToHamming(std::move(ss));
    auto partition = createUniformPartition(ss[0].pi.size(), length);
    return {std::move(ss), std::move(partition)};
}

}  // namespace fmb200::search_scheme
