// Stand-in for libsais 2.10.4 (IlyaGrebnov/libsais; call sites utils.h:104-124,592-595 of the reference).
// TEST INFRASTRUCTURE ONLY.  Any correct suffix sorter yields byte-identical output because all suffixes of
// the 0-delimited text are distinct; this one is prefix doubling with std::sort, fine for test-sized inputs.
#pragma once
#include <algorithm>
#include <cstdint>
#include <numeric>
#include <vector>

namespace sais_standin {
template <typename SymT, typename IdxT>
int suffix_sort(const SymT* T, IdxT* SA, int64_t n) {
    if (n <= 0) return 0;
    std::vector<int64_t> rank(n), tmp(n);
    std::vector<int64_t> sa(n);
    std::iota(sa.begin(), sa.end(), int64_t{0});
    for (int64_t i = 0; i < n; ++i) rank[i] = static_cast<int64_t>(T[i]) + 1;
    for (int64_t h = 1;; h *= 2) {
        auto key2 = [&](int64_t i) { return i + h < n ? rank[i + h] : int64_t{0}; };
        auto cmp = [&](int64_t a, int64_t b) {
            if (rank[a] != rank[b]) return rank[a] < rank[b];
            return key2(a) < key2(b);
        };
        std::sort(sa.begin(), sa.end(), cmp);
        tmp[sa[0]] = 1;
        for (int64_t i = 1; i < n; ++i) tmp[sa[i]] = tmp[sa[i - 1]] + (cmp(sa[i - 1], sa[i]) ? 1 : 0);
        rank.swap(tmp);
        if (rank[sa[n - 1]] == n) break;
    }
    for (int64_t i = 0; i < n; ++i) SA[i] = static_cast<IdxT>(sa[i]);
    return 0;
}
}
