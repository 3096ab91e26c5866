// fmb200/search.hpp -- the reference's batch search entry points, executed by the CUDA kernels of libfmb200.so.
//
// Same names, argument meaning and delegate shapes as the reference (paths relative to
// /root/reference/src/fmindex-collection/):
//   search/SearchNoErrors.h:13-26,28-85   search_no_errors::search(index, query) / (index, queries, delegate(qidx, cursor))
//   search/SearchNg26.h:426-444           search_ng26::search<Edit>(index, queries, scheme, partition, delegate(qidx, cursor, e))
//                                         search_ng26::search<Edit>(index, queries, maxErrors, delegate)
//   search/Backtracking.h:85-98           search_backtracking::search(index, queries, maxError, delegate(qidx, cursor, e)) / one query
//   search/BacktrackingWithBuffers.h:93   search_backtracking_with_buffers::search(index, query, maxError, buffer1, buffer2, delegate(cursor, e))
//   search/SearchOneError.h:126-145       search_one_error::search(index, queries, delegate(qidx, cursor, e))
//   search/SearchPseudo.h:171-186         search_pseudo::search<Edit>(index, queries, expanded scheme, delegate(qidx, cursor, e))
//   search/search.h:14-75                 fmc::search<Edit>(index, queries, errors, delegate), fmc::Search{...}()
//   locate.h:15-57                        LocateLinear{index, cursor}
// Differences a caller can observe: delegates are invoked after the device finished, grouped by
// ascending qidx (the reference's batched exact search retires queries out of order; tests compare sorted).  With a hit
// limit `n` (search_ng26::search(..., n), search_best(..., n), fmc::search_n, Search::maxResults) the delegate sees exactly
// the reference's calls in the reference's order: ascending qidx, within a query the order of its depth-first search,
// the cursor that crosses the limit clipped (SearchNg26.h:408-423).  `*_bulk` variants return the raw records without
// per-hit callbacks.
#pragma once
#include <algorithm>
#include <limits>
#include <map>
#include <memory>
#include <optional>

#include "index.hpp"
#include "search_scheme.hpp"

namespace fmb200 {

namespace detail {
struct QueriesDeleter { void operator()(fmb_queries* p) const { fmb_queries_destroy(p); } };
struct ResultsDeleter { void operator()(fmb_results* p) const { fmb_results_destroy(p); } };
using QueriesHandle = std::unique_ptr<fmb_queries, QueriesDeleter>;
using ResultsHandle = std::unique_ptr<fmb_results, ResultsDeleter>;

inline QueriesHandle upload(fmb_index const* ix, FlatSequences const& f, size_t first = 0, size_t count = std::numeric_limits<size_t>::max()) {
    count = std::min(count, f.size() - first);
    fmb_queries* q{};
    check(fmb_queries_upload(&q, ix, f.symbols.data(), f.offsets.data() + first, count));
    return QueriesHandle{q};
}
inline std::vector<fmb_hit> fetch_hits(fmb_results* r) {
    std::vector<fmb_hit> hits(fmb_results_count(r));
    check(fmb_results_fetch_hits(r, hits.data(), hits.size()));
    return hits;
}
constexpr size_t kNoLimit = std::numeric_limits<size_t>::max();
// hit-limited results keep the discovery order inside a query: only group them by ascending qidx
inline void group_hits_by_query(std::vector<fmb_hit>& hits) {
    std::stable_sort(hits.begin(), hits.end(), [](fmb_hit const& a, fmb_hit const& b) { return a.qidx < b.qidx; });
}
inline void sort_hits(std::vector<fmb_hit>& hits) {
    std::sort(hits.begin(), hits.end(), [](fmb_hit const& a, fmb_hit const& b) {
        return std::tie(a.qidx, a.e, a.lb, a.len, a.steps, a.lb_rev) < std::tie(b.qidx, b.e, b.lb, b.len, b.steps, b.lb_rev);
    });
}
struct FlatScheme {
    uint32_t n_searches{}, n_parts{};
    std::vector<uint32_t> pi, l, u, partition;
};
// any range of {pi, l, u} records: fmb200::search_scheme::Scheme or the reference's fmc::search_scheme::Scheme
template <typename T>
concept SchemeLike = std::ranges::range<T> && requires(T const& t) {
    { std::ranges::begin(t)->pi.size() } -> std::convertible_to<size_t>;
    { std::ranges::begin(t)->l[0] } -> std::convertible_to<size_t>;
    { std::ranges::begin(t)->u[0] } -> std::convertible_to<size_t>;
};
template <SchemeLike scheme_t>
FlatScheme flatten(scheme_t const& ss, std::vector<size_t> const& partition) {
    if (std::ranges::empty(ss)) throw std::invalid_argument("fmb200: empty search scheme");
    FlatScheme f;
    f.n_searches = static_cast<uint32_t>(std::ranges::size(ss));
    f.n_parts = static_cast<uint32_t>(std::ranges::begin(ss)->pi.size());
    if (partition.size() != f.n_parts) throw std::invalid_argument("fmb200: partition size does not match the scheme");
    for (auto const& s : ss) {
        if (s.pi.size() != f.n_parts || s.l.size() != f.n_parts || s.u.size() != f.n_parts) throw std::invalid_argument("fmb200: ragged search scheme");
        for (size_t p = 0; p < f.n_parts; ++p) {
            f.pi.push_back(static_cast<uint32_t>(s.pi[p]));
            f.l.push_back(static_cast<uint32_t>(s.l[p]));
            f.u.push_back(static_cast<uint32_t>(s.u[p]));
        }
    }
    for (auto p : partition) f.partition.push_back(static_cast<uint32_t>(p));
    return f;
}
template <typename index_t>
constexpr bool is_bidirectional = requires(index_t const& ix) { ix.bwtRev; };

template <typename index_t>
auto make_cursor(index_t const& index, fmb_hit const& h) {
    using cursor_t = select_cursor_t<index_t>;
    if constexpr (requires { cursor_t{index, size_t{}, size_t{}, size_t{}, size_t{}}; }) {
        return cursor_t{index, h.lb, h.lb_rev, h.len, h.steps};
    } else {
        return cursor_t{index, h.lb, h.len};
    }
}
template <typename index_t>
auto make_left_cursor(index_t const& index, fmb_hit const& h) {
    using cursor_t = select_left_cursor_t<index_t>;
    if constexpr (requires { cursor_t{index, size_t{}, size_t{}, size_t{}}; }) {
        return cursor_t{index, h.lb, h.len, h.steps};
    } else {
        return cursor_t{index, h.lb, h.len};
    }
}
}  // namespace detail

// =====================================================================================================================
namespace search_no_errors {

// bulk form: one record per query with a non-empty interval
template <typename index_t, Sequences queries_t>
std::vector<fmb_hit> search_bulk(index_t const& index, queries_t const& queries, fmb_stats* stats = nullptr) {
    auto flat = flatten(queries);
    auto q = detail::upload(index.handle(), flat);
    fmb_results* r{};
    check(fmb_search_exact(index.handle(), q.get(), &r));
    detail::ResultsHandle res{r};
    if (stats) check(fmb_results_get_stats(r, stats));
    return detail::fetch_hits(r);
}

// SearchNoErrors.h:28-85
template <typename index_t, Sequences queries_t, typename delegate_t>
void search(index_t const& index, queries_t const& queries, delegate_t&& delegate, size_t /*BatchSize*/ = 32) {
    for (auto const& h : search_bulk(index, queries)) delegate(static_cast<size_t>(h.qidx), detail::make_left_cursor(index, h));
}

// SearchNoErrors.h:13-26: single query, returns the cursor (possibly empty)
template <typename index_t, Sequence query_t>
auto search(index_t const& index, query_t const& query) {
    std::array<std::span<uint8_t const>, 1> one;
    std::vector<uint8_t> copy(std::ranges::size(query));
    std::ranges::copy(query, copy.begin());
    one[0] = copy;
    auto hits = search_bulk(index, one);
    fmb_hit h{0, 0, 0, 0, copy.size(), 0};
    if (!hits.empty()) h = hits[0];
    return detail::make_left_cursor(index, h);
}

}  // namespace search_no_errors

// =====================================================================================================================
namespace search_ng26 {

template <bool Edit = true, typename index_t, Sequences queries_t, detail::SchemeLike scheme_t>
std::vector<fmb_hit> search_bulk(index_t const& index, queries_t const& queries, scheme_t const& scheme,
                                 std::vector<size_t> const& partition, fmb_stats* stats = nullptr, size_t n = detail::kNoLimit) {
    static_assert(detail::is_bidirectional<index_t>, "search schemes need a bidirectional index (extendRight)");
    auto fs = detail::flatten(scheme, partition);
    auto flat = flatten(queries);
    auto q = detail::upload(index.handle(), flat);
    fmb_results* r{};
    if (n == detail::kNoLimit) check(fmb_search_scheme(index.handle(), q.get(), Edit ? 1 : 0, fs.n_searches, fs.n_parts, fs.pi.data(), fs.l.data(), fs.u.data(), fs.partition.data(), &r));
    else check(fmb_search_scheme_n(index.handle(), q.get(), Edit ? 1 : 0, fs.n_searches, fs.n_parts, fs.pi.data(), fs.l.data(), fs.u.data(), fs.partition.data(), n, &r));
    detail::ResultsHandle res{r};
    if (stats) check(fmb_results_get_stats(r, stats));
    auto hits = detail::fetch_hits(r);
    if (n == detail::kNoLimit) detail::sort_hits(hits);      // with a limit the device returns (qidx, discovery order)
    return hits;
}

// SearchNg26.h:426-433: explicit scheme + partition
template <bool Edit = true, typename index_t, Sequences queries_t, detail::SchemeLike scheme_t, typename delegate_t>
void search(index_t const& index, queries_t&& queries, scheme_t const& scheme, std::vector<size_t> const& partition,
            delegate_t&& delegate, size_t n = std::numeric_limits<size_t>::max()) {
    if (n == 0) return;                                       // SearchNg26.h:410
    for (auto const& h : search_bulk<Edit>(index, queries, scheme, partition, nullptr, n)) delegate(static_cast<size_t>(h.qidx), detail::make_cursor(index, h), static_cast<size_t>(h.e));
}

// SearchNg26.h:436-444: scheme selected per query length (h2 with maxErrors+2 parts, CachedSearchScheme.h:15-36).
// Queries are grouped by length; every group is one device call.
template <bool Edit = true, typename index_t, Sequences queries_t>
std::vector<fmb_hit> search_bulk(index_t const& index, queries_t const& queries, size_t maxErrors, size_t n = detail::kNoLimit) {
    std::map<size_t, std::vector<size_t>> by_len;
    size_t qidx = 0;
    for (auto const& q : queries) by_len[std::ranges::size(q)].push_back(qidx++);
    std::vector<fmb_hit> all;
    for (auto const& [len, ids] : by_len) {
        if (len == 0) continue;      // an empty query has no scheme (createUniformPartition asserts totalSum > 0)
        if (len < maxErrors + (len == 2 ? 1 : 2)) throw std::runtime_error("fmb200: query shorter than the number of scheme parts");
        auto [scheme, partition] = search_scheme::facadeScheme<Edit>(maxErrors, len);
        std::vector<std::span<uint8_t const>> group;
        std::vector<std::vector<uint8_t>> store;
        store.reserve(ids.size());
        for (auto id : ids) {
            auto const& q = queries[id];
            store.emplace_back(std::ranges::size(q));
            std::ranges::copy(q, store.back().begin());
            group.emplace_back(store.back());
        }
        auto hits = search_bulk<Edit>(index, group, scheme, partition, nullptr, n);
        for (auto& h : hits) h.qidx = ids[h.qidx];
        all.insert(all.end(), hits.begin(), hits.end());
    }
    if (n == detail::kNoLimit) detail::sort_hits(all);
    else detail::group_hits_by_query(all);
    return all;
}
template <bool Edit = true, typename index_t, Sequences queries_t, typename delegate_t>
void search(index_t const& index, queries_t&& queries, size_t maxErrors, delegate_t&& delegate, size_t n = std::numeric_limits<size_t>::max()) {
    if (n == 0) return;
    for (auto const& h : search_bulk<Edit>(index, queries, maxErrors, n)) delegate(static_cast<size_t>(h.qidx), detail::make_cursor(index, h), static_cast<size_t>(h.e));
}

// SearchNg26.h:448-470: search_best with a list of (scheme, partition) pairs -- per query, the first pair that yields a hit wins.
// Device form: pair 0 runs on all queries, pair 1 on the queries still without a hit, and so on.  The hit limit n counts per
// pair (`ct` is reset for every pair, :453).  All queries of a call must have the length the partitions sum to.
template <bool Edit = true, typename index_t, Sequences queries_t, typename schemes_t, typename delegate_t>
    requires requires(schemes_t const& ss) { std::get<0>(*std::begin(ss)); std::get<1>(*std::begin(ss)); }
void search_best(index_t const& index, queries_t&& queries, schemes_t const& search_schemes, delegate_t&& delegate,
                 size_t n = std::numeric_limits<size_t>::max()) {
    // n == 0: `ct == n` never holds and `cur.len = n - ct` wraps in the reference -- nothing sensible to mirror
    if (n == 0) throw std::invalid_argument("fmb200: search_best with n == 0");
    size_t const Q = std::ranges::size(queries);
    std::vector<size_t> pending(Q);
    for (size_t i = 0; i < Q; ++i) pending[i] = i;
    std::vector<fmb_hit> all;
    for (auto const& entry : search_schemes) {
        if (pending.empty()) break;
        auto const& scheme = std::get<0>(entry);
        auto const& partition = std::get<1>(entry);
        std::vector<std::span<uint8_t const>> group;
        std::vector<std::vector<uint8_t>> store;
        store.reserve(pending.size());
        for (auto id : pending) {
            auto const& q = queries[id];
            store.emplace_back(std::ranges::size(q));
            std::ranges::copy(q, store.back().begin());
            group.emplace_back(store.back());
        }
        auto hits = search_bulk<Edit>(index, group, scheme, partition, nullptr, n);
        std::vector<char> found(pending.size(), 0);
        for (auto& h : hits) {
            found[h.qidx] = 1;
            h.qidx = pending[h.qidx];
        }
        all.insert(all.end(), hits.begin(), hits.end());
        std::vector<size_t> rest;
        for (size_t i = 0; i < pending.size(); ++i)
            if (!found[i]) rest.push_back(pending[i]);
        pending.swap(rest);
    }
    if (n == detail::kNoLimit) detail::sort_hits(all);
    else detail::group_hits_by_query(all);
    for (auto const& h : all) delegate(static_cast<size_t>(h.qidx), detail::make_cursor(index, h), static_cast<size_t>(h.e));
}

// SearchNg26.h:472-487: search_best with a maximal error count -- error levels 0 .. maxErrors-1 are tried in turn on ALL queries
// and the first level at which any query has a hit ends the loop.  The reference's semantics are kept as they are, including
// that the inner call is `search(index, queries, i, …)` with its default Edit = true whatever search_best's own parameter is.
template <bool Edit = true, typename index_t, Sequences queries_t, typename delegate_t>
void search_best(index_t const& index, queries_t&& queries, size_t maxErrors, delegate_t&& delegate, size_t n = std::numeric_limits<size_t>::max()) {
    if (n == 0) return;                                       // every inner search returns at once (:410)
    for (size_t i = 0; i < maxErrors; ++i) {
        auto hits = search_bulk<true>(index, queries, i, n);
        if (hits.empty()) continue;
        for (auto const& h : hits) delegate(static_cast<size_t>(h.qidx), detail::make_cursor(index, h), static_cast<size_t>(h.e));
        break;
    }
}

}  // namespace search_ng26

// =====================================================================================================================
// search/SearchOneError.h:126-145: at most one substitution.  The reference hard-codes a two-search scheme: left half exact then
// right half with <= 1 mismatch (reports e = 0 and e = 1), right half exact then left half with exactly one mismatch -- i.e.
// {pi 01, l 00, u 01}, {pi 10, l 01, u 01} over the partition {L - L/2, L/2} (:27,73), searched with Hamming distance.
namespace search_one_error {

template <typename index_t, Sequences queries_t>
std::vector<fmb_hit> search_bulk(index_t const& index, queries_t const& queries) {
    std::map<size_t, std::vector<size_t>> by_len;
    size_t qidx = 0;
    for (auto const& q : queries) by_len[std::ranges::size(q)].push_back(qidx++);
    std::vector<fmb_hit> all;
    for (auto const& [len, ids] : by_len) {
        if (len == 0) {
            // no symbol to search: search_left_to_right reports the cursor of the whole index with 0 errors (:24-37, 53)
            for (auto id : ids) all.push_back(fmb_hit{id, 0, 0, index.size(), 0, 0});
            continue;
        }
        search_scheme::Scheme scheme;
        std::vector<size_t> partition;
        if (len == 1) {
            // the right-to-left search has an empty exact half: every symbol other than the query's is a one-error hit (:86-100)
            scheme = {search_scheme::Search{{0}, {0}, {1}}};
            partition = {1};
        } else {
            scheme = {search_scheme::Search{{0, 1}, {0, 0}, {0, 1}}, search_scheme::Search{{1, 0}, {0, 1}, {0, 1}}};
            partition = {len - len / 2, len / 2};
        }
        std::vector<std::span<uint8_t const>> group;
        std::vector<std::vector<uint8_t>> store;
        store.reserve(ids.size());
        for (auto id : ids) {
            auto const& q = queries[id];
            store.emplace_back(std::ranges::size(q));
            std::ranges::copy(q, store.back().begin());
            group.emplace_back(store.back());
        }
        auto hits = search_ng26::search_bulk<false>(index, group, scheme, partition);
        for (auto& h : hits) h.qidx = ids[h.qidx];
        all.insert(all.end(), hits.begin(), hits.end());
    }
    detail::sort_hits(all);
    return all;
}

template <typename index_t, Sequences queries_t, typename delegate_t>
void search(index_t const& index, queries_t const& queries, delegate_t&& delegate) {
    for (auto const& h : search_bulk(index, queries)) delegate(static_cast<size_t>(h.qidx), detail::make_cursor(index, h), static_cast<size_t>(h.e));
}

}  // namespace search_one_error

// =====================================================================================================================
// search/SearchPseudo.h:171-186: search with an EXPANDED scheme (one pi / l / u entry per query symbol, search_scheme/expand.h:146-165),
// Hamming distance.  The device kernel works on parts, so the expanded searches are folded back: a part ends where the walk
// changes direction, where u changes, and at every position whose lower bound is not already implied by an earlier part end
// (e never decreases, so `l[pos] <= e` holds once a part end with the same or a larger bound was passed -- the argument that makes
// expand()'s own lower bounds, SearchNg26's per-part check and this per-symbol check agree).  All searches are cut at the union of
// those boundaries.  The edit-distance form of search_pseudo enumerates alignments without the redundancy filter of search_ng26
// (more duplicates): it runs on the PSEUDO instantiation of the same kernel (fmb_search_scheme_pseudo).
namespace search_pseudo {

namespace detail_pseudo {
template <fmb200::detail::SchemeLike scheme_t>
std::tuple<search_scheme::Scheme, std::vector<size_t>> fold(scheme_t const& expanded) {
    size_t const L = std::ranges::begin(expanded)->pi.size();
    if (L == 0) throw std::invalid_argument("fmb200: empty expanded search");
    std::vector<char> cut(L + 1, 0);          // cut[t]: a part boundary between text positions t-1 and t
    for (auto const& s : expanded) {
        if (s.pi.size() != L || s.l.size() != L || s.u.size() != L) throw std::invalid_argument("fmb200: ragged expanded search scheme");
        size_t implied = 0;                   // largest lower bound checked at a part end so far
        for (size_t pos = 0; pos + 1 < L; ++pos) {
            bool const right = pos == 0 ? s.pi[1] > s.pi[0] : s.pi[pos] > s.pi[pos - 1];
            bool const next_right = s.pi[pos + 1] > s.pi[pos];
            bool const adjacent = next_right ? s.pi[pos + 1] == s.pi[pos] + 1 : s.pi[pos + 1] + 1 == s.pi[pos];
            bool end = !adjacent || right != next_right || s.u[pos + 1] != s.u[pos] || s.l[pos] > implied;
            if (!end) continue;
            implied = std::max<size_t>(implied, s.l[pos]);
            cut[next_right ? s.pi[pos + 1] : s.pi[pos + 1] + 1] = 1;
        }
    }
    std::vector<size_t> partition, part_of(L);
    size_t start = 0;
    for (size_t t = 1; t <= L; ++t)
        if (t == L || cut[t]) {
            for (size_t i = start; i < t; ++i) part_of[i] = partition.size();
            partition.push_back(t - start);
            start = t;
        }
    search_scheme::Scheme folded;
    for (auto const& s : expanded) {
        search_scheme::Search f;
        for (size_t pos = 0; pos < L; ++pos) {
            size_t const part = part_of[s.pi[pos]];
            if (f.pi.empty() || f.pi.back() != part) {
                if (std::find(f.pi.begin(), f.pi.end(), part) != f.pi.end()) throw std::invalid_argument("fmb200: expanded search visits a part twice");
                f.pi.push_back(part);
                f.l.push_back(s.l[pos]);
                f.u.push_back(s.u[pos]);
            } else {
                if (s.u[pos] != f.u.back()) throw std::invalid_argument("fmb200: upper bound changes inside a part");
                f.l.back() = s.l[pos];       // the bound of the part's last symbol
            }
        }
        folded.push_back(std::move(f));
    }
    return {std::move(folded), std::move(partition)};
}
}  // namespace detail_pseudo

template <bool EditDistance, typename index_t, Sequences queries_t, fmb200::detail::SchemeLike scheme_t>
std::vector<fmb_hit> search_bulk(index_t const& index, queries_t const& queries, scheme_t const& expanded) {
    auto [scheme, partition] = detail_pseudo::fold(expanded);
    if constexpr (!EditDistance) {
        return search_ng26::search_bulk<false>(index, queries, scheme, partition);
    } else {
        // search_distance (SearchPseudo.h:100-165): the scheme kernel with the redundancy filter switched off.  The device walks the
        // first part to the right; with errors allowed inside a first part that the expanded scheme walks to the left the two would
        // enumerate different alignments -- refused (every generator of the reference starts with an error-free part or walks right).
        size_t i = 0;
        for (auto const& e : expanded) {
            if (scheme[i].u[0] > 0 && e.pi.size() > 1 && e.pi[1] < e.pi[0]) throw std::invalid_argument("fmb200::search_pseudo<true>: errors inside a first part that is searched to the left");
            ++i;
        }
        static_assert(fmb200::detail::is_bidirectional<index_t>, "search schemes need a bidirectional index (extendRight)");
        auto fs = fmb200::detail::flatten(scheme, partition);
        auto flat = flatten(queries);
        auto q = fmb200::detail::upload(index.handle(), flat);
        fmb_results* r{};
        check(fmb_search_scheme_pseudo(index.handle(), q.get(), 1, fs.n_searches, fs.n_parts, fs.pi.data(), fs.l.data(), fs.u.data(), fs.partition.data(), &r));
        fmb200::detail::ResultsHandle res{r};
        auto hits = fmb200::detail::fetch_hits(r);
        fmb200::detail::sort_hits(hits);
        return hits;
    }
}

template <bool EditDistance, typename index_t, Sequences queries_t, fmb200::detail::SchemeLike scheme_t, typename delegate_t>
void search(index_t const& index, queries_t&& queries, scheme_t const& expanded, delegate_t&& delegate) {
    for (auto const& h : search_bulk<EditDistance>(index, queries, expanded)) delegate(static_cast<size_t>(h.qidx), fmb200::detail::make_cursor(index, h), static_cast<size_t>(h.e));
}

}  // namespace search_pseudo

// =====================================================================================================================
namespace search_backtracking {

template <typename index_t, Sequences queries_t>
std::vector<fmb_hit> search_bulk(index_t const& index, queries_t const& queries, size_t maxError) {
    // the device kernel takes batches of equal length; group like search_ng26 above
    std::map<size_t, std::vector<size_t>> by_len;
    size_t qidx = 0;
    for (auto const& q : queries) by_len[std::ranges::size(q)].push_back(qidx++);
    std::vector<fmb_hit> all;
    for (auto const& [len, ids] : by_len) {
        if (len == 0) continue;
        FlatSequences flat;
        flat.symbols.reserve(len * ids.size());
        for (auto id : ids) {
            for (auto c : queries[id]) flat.symbols.push_back(static_cast<uint8_t>(c));
            flat.offsets.push_back(flat.symbols.size());
        }
        auto q = detail::upload(index.handle(), flat);
        fmb_results* r{};
        check(fmb_search_backtracking(index.handle(), q.get(), static_cast<uint32_t>(maxError), &r));
        detail::ResultsHandle res{r};
        auto hits = detail::fetch_hits(r);
        for (auto& h : hits) h.qidx = ids[h.qidx];
        all.insert(all.end(), hits.begin(), hits.end());
    }
    detail::sort_hits(all);
    return all;
}

// Backtracking.h:85-98
template <typename index_t, Sequences queries_t, typename delegate_t>
void search(index_t const& index, queries_t&& queries, size_t maxError, delegate_t&& delegate) {
    for (auto const& h : search_bulk(index, queries, maxError)) delegate(static_cast<size_t>(h.qidx), detail::make_cursor(index, h), static_cast<size_t>(h.e));
}

// Backtracking.h:90-98: one query, delegate(cursor, errors)
template <typename index_t, Sequence query_t, typename delegate_t>
    requires(!Sequences<query_t>)
void search(index_t const& index, query_t const& query, size_t maxError, delegate_t&& delegate) {
    std::vector<std::vector<uint8_t>> one(1);
    for (auto c : query) one[0].push_back(static_cast<uint8_t>(c));
    for (auto const& h : search_bulk(index, one, maxError)) delegate(detail::make_cursor(index, h), static_cast<size_t>(h.e));
}

}  // namespace search_backtracking

// search/BacktrackingWithBuffers.h:93-106: the same search with caller-provided frontier buffers, one query, delegate(cursor, errors).
// The reference keeps its breadth-first frontier in buffer1 / buffer2; here the frontier lives in the shared memory of the kernel, the
// buffers are accepted for signature compatibility and left empty, as the reference leaves them (:66).
namespace search_backtracking_with_buffers {
template <typename index_t, Sequence query_t, typename buffer_t, typename delegate_t>
void search(index_t const& index, query_t const& query, size_t maxError, buffer_t& buffer1, buffer_t& buffer2, delegate_t&& delegate) {
    buffer1.clear();
    buffer2.clear();
    search_backtracking::search(index, query, maxError, delegate);
}
}  // namespace search_backtracking_with_buffers

// =====================================================================================================================
// locate.h:15-57: range over the (seqId, pos, offset) entries of every row of a cursor.  All rows are located by
// one kernel launch when the range is constructed.
template <typename index_t, typename cursor_t>
struct LocateLinear {
    using LEntry = std::tuple<uint32_t, uint32_t, size_t>;
    index_t const& index;
    cursor_t cursor;
    std::vector<LEntry> entries;
    LocateLinear(index_t const& ix, cursor_t const& cur) : index{ix}, cursor{cur} {
        std::vector<uint64_t> rows(cur.len);
        for (size_t i = 0; i < cur.len; ++i) rows[i] = cur.lb + i;
        entries = fmb200::locate_rows(ix, rows);
    }
    auto begin() const { return entries.begin(); }
    auto end() const { return entries.end(); }
};
template <typename index_t, typename cursor_t>
LocateLinear(index_t const&, cursor_t const&) -> LocateLinear<index_t, cursor_t>;

// =====================================================================================================================
// search/search.h:26-35: k == 0 -> exact search, else search_ng26 with the h2 scheme
template <bool EditDistance, typename index_t, Sequences queries_t, typename delegate_t>
void search(index_t const& index, queries_t const& queries, size_t errors, delegate_t&& delegate) {
    if (errors == 0) {
        for (auto const& h : search_no_errors::search_bulk(index, queries)) {
            if constexpr (detail::is_bidirectional<index_t>) {
                delegate(static_cast<size_t>(h.qidx), detail::make_left_cursor(index, h), size_t{0});
            } else {
                delegate(static_cast<size_t>(h.qidx), detail::make_cursor(index, h), size_t{0});
            }
        }
    } else {
        search_ng26::search<EditDistance>(index, queries, errors, std::forward<delegate_t>(delegate));
    }
}

// search/search.h:37-45: at most n rows per query, in the order the reference finds them (always through search_ng26, also for
// errors == 0: the h2 scheme with two error-free parts)
template <bool EditDistance, typename index_t, Sequences queries_t, typename delegate_t>
void search_n(index_t const& index, queries_t const& queries, size_t errors, size_t n, delegate_t&& delegate) {
    search_ng26::search<EditDistance>(index, queries, errors, std::forward<delegate_t>(delegate), n);
}

// search/search.h:47-75: search + locate, reportFunc(qidx, seqId, pos + offset, errors).  One device pass
// (fmb_search_and_locate for explicit schemes is used by search_and_locate_bulk below).
template <typename index_t, Sequences queries_t, typename delegate_t>
struct Search {
    index_t const& index;
    queries_t const& queries;
    bool editDistance{true};
    size_t errors{0};
    std::optional<size_t> maxResults{};
    delegate_t const& reportFunc;
    void operator()() {
        std::vector<fmb_hit> hits;
        if (maxResults) {
            if (*maxResults == 0) return;
            hits = editDistance ? search_ng26::search_bulk<true>(index, queries, errors, *maxResults) : search_ng26::search_bulk<false>(index, queries, errors, *maxResults);
        } else if (errors == 0) hits = search_no_errors::search_bulk(index, queries);
        else hits = editDistance ? search_ng26::search_bulk<true>(index, queries, errors) : search_ng26::search_bulk<false>(index, queries, errors);
        // locate all rows of all hits in one launch
        std::vector<uint64_t> rows;
        for (auto const& h : hits)
            for (uint64_t i = 0; i < h.len; ++i) rows.push_back(h.lb + i);
        auto entries = locate_rows(index, rows);
        size_t at = 0;
        for (auto const& h : hits)
            for (uint64_t i = 0; i < h.len; ++i, ++at) {
                auto [sid, spos, offset] = entries[at];
                reportFunc(static_cast<size_t>(h.qidx), sid, spos + offset, static_cast<size_t>(h.e));
            }
    }
};
template <typename index_t, typename queries_t, typename delegate_t>
Search(index_t const&, queries_t const&, bool, size_t, std::optional<size_t>, delegate_t const&) -> Search<index_t, queries_t, delegate_t>;

// Bulk one-call path (fmb_search_and_locate): host queries in, located rows (qidx, seq, pos + offset, e) out, upload /
// kernels / download pipelined over query chunks.  scheme == nullptr selects exact search.
template <typename index_t, Sequences queries_t, detail::SchemeLike scheme_t = search_scheme::Scheme>
std::vector<fmb_loc32> search_and_locate_bulk(index_t const& index, queries_t const& queries, bool edit = false,
                                              scheme_t const* scheme = nullptr, std::vector<size_t> const* partition = nullptr,
                                              fmb_stats* stats = nullptr) {
    auto flat = flatten(queries);
    detail::FlatScheme fs;
    if (scheme) fs = detail::flatten(*scheme, *partition);
    std::vector<fmb_loc32> out(std::max<size_t>(flat.size() * 2, 1024));
    for (;;) {
        uint64_t n_out = 0;
        int rc = fmb_search_and_locate(index.handle(), flat.symbols.data(), flat.offsets.data(), flat.size(), edit ? 1 : 0, fs.n_searches, fs.n_parts,
                                       fs.pi.data(), fs.l.data(), fs.u.data(), fs.partition.data(), out.data(), out.size(), &n_out, stats);
        if (rc == FMB_EOVERFLOW && n_out > out.size()) {
            out.resize(n_out);
            continue;
        }
        check(rc);
        out.resize(n_out);
        return out;
    }
}

}  // namespace fmb200
