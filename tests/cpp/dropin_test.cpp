// dropin_test.cpp -- drop-in proof: the SAME user code (same queries, same scheme objects, same delegate) runs once
// through the reference's own search functions on the reference's own index (CPU) and once through the fmb200 functions
// on `fmb200::attach(index)` (GPU).  The delegates receive the reference's cursor types in both cases, so the code
// behind the delegate (here: fmc::LocateLinear on the reported cursor) is shared.
// Built by tests/cpp/build.sh here (needs /root/reference), run by tests/test_gpu_cpp_shim.py on the GPU box.
#include <fmindex-collection/fmindex/BiFMIndex.h>
#include <fmindex-collection/fmindex/FMIndex.h>
#include <fmindex-collection/locate.h>
#include <fmindex-collection/search/Backtracking.h>
#include <fmindex-collection/search/SearchNg26.h>
#include <fmindex-collection/search/SearchNoErrors.h>
#include <fmindex-collection/search/SearchOneError.h>
#include <fmindex-collection/search/SearchPseudo.h>
#include <fmindex-collection/search/search.h>
#include <fmindex-collection/search_scheme/expand.h>
#include <fmindex-collection/search_scheme/generator/all.h>
#include <fmindex-collection/string/InterleavedBitvector.h>

#include <algorithm>
#include <cstdio>
#include <random>

#include "fmb200/fmb200.hpp"

using RefIndex = fmc::BiFMIndex<5, fmc::string::InterleavedBitvector16>;
using Row = std::array<uint64_t, 6>;

static int g_checks = 0, g_fail = 0;
#define CHECK(cond)                                                                         \
    do {                                                                                    \
        ++g_checks;                                                                         \
        if (!(cond)) { ++g_fail; std::fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); } \
    } while (0)

// the shared user code behind the delegate: collect the cursor and locate it with the REFERENCE's LocateLinear
struct Collector {
    RefIndex const& index;
    std::vector<Row> cursors;
    std::vector<Row> located;
    template <typename cursor_t>      // fmc::BiFMIndexCursor, or fmc::LeftBiFMIndexCursor from the exact branch of the façade
    void operator()(size_t qidx, cursor_t const& cursor, size_t e) {
        uint64_t lbRev = 0;
        if constexpr (requires { cursor.lbRev; }) lbRev = cursor.lbRev;
        cursors.push_back({qidx, cursor.lb, lbRev, cursor.len, cursor.steps, e});
        for (auto [sid, spos, offset] : fmc::LocateLinear{index, cursor}) located.push_back({qidx, sid, spos + offset, e, 0, 0});
    }
    void sort() { std::sort(cursors.begin(), cursors.end()); std::sort(located.begin(), located.end()); }
};

int main() {
    if (fmb_device_count() < 1) { std::fprintf(stderr, "no CUDA device: libfmb200 has no CPU fallback\n"); return 2; }
    std::mt19937_64 rng(7);
    std::vector<std::vector<uint8_t>> seqs;
    for (size_t len : {40000u, 9000u}) {
        std::vector<uint8_t> s(len);
        for (auto& c : s) c = 1 + rng() % 4;
        seqs.push_back(std::move(s));
    }
    // a repeat family so that cursors with len > 1 and multi-row locates occur
    for (int r = 0; r < 6; ++r) std::copy(seqs[0].begin() + 100, seqs[0].begin() + 400, seqs[0].begin() + 3000 + 2500 * r);
    auto index = RefIndex{seqs, /*samplingRate*/ 16, /*threadNbr*/ 1};
    auto dev = fmb200::attach<fmc::BiFMIndexCursor, fmc::LeftBiFMIndexCursor>(index);

    std::vector<std::vector<uint8_t>> queries;
    for (int i = 0; i < 500; ++i) {
        auto const& s = seqs[i % 2];
        size_t off = (i % 5 == 0) ? 100 + rng() % 200 : rng() % (s.size() - 50);
        std::vector<uint8_t> q(s.begin() + off, s.begin() + off + 50);
        if (i % 3 == 1) q[rng() % 50] = 1 + rng() % 4;
        if (i % 3 == 2) { q.insert(q.begin() + rng() % 50, uint8_t(1 + rng() % 4)); q.pop_back(); }
        queries.push_back(std::move(q));
    }

    for (size_t k : {1, 2}) {
        auto scheme = fmc::search_scheme::generator::optimum(0, k);                    // the reference's own scheme object
        auto partition = fmc::search_scheme::createUniformPartition(scheme, 50);
        Collector ref{index}, gpu{index};
        fmc::search_ng26::search<true>(index, queries, scheme, partition, [&](size_t q, auto c, size_t e) { ref(q, c, e); });
        fmb200::search_ng26::search<true>(dev, queries, scheme, partition, [&](size_t q, auto c, size_t e) {
            static_assert(std::same_as<decltype(c), fmc::BiFMIndexCursor<RefIndex>>);
            CHECK(c.index == &index);
            gpu(q, c, e);
        });
        ref.sort(); gpu.sort();
        CHECK(ref.cursors == gpu.cursors);
        CHECK(ref.located == gpu.located);
        CHECK(ref.cursors.size() > 400);
        Collector refh{index}, gpuh{index};
        fmc::search_ng26::search<false>(index, queries, scheme, partition, [&](size_t q, auto c, size_t e) { refh(q, c, e); });
        fmb200::search_ng26::search<false>(dev, queries, scheme, partition, [&](size_t q, auto c, size_t e) { gpuh(q, c, e); });
        refh.sort(); gpuh.sort();
        CHECK(refh.cursors == gpuh.cursors);
        // façade with the h2 scheme: fmc::search<Edit>(index, queries, errors, delegate)
        Collector reff{index}, gpuf{index};
        fmc::search<true>(index, queries, k, [&](size_t q, auto c, size_t e) { reff(q, c, e); });
        fmb200::search<true>(dev, queries, k, [&](size_t q, auto c, size_t e) { gpuf(q, c, e); });
        reff.sort(); gpuf.sort();
        CHECK(reff.cursors == gpuf.cursors);
        CHECK(reff.located == gpuf.located);
    }
    // exact search: delegate(qidx, LeftBiFMIndexCursor)
    {
        std::vector<std::array<uint64_t, 4>> ref, gpu;
        fmc::search_no_errors::search(index, queries, [&](size_t q, auto const& c) { ref.push_back({q, c.lb, c.len, c.steps}); });
        fmb200::search_no_errors::search(dev, queries, [&](size_t q, auto const& c) {
            static_assert(std::same_as<std::decay_t<decltype(c)>, fmc::LeftBiFMIndexCursor<RefIndex>>);
            gpu.push_back({q, c.lb, c.len, c.steps});
        });
        std::sort(ref.begin(), ref.end());
        std::sort(gpu.begin(), gpu.end());
        CHECK(ref == gpu);
        CHECK(!ref.empty());
    }
    // backtracking
    {
        std::vector<std::vector<uint8_t>> shortq(queries.begin(), queries.begin() + 80);
        for (auto& q : shortq) q.resize(16);
        Collector ref{index}, gpu{index};
        fmc::search_backtracking::search(index, shortq, 1, [&](size_t q, auto c, size_t e) { ref(q, c, e); });
        fmb200::search_backtracking::search(dev, shortq, 1, [&](size_t q, auto c, size_t e) { gpu(q, c, e); });
        ref.sort(); gpu.sort();
        CHECK(ref.cursors == gpu.cursors);
        CHECK(ref.located == gpu.located);
    }
    // search_best: list of (scheme, partition) pairs, and the maxErrors form (SearchNg26.h:448-487)
    {
        using Pair = std::tuple<fmc::search_scheme::Scheme, std::vector<size_t>>;
        std::vector<Pair> schemes;
        for (size_t k : {0, 1, 2}) {
            auto sch = fmc::search_scheme::generator::optimum(0, k);
            schemes.emplace_back(sch, fmc::search_scheme::createUniformPartition(sch, 50));
        }
        Collector ref{index}, gpu{index};
        fmc::search_ng26::search_best<true>(index, queries, schemes, [&](size_t q, auto c, size_t e) { ref(q, c, e); });
        fmb200::search_ng26::search_best<true>(dev, queries, schemes, [&](size_t q, auto c, size_t e) { gpu(q, c, e); });
        ref.sort(); gpu.sort();
        CHECK(ref.cursors == gpu.cursors);
        CHECK(ref.cursors.size() >= 400);
        // a batch in which no query matches exactly: level 0 finds nothing, level 1 ends the loop
        std::vector<std::vector<uint8_t>> damaged;
        for (size_t i = 0; i < queries.size(); ++i)
            if (i % 3 == 1) damaged.push_back(queries[i]);
        for (size_t maxErrors : {1, 2, 3}) {
            Collector r2{index}, g2{index};
            fmc::search_ng26::search_best<true>(index, damaged, maxErrors, [&](size_t q, auto c, size_t e) { r2(q, c, e); });
            fmb200::search_ng26::search_best<true>(dev, damaged, maxErrors, [&](size_t q, auto c, size_t e) { g2(q, c, e); });
            r2.sort(); g2.sort();
            CHECK(r2.cursors == g2.cursors);
        }
    }
    // hit limit n (search_n): the delegate must see the reference's calls IN THE REFERENCE'S ORDER -- nothing is sorted here.
    // search_ng26::search(..., n) with an explicit scheme and with maxErrors (SearchNg26.h:426-444), fmc::search_n
    // (search/search.h:37-45), search_best(..., n) (:448-470) and fmc::Search{... maxResults ...} (:47-75)
    for (size_t n : {1, 3, 8, 100000}) {
        for (size_t k : {1, 2}) {
            auto scheme = fmc::search_scheme::generator::optimum(0, k);
            auto partition = fmc::search_scheme::createUniformPartition(scheme, 50);
            Collector ref{index}, gpu{index}, refh{index}, gpuh{index}, reff{index}, gpuf{index};
            fmc::search_ng26::search<true>(index, queries, scheme, partition, [&](size_t q, auto c, size_t e) { ref(q, c, e); }, n);
            fmb200::search_ng26::search<true>(dev, queries, scheme, partition, [&](size_t q, auto c, size_t e) { gpu(q, c, e); }, n);
            CHECK(ref.cursors == gpu.cursors);
            CHECK(ref.located == gpu.located);
            fmc::search_ng26::search<false>(index, queries, scheme, partition, [&](size_t q, auto c, size_t e) { refh(q, c, e); }, n);
            fmb200::search_ng26::search<false>(dev, queries, scheme, partition, [&](size_t q, auto c, size_t e) { gpuh(q, c, e); }, n);
            CHECK(refh.cursors == gpuh.cursors);
            fmc::search_n<true>(index, queries, k, n, [&](size_t q, auto c, size_t e) { reff(q, c, e); });
            fmb200::search_n<true>(dev, queries, k, n, [&](size_t q, auto c, size_t e) { gpuf(q, c, e); });
            CHECK(reff.cursors == gpuf.cursors);
            CHECK(reff.located == gpuf.located);
            if (n == 1) CHECK(reff.located.size() <= queries.size());
        }
        {
            using Pair = std::tuple<fmc::search_scheme::Scheme, std::vector<size_t>>;
            std::vector<Pair> schemes;
            for (size_t k : {0, 1, 2}) {
                auto sch = fmc::search_scheme::generator::optimum(0, k);
                schemes.emplace_back(sch, fmc::search_scheme::createUniformPartition(sch, 50));
            }
            Collector ref{index}, gpu{index};
            fmc::search_ng26::search_best<true>(index, queries, schemes, [&](size_t q, auto c, size_t e) { ref(q, c, e); }, n);
            fmb200::search_ng26::search_best<true>(dev, queries, schemes, [&](size_t q, auto c, size_t e) { gpu(q, c, e); }, n);
            CHECK(ref.cursors == gpu.cursors);
        }
        {
            std::vector<std::array<uint64_t, 4>> ref, gpu;
            auto r1 = [&](size_t q, size_t sid, size_t pos, size_t e) { ref.push_back({q, sid, pos, e}); };
            auto r2 = [&](size_t q, size_t sid, size_t pos, size_t e) { gpu.push_back({q, sid, pos, e}); };
            fmc::Search{index, queries, true, size_t{2}, std::optional<size_t>{n}, r1}();
            fmb200::Search{dev, queries, true, size_t{2}, std::optional<size_t>{n}, r2}();
            CHECK(ref == gpu);
            CHECK(!ref.empty());
        }
    }
    // search_one_error (SearchOneError.h:126-145): queries of mixed lengths, including the degenerate lengths 0, 1, 2 and 3
    {
        std::vector<std::vector<uint8_t>> mixed(queries.begin(), queries.begin() + 120);
        for (size_t i = 0; i < mixed.size(); ++i) mixed[i].resize(i < 8 ? i % 4 : 12 + i % 30);
        Collector ref{index}, gpu{index};
        fmc::search_one_error::search(index, mixed, [&](size_t q, auto c, size_t e) { ref.cursors.push_back({q, c.lb, c.lbRev, c.len, c.steps, e}); });
        fmb200::search_one_error::search(dev, mixed, [&](size_t q, auto c, size_t e) {
            static_assert(std::same_as<decltype(c), fmc::BiFMIndexCursor<RefIndex>>);
            gpu.cursors.push_back({q, c.lb, c.lbRev, c.len, c.steps, e});
        });
        ref.sort(); gpu.sort();
        CHECK(ref.cursors == gpu.cursors);
        CHECK(ref.cursors.size() > 100);
    }
    // search_pseudo<false> (SearchPseudo.h:171-186) with the reference's expanded schemes (expand.h:146-165)
    for (size_t k : {1, 2, 3}) {
        for (int kind = 0; kind < 2; ++kind) {
            auto scheme = kind == 0 ? fmc::search_scheme::generator::optimum(0, std::min<size_t>(k, 2)) : fmc::search_scheme::generator::h2(k + 2, 0, k);
            auto expanded = fmc::search_scheme::expand(scheme, 50);
            Collector ref{index}, gpu{index};
            fmc::search_pseudo::search<false>(index, queries, expanded, [&](size_t q, auto c, size_t e) { ref(q, c, e); });
            fmb200::search_pseudo::search<false>(dev, queries, expanded, [&](size_t q, auto c, size_t e) { gpu(q, c, e); });
            ref.sort(); gpu.sort();
            CHECK(ref.cursors == gpu.cursors);
            CHECK(ref.located == gpu.located);
            CHECK(ref.cursors.size() > 300);
        }
    }
    // search_pseudo<true> (SearchPseudo.h:100-165): edit distance without redundancy filter, every duplicate included
    for (size_t k : {1, 2}) {
        for (int kind = 0; kind < 3; ++kind) {
            auto scheme = kind == 0 ? fmc::search_scheme::generator::optimum(0, k) : kind == 1 ? fmc::search_scheme::generator::h2(k + 2, 0, k)
                                                                                                 : fmc::search_scheme::generator::backtracking(k + 1, 0, k);
            std::vector<std::vector<uint8_t>> shortq(queries.begin(), queries.begin() + (kind == 2 ? 40 : 200));
            for (auto& q : shortq) q.resize(kind == 2 ? 14 : 30);
            auto expanded = fmc::search_scheme::expand(scheme, shortq[0].size());
            Collector ref{index}, gpu{index};
            fmc::search_pseudo::search<true>(index, shortq, expanded, [&](size_t q, auto c, size_t e) { ref.cursors.push_back({q, c.lb, c.lbRev, c.len, c.steps, e}); });
            fmb200::search_pseudo::search<true>(dev, shortq, expanded, [&](size_t q, auto c, size_t e) { gpu.cursors.push_back({q, c.lb, c.lbRev, c.len, c.steps, e}); });
            ref.sort(); gpu.sort();
            CHECK(ref.cursors == gpu.cursors);
            CHECK(ref.cursors.size() > shortq.size());
        }
    }
    // fmc::Search functor vs fmb200::Search: reportFunc(qidx, seqId, pos + offset, errors)
    {
        std::vector<std::array<uint64_t, 4>> ref, gpu;
        auto r1 = [&](size_t q, size_t sid, size_t pos, size_t e) { ref.push_back({q, sid, pos, e}); };
        auto r2 = [&](size_t q, size_t sid, size_t pos, size_t e) { gpu.push_back({q, sid, pos, e}); };
        fmc::Search{index, queries, true, size_t{1}, std::optional<size_t>{}, r1}();
        fmb200::Search{dev, queries, true, size_t{1}, std::optional<size_t>{}, r2}();
        std::sort(ref.begin(), ref.end());
        std::sort(gpu.begin(), gpu.end());
        CHECK(ref == gpu);
    }
    // String_c view of the attached image == the reference string, row by row on a stride
    for (size_t i = 0; i <= index.size(); i += 97) {
        for (uint8_t c = 0; c < 5; ++c) {
            CHECK(dev.bwt.rank(i, c) == index.bwt.rank(i, c));
            CHECK(dev.bwtRev.prefix_rank(i, c) == index.bwtRev.prefix_rank(i, c));
        }
        if (i < index.size()) {
            CHECK(dev.bwt.symbol(i) == index.bwt.symbol(i));
            auto a = dev.locate(i);
            auto b = index.locate(i);
            CHECK(std::get<0>(a) == std::get<0>(b) && std::get<1>(a) == std::get<1>(b) && std::get<2>(a) == std::get<2>(b));
        }
    }
    std::printf("dropin_test: %d checks, %d failed\n", g_checks, g_fail);
    return g_fail ? 1 : 0;
}
