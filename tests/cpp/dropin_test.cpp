// dropin_test.cpp -- drop-in proof: the SAME user code (same queries, same scheme objects, same delegate) runs once
// through the reference's own search functions on the reference's own index (CPU) and once through the fmb200 functions
// on `fmb200::attach(index)` (GPU).  The delegates receive the reference's cursor types in both cases, so the code
// behind the delegate (here: fmc::LocateLinear on the reported cursor) is shared.
// Built by tests/cpp/build.sh here (needs /root/reference), run by tests/test_gpu_cpp_shim.py on the GPU box.
#include <fmindex-collection/fmindex/BiFMIndex.h>
#include <fmindex-collection/fmindex/FMIndex.h>
#include <fmindex-collection/locate.h>
#include <fmindex-collection/search/Backtracking.h>
#include <fmindex-collection/search/BacktrackingWithBuffers.h>
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
        // one query at a time: Backtracking.h:90-98 and the buffered variant (BacktrackingWithBuffers.h:93-106)
        Collector ref1{index}, gpu1{index}, refb{index}, gpub{index};
        using RefBuf = std::vector<std::pair<fmc::select_cursor_t<RefIndex>, size_t>>;
        RefBuf rb1, rb2;
        std::vector<int> gb1, gb2;
        for (size_t q = 0; q < 20; ++q) {
            fmc::search_backtracking::search(index, shortq[q], 1, [&](auto c, size_t e) { ref1(q, c, e); });
            fmb200::search_backtracking::search(dev, shortq[q], 1, [&](auto c, size_t e) { gpu1(q, c, e); });
            fmc::search_backtracking_with_buffers::search(index, shortq[q], 1, rb1, rb2, [&](auto c, size_t e) { refb(q, c, e); });
            fmb200::search_backtracking_with_buffers::search(dev, shortq[q], 1, gb1, gb2, [&](auto c, size_t e) { gpub(q, c, e); });
        }
        ref1.sort(); gpu1.sort(); refb.sort(); gpub.sort();
        CHECK(!ref1.cursors.empty() && ref1.cursors == gpu1.cursors && ref1.located == gpu1.located);
        CHECK(refb.cursors == gpub.cursors && refb.located == gpub.located && refb.cursors == ref1.cursors);
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
    // ---- seam 2: the device block layout as a String_c inside the REFERENCE's index type ----------------
    {
        static_assert(fmc::String_c<fmb200::HostMirror<5>> && fmc::String_c<fmb200::HostMirror<21>> && fmc::String_c<fmb200::HostMirror<2>>);
        using MirrorIndex = fmc::BiFMIndex<5, fmb200::HostMirror>;
        auto mirror = MirrorIndex{seqs, /*samplingRate*/ 16, /*threadNbr*/ 1};
        // the mirror's blocks are the bytes the kernels read, for both directions
        for (int dir = 0; dir < 2; ++dir) {
            uint64_t bytes = 0;
            uint32_t stride = 0;
            CHECK(fmb_index_export_blocks(dev.handle(), dir, nullptr, 0, &bytes, &stride) == 0 && stride == 32);
            std::vector<uint8_t> devBlocks(bytes);
            CHECK(fmb_index_export_blocks(dev.handle(), dir, devBlocks.data(), devBlocks.size(), &bytes, &stride) == 0);
            auto const host = dir == 0 ? mirror.bwt.blocks() : mirror.bwtRev.blocks();
            CHECK(host.size() == devBlocks.size() && std::equal(host.begin(), host.end(), devBlocks.begin()));
        }
        // String_c answers of the mirror == the reference's interleaved bit vectors, row by row
        bool same = true;
        for (size_t idx = 0; idx <= index.size() && same; idx += 1 + idx % 7) {
            for (uint8_t s = 0; s < 5 && same; ++s)
                same = mirror.bwt.rank(idx, s) == index.bwt.rank(idx, s) && mirror.bwt.prefix_rank(idx, s) == index.bwt.prefix_rank(idx, s) &&
                       mirror.bwtRev.rank(idx, s) == index.bwtRev.rank(idx, s);
            if (idx < index.size()) same = same && mirror.bwt.symbol(idx) == index.bwt.symbol(idx);
            same = same && mirror.bwt.all_ranks_and_prefix_ranks(idx) == index.bwt.all_ranks_and_prefix_ranks(idx);
        }
        CHECK(same);
        // the reference's k-error search on the mirror layout == the reference on its own layout == the GPU
        auto scheme = fmc::search_scheme::generator::optimum(0, 2);
        auto partition = fmc::search_scheme::createUniformPartition(scheme, 50);
        std::vector<Row> a, b, c;
        fmc::search_ng26::search<true>(mirror, queries, scheme, partition, [&](size_t q, auto cur, size_t e) { a.push_back({q, cur.lb, cur.lbRev, cur.len, cur.steps, e}); });
        fmc::search_ng26::search<true>(index, queries, scheme, partition, [&](size_t q, auto cur, size_t e) { b.push_back({q, cur.lb, cur.lbRev, cur.len, cur.steps, e}); });
        fmb200::search_ng26::search<true>(dev, queries, scheme, partition, [&](size_t q, auto cur, size_t e) { c.push_back({q, cur.lb, cur.lbRev, cur.len, cur.steps, e}); });
        std::sort(a.begin(), a.end()); std::sort(b.begin(), b.end()); std::sort(c.begin(), c.end());
        CHECK(!a.empty() && a == b && a == c);
        // generic layout (sigma = 21) against the device's blocks
        std::vector<std::vector<uint8_t>> prot(2);
        for (auto& s : prot) { s.resize(3000); for (auto& ch : s) ch = 1 + rng() % 20; }
        fmb200::BiFMIndex<21> pdev{prot, 8};
        auto pmir = fmc::BiFMIndex<21, fmb200::HostMirror>{prot, 8, 1};
        uint64_t bytes = 0;
        uint32_t stride = 0;
        CHECK(fmb_index_export_blocks(pdev.handle(), 0, nullptr, 0, &bytes, &stride) == 0 && stride == 128);
        std::vector<uint8_t> pb(bytes);
        CHECK(fmb_index_export_blocks(pdev.handle(), 0, pb.data(), pb.size(), &bytes, &stride) == 0);
        CHECK(pmir.bwt.blocks().size() == pb.size() && std::equal(pb.begin(), pb.end(), pmir.bwt.blocks().begin()));
    }
    // ---- index variants: BiFMIndex<...>::ReuseRev (one BWT of text + reversed text for both directions) and ::NoDelim (FirstSymb = 0,
    //      omega-sorted text), fmindex/BiFMIndex.h:22-28, 53-72 -- the reference's own objects, attached, searched by both sides ----------
    {
        auto compare = [&](auto const& rindex, auto const& rdev, std::vector<std::vector<uint8_t>> const& qs, char const* what) {
            using RI = std::remove_cvref_t<decltype(rindex)>;
            for (size_t k : {1, 2}) {
                auto scheme = fmc::search_scheme::generator::optimum(0, k);
                auto partition = fmc::search_scheme::createUniformPartition(scheme, 50);
                for (int edit = 0; edit < 2; ++edit) {
                    std::vector<Row> a, b, la, lb;
                    auto ra = [&](size_t q, auto cur, size_t e) {
                        a.push_back({q, cur.lb, cur.lbRev, cur.len, cur.steps, e});
                        for (auto [sid, spos, offset] : fmc::LocateLinear{rindex, cur}) la.push_back({q, sid, spos + offset, e, 0, 0});
                    };
                    auto rb = [&](size_t q, auto cur, size_t e) {
                        static_assert(std::same_as<decltype(cur), fmc::BiFMIndexCursor<RI>>);
                        b.push_back({q, cur.lb, cur.lbRev, cur.len, cur.steps, e});
                        for (auto [sid, spos, offset] : fmb200::LocateLinear{rdev, cur}) lb.push_back({q, sid, spos + offset, e, 0, 0});
                    };
                    if (edit) { fmc::search_ng26::search<true>(rindex, qs, scheme, partition, ra); fmb200::search_ng26::search<true>(rdev, qs, scheme, partition, rb); }
                    else { fmc::search_ng26::search<false>(rindex, qs, scheme, partition, ra); fmb200::search_ng26::search<false>(rdev, qs, scheme, partition, rb); }
                    std::sort(a.begin(), a.end()); std::sort(b.begin(), b.end()); std::sort(la.begin(), la.end()); std::sort(lb.begin(), lb.end());
                    if (a != b || la != lb || a.empty()) std::fprintf(stderr, "variant %s k=%zu edit=%d: %zu vs %zu cursors, %zu vs %zu rows\n", what, k, edit, a.size(), b.size(), la.size(), lb.size());
                    CHECK(!a.empty() && a == b);
                    CHECK(la == lb);
                }
            }
            std::vector<Row> a, b;
            fmc::search_no_errors::search(rindex, qs, [&](size_t q, auto cur) { a.push_back({q, cur.lb, 0, cur.len, cur.steps, 0}); });
            fmb200::search_no_errors::search(rdev, qs, [&](size_t q, auto cur) { b.push_back({q, cur.lb, 0, cur.len, cur.steps, 0}); });
            std::sort(a.begin(), a.end()); std::sort(b.begin(), b.end());
            CHECK(!a.empty() && a == b);
        };
        // ReuseRev: the text must hold every sequence and its reversal (includeReversedInput)
        using RevIndex = RefIndex::ReuseRev;
        auto rindex = RevIndex{seqs, /*samplingRate*/ 16, /*threadNbr*/ 1, /*seqOffset*/ 0, /*includeReversedInput*/ true};
        auto rdev = fmb200::attach<fmc::BiFMIndexCursor, fmc::LeftBiFMIndexCursor>(rindex);
        fmb_index_info info{};
        CHECK(fmb_index_get_info(rdev.handle(), &info) == 0 && info.bidirectional == 1 && (info.flags & FMB_INDEX_REUSE_REV));
        compare(rindex, rdev, queries, "ReuseRev");
        // NoDelim over the alphabet {0, 1, 2, 3}: symbol 0 is an ordinary symbol
        using NoDelimIndex = fmc::BiFMIndex<4, fmc::string::InterleavedBitvector16>::NoDelim;
        auto seqs0 = seqs;
        for (auto& sq : seqs0) for (auto& ch : sq) ch -= 1;
        auto queries0 = queries;
        for (auto& q : queries0) for (auto& ch : q) ch -= 1;
        auto nindex = NoDelimIndex{seqs0, /*samplingRate*/ 16, /*threadNbr*/ 1};
        auto ndev = fmb200::attach<fmc::BiFMIndexCursor, fmc::LeftBiFMIndexCursor>(nindex);
        CHECK(fmb_index_get_info(ndev.handle(), &info) == 0 && (info.flags & FMB_INDEX_NO_DELIM) && info.occ_block_bytes == 32);
        compare(nindex, ndev, queries0, "NoDelim");
        // both at once
        using MirroredIndex = NoDelimIndex::ReuseRev;
        auto mindex = MirroredIndex{seqs0, /*samplingRate*/ 16, /*threadNbr*/ 1, /*seqOffset*/ 0, /*includeReversedInput*/ true};
        auto mdev = fmb200::attach<fmc::BiFMIndexCursor, fmc::LeftBiFMIndexCursor>(mindex);
        compare(mindex, mdev, queries0, "NoDelim::ReuseRev");
    }
    std::printf("dropin_test: %d checks, %d failed\n", g_checks, g_fail);
    return g_fail ? 1 : 0;
}
