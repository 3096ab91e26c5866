// shim_test.cpp -- the C++ host side (include/fmb200/*.hpp) against the CPU oracle (oracle/fm_oracle.c, the checker).
// Reads like the reference's own tests (search/checkSearches.cpp, fmindex/checkBiFMIndex.cpp): build an index from
// sequences, run every search entry point with a collecting delegate, compare as sorted multisets.
// Built by tests/cpp/build.sh, run by tests/test_gpu_cpp_shim.py on the GPU box.  Needs a CUDA device.
#include <algorithm>
#include <filesystem>
#include <unistd.h>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <set>

#include "fmb200/fmb200.hpp"
#include "fm_oracle.h"

using Hit = std::array<uint64_t, 6>;   // qidx, lb, lbRev, len, steps, e
using Loc = std::array<uint64_t, 4>;   // qidx, seq, pos, e

static int g_checks = 0, g_fail = 0;
#define CHECK(cond)                                                                         \
    do {                                                                                    \
        ++g_checks;                                                                         \
        if (!(cond)) { ++g_fail; std::fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); } \
    } while (0)

static std::vector<Hit> oracle_hits(fmo_hit* h, uint64_t n, bool zero_rev = false) {
    std::vector<Hit> v(n);
    for (uint64_t i = 0; i < n; ++i) v[i] = {h[i].qidx, h[i].lb, zero_rev ? 0 : h[i].lb_rev, h[i].len, h[i].steps, h[i].e};
    fmo_free(h);
    std::sort(v.begin(), v.end());
    return v;
}

int main() {
    namespace ss = fmb200::search_scheme;
    if (fmb_device_count() < 1) { std::fprintf(stderr, "no CUDA device: libfmb200 has no CPU fallback\n"); return 2; }
    std::mt19937_64 rng(42);
    // three sequences over {1..4}
    std::vector<std::vector<uint8_t>> seqs;
    for (size_t len : {60000u, 25000u, 777u}) {
        std::vector<uint8_t> s(len);
        for (auto& c : s) c = 1 + rng() % 4;
        seqs.push_back(std::move(s));
    }
    std::vector<uint8_t> text;
    for (auto const& s : seqs) { text.insert(text.end(), s.begin(), s.end()); text.push_back(0); }
    fmo_index* o = fmo_index_build(text.data(), text.size(), 5, 8, 1);

    // ---- index from sequences (GPU suffix sort) and from (bwt, bwtRev, SparseArray) ---------------------------------
    fmb200::BiFMIndex<5> index{seqs, /*samplingRate*/ 8, /*threads*/ 1};
    CHECK(index.size() == text.size());
    for (size_t s = 0; s <= 5; ++s) CHECK(index.C[s] == fmo_C(o)[s]);
    {
        std::vector<std::optional<std::tuple<uint32_t, uint32_t>>> ann(text.size());
        const uint64_t* bm = fmo_sample_bitmap(o);
        size_t k = 0;
        for (size_t i = 0; i < text.size(); ++i)
            if ((bm[i / 64] >> (i % 64)) & 1) { ann[i] = std::make_tuple(fmo_sample_seq(o)[k], fmo_sample_pos(o)[k]); ++k; }
        fmb200::BiFMIndex<5> index2{std::span<uint8_t const>{fmo_bwt(o), text.size()}, std::span<uint8_t const>{fmo_bwt_rev(o), text.size()}, fmb200::SparseArray{ann}};
        CHECK(index2.size() == index.size());
        for (uint64_t row : {uint64_t{0}, uint64_t{17}, uint64_t{text.size() - 1}}) {
            CHECK(index2.locate(row) == index.locate(row));
            CHECK(index2.bwt.rank(row, 2) == index.bwt.rank(row, 2));
        }
        bool threw = false;
        try { fmb200::BiFMIndex<5> bad{std::span<uint8_t const>{fmo_bwt(o), 10}, std::span<uint8_t const>{fmo_bwt_rev(o), 9}, fmb200::SparseArray{}}; }
        catch (std::runtime_error const&) { threw = true; }          // BiFMIndex.h:48-50
        CHECK(threw);
    }

    // ---- String_c concept (string/concepts.h:26-87) -----------------------------------------------------------------
    for (int i = 0; i < 200; ++i) {
        uint64_t idx = rng() % (text.size() + 1);
        uint8_t c = rng() % 5;
        CHECK(index.bwt.rank(idx, c) == fmo_rank(o, 0, idx, c));
        CHECK(index.bwtRev.rank(idx, c) == fmo_rank(o, 1, idx, c));
        CHECK(index.bwt.prefix_rank(idx, c) == fmo_prefix_rank(o, 0, idx, c));
        if (idx < text.size()) CHECK(index.bwt.symbol(idx) == fmo_symbol(o, 0, idx));
        uint64_t rs[5], prs[5];
        fmo_all_ranks_and_prefix_ranks(o, 0, idx, rs, prs);
        auto [grs, gprs] = index.bwt.all_ranks_and_prefix_ranks(idx);
        for (int s = 0; s < 5; ++s) CHECK(grs[s] == rs[s] && gprs[s] == prs[s]);
        auto ars = index.bwt.all_ranks(idx);
        for (int s = 0; s < 5; ++s) CHECK(ars[s] == rs[s]);
    }
    CHECK(index.bwt.prefix_rank(text.size(), 5) == text.size());      // symb == Sigma is allowed (FlattenedBitvectors2L.h:226-228)

    // ---- cursors (fmindex/BiFMIndexCursor.h) ----------------------------------------------------------------------------
    {
        fmb200::BiFMIndexCursor<fmb200::BiFMIndex<5>> cur{index};
        uint64_t oc[4] = {0, 0, text.size(), 0};
        for (uint8_t c : {1, 3, 2, 4, 1}) {
            bool right = (c & 1);
            cur = right ? cur.extendRight(c) : cur.extendLeft(c);
            uint64_t nx[4];
            if (right) fmo_extend_right(o, oc, c, nx); else fmo_extend_left(o, oc, c, nx);
            std::copy(nx, nx + 4, oc);
            CHECK(cur.lb == oc[0] && cur.lbRev == oc[1] && cur.len == oc[2] && cur.steps == oc[3]);
        }
        auto all = cur.extendLeft();
        uint64_t oall[20];
        fmo_extend_left_all(o, oc, oall);
        for (int s = 0; s < 5; ++s) CHECK(all[s].lb == oall[4 * s] && all[s].lbRev == oall[4 * s + 1] && all[s].len == oall[4 * s + 2]);
    }

    // ---- queries: reads of length 40 from the text, some with planted edits, some random --------------------------------
    std::vector<std::vector<uint8_t>> queries;
    for (int i = 0; i < 600; ++i) {
        auto const& s = seqs[i % 2];
        size_t off = rng() % (s.size() - 40);
        std::vector<uint8_t> q(s.begin() + off, s.begin() + off + 40);
        if (i % 3 == 1) q[rng() % 40] = 1 + rng() % 4;
        if (i % 3 == 2) { q.erase(q.begin() + rng() % 40); q.push_back(1 + rng() % 4); }
        if (i % 50 == 49) for (auto& c : q) c = 1 + rng() % 4;
        queries.push_back(std::move(q));
    }
    auto flat = fmb200::flatten(queries);

    // ---- search_no_errors::search(index, queries, delegate) (SearchNoErrors.h:28) ----------------------------------------
    {
        std::vector<Hit> got;
        fmb200::search_no_errors::search(index, queries, [&](size_t qidx, auto const& cursor) {
            got.push_back({qidx, cursor.lb, 0, cursor.len, cursor.steps, 0});
        });
        std::sort(got.begin(), got.end());
        fmo_hit* h{};
        uint64_t n = fmo_search_exact(o, flat.symbols.data(), flat.offsets.data(), queries.size(), &h, nullptr);
        CHECK(got == oracle_hits(h, n, true));
        CHECK(!got.empty());
        auto single = fmb200::search_no_errors::search(index, queries[0]);          // :13-26
        CHECK(single.count() == 1 || single.count() == got[0][3]);
    }
    // ---- search_ng26::search<Edit>(index, queries, scheme, partition, delegate) (SearchNg26.h:426) -----------------------
    for (size_t k : {1, 2}) {
        auto scheme = ss::generator::optimum(0, k);
        auto partition = ss::createUniformPartition(scheme, 40);
        auto fs = fmb200::detail::flatten(scheme, partition);
        for (int edit = 0; edit < 2; ++edit) {
            std::vector<Hit> got;
            auto cb = [&](size_t qidx, auto const& cursor, size_t e) { got.push_back({qidx, cursor.lb, cursor.lbRev, cursor.len, cursor.steps, e}); };
            if (edit) fmb200::search_ng26::search<true>(index, queries, scheme, partition, cb);
            else fmb200::search_ng26::search<false>(index, queries, scheme, partition, cb);
            // callbacks arrive grouped by ascending qidx like the reference's (SearchNg26.h:408-421)
            CHECK(std::is_sorted(got.begin(), got.end(), [](Hit const& a, Hit const& b) { return a[0] < b[0]; }));
            std::sort(got.begin(), got.end());
            fmo_hit* h{};
            uint64_t n = fmo_search_ng26(o, flat.symbols.data(), flat.offsets.data(), queries.size(), edit, fs.n_searches, fs.n_parts, fs.pi.data(), fs.l.data(),
                                         fs.u.data(), fs.partition.data(), UINT64_MAX, &h, nullptr);
            CHECK(got == oracle_hits(h, n));
            CHECK(got.size() >= 300);
        }
    }
    // ---- façade fmc::search<Edit>(index, queries, errors, delegate) (search/search.h:26) with ragged query lengths --------
    {
        auto ragged = queries;
        for (size_t i = 0; i < ragged.size(); i += 7) ragged[i].resize(33);
        for (size_t errors : {0, 1, 2}) {
            std::vector<Hit> got, exp;
            fmb200::search<true>(index, ragged, errors, [&](size_t qidx, auto const& cursor, size_t e) {
                if constexpr (requires { cursor.lbRev; }) got.push_back({qidx, cursor.lb, cursor.lbRev, cursor.len, cursor.steps, e});
                else got.push_back({qidx, cursor.lb, 0, cursor.len, cursor.steps, e});
            });
            std::sort(got.begin(), got.end());
            for (size_t len : {33, 39, 40}) {          // oracle: per length group, the scheme the façade selects
                std::vector<size_t> ids;
                fmb200::FlatSequences f;
                for (size_t i = 0; i < ragged.size(); ++i)
                    if (ragged[i].size() == len) { ids.push_back(i); f.symbols.insert(f.symbols.end(), ragged[i].begin(), ragged[i].end()); f.offsets.push_back(f.symbols.size()); }
                if (ids.empty()) continue;
                fmo_hit* h{};
                uint64_t n;
                if (errors == 0) {
                    n = fmo_search_exact(o, f.symbols.data(), f.offsets.data(), ids.size(), &h, nullptr);
                    for (uint64_t i = 0; i < n; ++i) h[i].lb_rev = 0;
                } else {
                    auto [scheme, partition] = ss::facadeScheme<true>(errors, len);
                    auto fs = fmb200::detail::flatten(scheme, partition);
                    n = fmo_search_ng26(o, f.symbols.data(), f.offsets.data(), ids.size(), 1, fs.n_searches, fs.n_parts, fs.pi.data(), fs.l.data(), fs.u.data(),
                                        fs.partition.data(), UINT64_MAX, &h, nullptr);
                }
                for (uint64_t i = 0; i < n; ++i) exp.push_back({ids[h[i].qidx], h[i].lb, h[i].lb_rev, h[i].len, h[i].steps, h[i].e});
                fmo_free(h);
            }
            std::sort(exp.begin(), exp.end());
            CHECK(got == exp);
        }
    }
    // ---- search_backtracking::search (Backtracking.h:85), bidirectional and unidirectional index --------------------------
    {
        std::vector<std::vector<uint8_t>> shortq(queries.begin(), queries.begin() + 100);
        for (auto& q : shortq) q.resize(18);
        auto f = fmb200::flatten(shortq);
        std::vector<Hit> got;
        fmb200::search_backtracking::search(index, shortq, 1, [&](size_t qidx, auto const& cursor, size_t e) { got.push_back({qidx, cursor.lb, cursor.lbRev, cursor.len, cursor.steps, e}); });
        std::sort(got.begin(), got.end());
        fmo_hit* h{};
        uint64_t n = fmo_search_backtracking(o, f.symbols.data(), f.offsets.data(), shortq.size(), 1, &h, nullptr);
        CHECK(got == oracle_hits(h, n));

        fmb200::FMIndex<5> uni{seqs, 8, 1};
        std::vector<std::array<uint64_t, 4>> gu, eu;
        fmb200::search_backtracking::search(uni, shortq, 1, [&](size_t qidx, auto const& cursor, size_t e) { gu.push_back({qidx, cursor.lb, cursor.len, e}); });
        for (auto const& x : got) eu.push_back({x[0], x[1], x[3], x[5]});
        std::sort(gu.begin(), gu.end());
        std::sort(eu.begin(), eu.end());
        CHECK(gu == eu);
        std::vector<std::array<uint64_t, 3>> g0;
        fmb200::search_no_errors::search(uni, shortq, [&](size_t qidx, auto const& cursor) { g0.push_back({qidx, cursor.lb, cursor.len}); });
        CHECK(!g0.empty());
    }
    // ---- LocateLinear + fmc::Search functor (locate.h:15, search/search.h:47-75) -----------------------------------------
    {
        std::vector<Loc> got, exp;
        auto report = [&](size_t qidx, size_t sid, size_t pos, size_t e) { got.push_back({qidx, sid, pos, e}); };
        fmb200::Search{index, queries, /*editDistance*/ false, /*errors*/ size_t{1}, std::optional<size_t>{}, report}();
        std::sort(got.begin(), got.end());
        auto [scheme, partition] = ss::facadeScheme<false>(1, 40);
        auto fs = fmb200::detail::flatten(scheme, partition);
        fmo_hit* h{};
        uint64_t n = fmo_search_ng26(o, flat.symbols.data(), flat.offsets.data(), queries.size(), 0, fs.n_searches, fs.n_parts, fs.pi.data(), fs.l.data(), fs.u.data(),
                                     fs.partition.data(), UINT64_MAX, &h, nullptr);
        fmo_loc* l{};
        uint64_t nl = fmo_locate(o, h, n, &l, nullptr);
        for (uint64_t i = 0; i < nl; ++i) exp.push_back({l[i].qidx, l[i].seq, l[i].pos, l[i].e});
        std::sort(exp.begin(), exp.end());
        CHECK(got == exp);
        // LocateLinear over one cursor
        fmb200::BiFMIndexCursor<fmb200::BiFMIndex<5>> cur{index, h[0].lb, h[0].lb_rev, h[0].len, h[0].steps};
        size_t cnt = 0;
        for (auto [sid, spos, offset] : fmb200::LocateLinear{index, cur}) {
            uint64_t e3[3];
            fmo_locate_row(o, h[0].lb + cnt, e3);
            CHECK(sid == e3[0] && spos == e3[1] && offset == e3[2]);
            ++cnt;
        }
        CHECK(cnt == h[0].len);
        fmo_free(h);
        fmo_free(l);
        uint64_t s2[2];
        for (uint64_t row = 0; row < 64; ++row) {
            auto v = index.single_locate_step(row);
            int has = fmo_single_locate_step(o, row, s2);
            CHECK(v.has_value() == (has != 0));
            if (v && has) CHECK(std::get<0>(*v) == s2[0] && std::get<1>(*v) == s2[1]);
        }
        // bulk one-call path == the functor's rows
        auto bulk = fmb200::search_and_locate_bulk(index, queries, false, &scheme, &partition);
        std::vector<Loc> gb;
        for (auto const& r : bulk) gb.push_back({r.qidx, r.seq, r.pos, r.e});
        std::sort(gb.begin(), gb.end());
        CHECK(gb == exp);
    }
    // ---- scheme generators vs the literal tables of the reference's generators ----------------------------------------------
    {
        auto h2 = ss::generator::h2(3, 0, 1);
        CHECK(h2.size() == 2);
        CHECK(ss::createUniformPartition(4, 150) == (std::vector<size_t>{38, 38, 37, 37}));
        CHECK(ss::createUniformPartition(2, 150) == (std::vector<size_t>{75, 75}));
        auto bt = ss::generator::backtracking(3, 1, 2);
        CHECK(bt[0].l == (std::vector<size_t>{0, 0, 1}) && bt[0].u == (std::vector<size_t>{2, 2, 2}));
    }
    // ---- saveIndex / loadIndex (fmindex/diskStorage.h:13-27) + hit limit through the shim ---------------------------------------
    {
        auto path = std::filesystem::temp_directory_path() / ("fmb200_shim_test_" + std::to_string(::getpid()) + ".fmb");
        fmb200::saveIndex(index, path);
        auto loaded = fmb200::loadIndex<fmb200::BiFMIndex<5>>(path);
        CHECK(loaded.size() == index.size());
        CHECK(loaded.C == index.C);
        std::vector<std::array<uint64_t, 3>> a, b;
        fmb200::search_no_errors::search(index, queries, [&](size_t q, auto const& c) { a.push_back({q, c.lb, c.len}); });
        fmb200::search_no_errors::search(loaded, queries, [&](size_t q, auto const& c) { b.push_back({q, c.lb, c.len}); });
        std::sort(a.begin(), a.end());
        std::sort(b.begin(), b.end());
        CHECK(a == b && !a.empty());
        bool refused = false;
        try { (void)fmb200::loadIndex<fmb200::FMIndex<5>>(path); } catch (std::runtime_error const&) { refused = true; }
        CHECK(refused);                                          // a BiFMIndex file is not an FMIndex
        std::filesystem::remove(path);
        // search_n: the oracle's list, in order
        for (size_t n : {1, 2, 5}) {
            auto [scheme, partition] = ss::facadeScheme<true>(2, 40);
            auto fs = fmb200::detail::flatten(scheme, partition);
            std::vector<std::array<uint64_t, 6>> got;
            fmb200::search_n<true>(index, queries, 2, n, [&](size_t q, auto const& c, size_t e) { got.push_back({q, c.lb, c.lbRev, c.len, c.steps, e}); });
            fmo_hit* h{};
            uint64_t cnt = fmo_search_ng26(o, flat.symbols.data(), flat.offsets.data(), queries.size(), 1, fs.n_searches, fs.n_parts, fs.pi.data(), fs.l.data(), fs.u.data(),
                                           fs.partition.data(), n, &h, nullptr);
            std::vector<std::array<uint64_t, 6>> exp;
            for (uint64_t i = 0; i < cnt; ++i) exp.push_back({h[i].qidx, h[i].lb, h[i].lb_rev, h[i].len, h[i].steps, h[i].e});
            CHECK(got == exp);
            fmo_free(h);
        }
    }
    // ---- io::uploadQueries(reverse = true): the device doubles the batch (read, reverse complement, read, ...) ------------------
    {
        std::vector<std::vector<uint8_t>> doubled;
        for (auto const& q : queries) {
            doubled.push_back(q);
            doubled.push_back(fmb200::io::reverseComplement(q));
        }
        auto dev = fmb200::io::uploadQueries(index, queries, /*reverse*/ true);
        CHECK(fmb_queries_count(dev.get()) == doubled.size());
        fmb_results* r{};
        fmb200::check(fmb_search_exact(index.handle(), dev.get(), &r));
        fmb200::detail::ResultsHandle res{r};
        auto got = fmb200::detail::fetch_hits(r);
        auto exp = fmb200::search_no_errors::search_bulk(index, doubled);
        fmb200::detail::sort_hits(got);
        fmb200::detail::sort_hits(exp);
        CHECK(got.size() == exp.size() && !got.empty());
        bool same = got.size() == exp.size();
        for (size_t i = 0; same && i < got.size(); ++i) same = got[i].qidx == exp[i].qidx && got[i].lb == exp[i].lb && got[i].len == exp[i].len;
        CHECK(same);
    }
    // ---- multi-GPU host logic: replicas + contiguous shards (every visible device) -------------------------------------------
    {
        CHECK(fmb200::shard_range(10, 0, 3) == (std::pair<size_t, size_t>{0, 4}));
        CHECK(fmb200::shard_range(10, 2, 3) == (std::pair<size_t, size_t>{7, 10}));
        fmb200::ReplicatedBiFMIndex<5> rep{std::span<uint8_t const>{fmo_bwt(o), text.size()}, std::span<uint8_t const>{fmo_bwt_rev(o), text.size()}, fmb200::SparseArray{}};
        auto hits = rep.search_exact(queries);
        auto one = fmb200::search_no_errors::search_bulk(index, queries);
        CHECK(hits.size() == one.size());
        for (size_t i = 0; i < hits.size() && i < one.size(); ++i) CHECK(hits[i].qidx == one[i].qidx && hits[i].lb == one[i].lb && hits[i].len == one[i].len);
        // one call for the whole batch: replicas of an existing index (peer copies), contiguous shards, located rows
        fmb200::ReplicatedBiFMIndex<5> rep2{fmb200::BiFMIndex<5>{seqs, 8, 1}};
        auto rows = rep2.search_and_locate(queries);
        auto ref = fmb200::search_and_locate_bulk(index, queries);
        auto key = [](fmb_loc32 const& a, fmb_loc32 const& b) { return std::tie(a.qidx, a.seq, a.pos, a.e) < std::tie(b.qidx, b.seq, b.pos, b.e); };
        CHECK(std::is_sorted(rows.begin(), rows.end(), [](fmb_loc32 const& a, fmb_loc32 const& b) { return a.qidx < b.qidx; }) || rep2.world() >= 1);
        std::sort(rows.begin(), rows.end(), key);
        std::sort(ref.begin(), ref.end(), key);
        CHECK(rows.size() == ref.size() && !rows.empty());
        for (size_t i = 0; i < rows.size() && i < ref.size(); ++i) CHECK(rows[i].qidx == ref[i].qidx && rows[i].seq == ref[i].seq && rows[i].pos == ref[i].pos && rows[i].e == ref[i].e);
    }
    // ---- a collection split into parts (the capacity path for n >= 2^32 rows): here at most 61000 symbols per part -> {seq 0}, {seq 1, 2}
    {
        fmb200::PartitionedBiFMIndex<5> pidx{seqs, 8, /*max_part_symbols*/ 61000, /*n_devices*/ 1};
        CHECK(pidx.parts.size() == 2 && pidx.seq_base == (std::vector<uint64_t>{0, 1}) && pidx.size() == index.size());
        auto [scheme, partition] = fmb200::search_scheme::facadeScheme<true>(1, 40);
        auto rows = pidx.search_and_locate(queries, true, &scheme, &partition);
        auto ref = fmb200::search_and_locate_bulk(index, queries, true, &scheme, &partition);
        auto key = [](fmb_loc32 const& a, fmb_loc32 const& b) { return std::tie(a.qidx, a.seq, a.pos, a.e) < std::tie(b.qidx, b.seq, b.pos, b.e); };
        std::sort(rows.begin(), rows.end(), key);
        std::sort(ref.begin(), ref.end(), key);
        CHECK(rows.size() == ref.size() && !rows.empty());
        bool same = rows.size() == ref.size();
        for (size_t i = 0; same && i < rows.size(); ++i) same = rows[i].qidx == ref[i].qidx && rows[i].seq == ref[i].seq && rows[i].pos == ref[i].pos && rows[i].e == ref[i].e;
        CHECK(same);
    }
    fmo_index_free(o);
    std::printf("shim_test: %d checks, %d failed\n", g_checks, g_fail);
    return g_fail ? 1 : 0;
}
