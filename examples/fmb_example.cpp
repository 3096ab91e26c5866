// fmb_example.cpp -- the reference's example program (src/example/main.cpp) on the B200 engine, reduced to its data path:
//   FASTA reference -> BiFMIndex (built on the GPU, cached next to the FASTA as <ref>.fmb like the example's <ref>.index,
//   example/utils.h:107-141) -> FASTA reads (+ reverse complements unless --no-reverse, main.cpp:71) -> k-error search
//   (--mode all | besthits, --maxhitsperquery, main.cpp:167-212) -> locate -> "queryId seqId pos" lines (--save_output, :260-266).
// usage: fmb_example --ref ref.fa --query reads.fa [--max_k 2] [--hamming] [--no-reverse] [--packed] [--mode all|besthits]
//                    [--maxhitsperquery N] [--save_output out.txt] [--sampling_rate 16] [--no-index-cache]
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>

#include "fmb200/fmb200.hpp"

int main(int argc, char** argv) {
    std::string ref, query, out, mode = "all";
    size_t k = 0, rate = 16, maxhits = 0;
    bool reverse = true, hamming = false, cache = true, packed = false;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() -> std::string { if (i + 1 >= argc) { std::fprintf(stderr, "%s needs a value\n", a.c_str()); std::exit(2); } return argv[++i]; };
        if (a == "--ref") ref = next();
        else if (a == "--query") query = next();
        else if (a == "--save_output") out = next();
        else if (a == "--max_k") k = std::stoul(next());
        else if (a == "--sampling_rate") rate = std::stoul(next());
        else if (a == "--maxhitsperquery") maxhits = std::stoul(next());
        else if (a == "--mode") mode = next();
        else if (a == "--no-reverse") reverse = false;
        else if (a == "--packed") packed = true;
        else if (a == "--hamming") hamming = true;
        else if (a == "--no-index-cache") cache = false;
        else { std::fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }
    if (ref.empty() || query.empty()) { std::fprintf(stderr, "usage: fmb_example --ref ref.fa --query reads.fa [options]\n"); return 2; }
    constexpr size_t Sigma = 5;
    using Index = fmb200::BiFMIndex<Sigma>;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto secs = [](auto a, auto b) { return std::chrono::duration<double>(b - a).count(); };
    try {
        auto t0 = now();
        auto index = [&]() -> Index {
            auto path = ref + ".fmb";
            if (cache && std::filesystem::exists(path)) return fmb200::loadIndex<Index>(path);
            auto [seqs, infos] = fmb200::io::loadQueries<Sigma>(ref, false, /*convertUnknownChar*/ true);
            if (seqs.empty()) throw std::runtime_error("no reference sequences in " + ref);
            auto ix = Index{seqs, rate, 1};
            if (cache) fmb200::saveIndex(ix, path);
            return ix;
        }();
        auto t1 = now();
        if (packed) {
            // the bulk path on 2-bit packed reads (io.hpp packs while parsing; fmb_search_and_locate_packed): mode "all" without a hit
            // limit, reads of one length -- what a read mapper feeds
            if (mode != "all" || maxhits) throw std::runtime_error("--packed serves --mode all without --maxhits");
            auto [pq, infos] = fmb200::io::loadQueriesPacked<Sigma>(query, reverse, true);
            size_t const L = pq.size() ? pq.length(0) : 0;
            for (size_t q = 0; q < pq.size(); ++q)
                if (pq.length(q) != L) throw std::runtime_error("--packed needs reads of one length");
            std::printf("index: %zu rows (%.2fs); loaded %zu queries (incl reverse complements), %zu packed words\n", index.size(), secs(t0, t1), pq.size(),
                        pq.words.size());
            auto t2 = now();
            auto [scheme, partition] = hamming ? fmb200::search_scheme::facadeScheme<false>(k, L) : fmb200::search_scheme::facadeScheme<true>(k, L);
            auto rows = fmb200::io::search_and_locate_bulk(index, pq, !hamming, &scheme, &partition);
            auto t3 = now();
            std::printf("k=%zu %s all (packed): %.3fs search+locate, %.0f q/s, %zu results\n", k, hamming ? "hamming" : "edit", secs(t2, t3),
                        pq.size() / std::max(secs(t2, t3), 1e-9), rows.size());
            if (!out.empty()) fmb200::io::saveResults(out, rows);
            return 0;
        }
        auto [queries, infos] = fmb200::io::loadQueries<Sigma>(query, reverse, true);
        std::printf("index: %zu rows (%.2fs); loaded %zu queries (incl reverse complements)\n", index.size(), secs(t0, t1), queries.size());
        std::vector<std::tuple<size_t, size_t, size_t, size_t>> results;
        auto report = [&](size_t qidx, size_t sid, size_t pos, size_t e) { results.emplace_back(qidx, sid, pos, e); };
        auto t2 = now();
        if (mode == "besthits") {
            // main.cpp:190-206: lowest error level at which a query has hits
            auto locate = [&](size_t qidx, auto const& cursor, size_t e) {
                for (auto [sid, spos, offset] : fmb200::LocateLinear{index, cursor}) report(qidx, sid, spos + offset, e);
            };
            std::vector<std::tuple<fmb200::search_scheme::Scheme, std::vector<size_t>>> schemes;
            size_t const L = queries.empty() ? 0 : queries[0].size();
            for (size_t e = 0; e <= k; ++e) {
                auto [s, p] = hamming ? fmb200::search_scheme::facadeScheme<false>(e, L) : fmb200::search_scheme::facadeScheme<true>(e, L);
                schemes.emplace_back(std::move(s), std::move(p));
            }
            size_t const n = maxhits ? maxhits : std::numeric_limits<size_t>::max();
            if (hamming) fmb200::search_ng26::search_best<false>(index, queries, schemes, locate, n);
            else fmb200::search_ng26::search_best<true>(index, queries, schemes, locate, n);
        } else {
            fmb200::Search{index, queries, !hamming, k, maxhits ? std::optional<size_t>{maxhits} : std::nullopt, report}();
        }
        auto t3 = now();
        std::printf("k=%zu %s %s: %.3fs search+locate, %.0f q/s, %zu results\n", k, hamming ? "hamming" : "edit", mode.c_str(), secs(t2, t3),
                    queries.size() / std::max(secs(t2, t3), 1e-9), results.size());
        if (!out.empty()) fmb200::io::saveResults(out, results);
    } catch (std::exception const& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
