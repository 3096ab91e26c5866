// fmb200/io.hpp -- the I/O shell around the search path: FASTA in, "queryId seqId pos" out.
//
// Mirrors the helpers of the reference's example program (paths relative to /root/reference/src/example/):
//   utils.h:18-25     struct Query{name, reverse}
//   utils.h:26-105    loadQueries<Sigma>(path, reverse, convertUnknownChar) -> (queries, queryInfos): FASTA records mapped to
//                     ranks ($ -> 0, A C G T -> 1 2 3 4, N -> 5 when Sigma == 6, anything else -> 5 / 1 with convertUnknownChar,
//                     else an error); with `reverse` every read is followed by its reverse complement
//   main.cpp:260-266  the result file: one line "queryId seqId pos" per located row
// Differences: sequences may span several lines and the last record does not need a trailing newline (the example drops the
// final byte of the file unseen, utils.h:63,73); a '>' is only a header at the start of a line.  The reverse-complement doubling
// can also be left to the device (fmb_queries_upload_revcomp): see uploadQueries below.  loadQueriesPacked / parseFastaPacked produce the
// 2-bit packed batch of fmb_queries_upload_packed while parsing (a quarter of the PCIe bytes; DNA indices only).
#pragma once
#include <algorithm>
#include <cstdio>
#include <filesystem>
#include <fstream>
#include <string>
#include <tuple>
#include <vector>

#include "search.hpp"

namespace fmb200::io {

struct Query {
    std::string name;
    bool reverse{};
    bool operator==(Query const&) const = default;
};

// example/utils.h:64-70: reverse the read and swap 1 <-> 4, 2 <-> 3 (other ranks stay)
inline constexpr uint8_t kDnaComplement[6] = {0, 4, 3, 2, 1, 5};
inline std::vector<uint8_t> reverseComplement(std::vector<uint8_t> const& q) {
    std::vector<uint8_t> r(q.rbegin(), q.rend());
    for (auto& c : r)
        if (c < 6) c = kDnaComplement[c];
    return r;
}

template <size_t Sigma>
uint8_t rankOf(char ch, bool convertUnknownChar) {
    switch (ch) {
    case '$': return 0;
    case 'A': case 'a': return 1;
    case 'C': case 'c': return 2;
    case 'G': case 'g': return 3;
    case 'T': case 't': return 4;
    default: break;
    }
    if ((ch == 'N' || ch == 'n') && Sigma == 6) return 5;
    if (convertUnknownChar) return Sigma == 6 ? 5 : 1;
    throw std::runtime_error("unknown alphabet");
}

// FASTA text -> (sequences as ranks, infos)
template <size_t Sigma>
auto parseFasta(std::string_view text, bool reverse, bool convertUnknownChar) {
    std::vector<std::vector<uint8_t>> queries;
    std::vector<Query> infos;
    if (text.empty()) return std::make_tuple(queries, infos);
    if (text[0] != '>') throw std::runtime_error("can't read fasta file");
    std::vector<uint8_t> cur;
    bool open = false;
    auto flush = [&] {
        if (!open) return;
        queries.push_back(cur);
        if (reverse) queries.push_back(reverseComplement(cur));
        cur.clear();
    };
    size_t i = 0;
    while (i < text.size()) {
        size_t eol = text.find('\n', i);
        if (eol == std::string_view::npos) eol = text.size();
        std::string_view line = text.substr(i, eol - i);
        if (!line.empty() && line.back() == '\r') line.remove_suffix(1);
        if (!line.empty() && line[0] == '>') {
            flush();
            open = true;
            line.remove_prefix(1);
            if (!line.empty() && line[0] == ' ') line.remove_prefix(1);
            infos.push_back(Query{std::string{line}, false});
            if (reverse) infos.push_back(Query{std::string{line}, true});
        } else {
            for (char ch : line) cur.push_back(rankOf<Sigma>(ch, convertUnknownChar));
        }
        i = eol + 1;
    }
    flush();
    return std::make_tuple(queries, infos);
}

// example/utils.h:26-105; a missing file yields empty vectors like the example
template <size_t Sigma>
auto loadQueries(std::string const& path, bool reverse, bool convertUnknownChar) {
    if (path.empty() || !std::filesystem::exists(path)) return std::make_tuple(std::vector<std::vector<uint8_t>>{}, std::vector<Query>{});
    std::ifstream ifs(path, std::ios::binary);
    std::string text((std::istreambuf_iterator<char>(ifs)), std::istreambuf_iterator<char>());
    return parseFasta<Sigma>(text, reverse, convertUnknownChar);
}

// 2-bit packed host batch (sigma <= 5): the form fmb_queries_upload_packed / fmb_search_and_locate_packed take.  A quarter of the
// bytes of the flattened `Sequences` crosses PCIe.  Symbol i is the field [2i, 2i + 2) of `words` (rank - 1); ranks without 2-bit
// code (0, and 5 = N) are listed in exc_pos / exc_sym.
struct PackedQueries {
    std::vector<uint32_t> words;
    std::vector<uint64_t> offsets{0};
    std::vector<uint64_t> exc_pos;
    std::vector<uint8_t> exc_sym;

    size_t size() const { return offsets.size() - 1; }
    uint64_t symbols() const { return total_; }
    size_t length(size_t q) const { return static_cast<size_t>(offsets[q + 1] - offsets[q]); }
    uint8_t symbol(uint64_t pos) const {        // for tests and small inputs: binary search in the exception list
        auto it = std::lower_bound(exc_pos.begin(), exc_pos.end(), pos);
        if (it != exc_pos.end() && *it == pos) return exc_sym[static_cast<size_t>(it - exc_pos.begin())];
        return static_cast<uint8_t>(((words[pos >> 4] >> (2 * (pos & 15))) & 3) + 1);
    }
    void push(uint8_t rank) {
        if ((total_ & 15) == 0) words.push_back(0);
        words.back() |= ((uint32_t(rank) - 1u) & 3u) << (2 * (total_ & 15));       // the field of an exception is ignored by the device
        if (rank < 1 || rank > 4) {
            exc_pos.push_back(total_);
            exc_sym.push_back(rank);
        }
        ++total_;
    }
    void endQuery() { offsets.push_back(total_); }
    // the device reads whole words past the last symbol: one word of slack (fmb_pack_symbols' contract)
    uint32_t const* data() {
        size_t const need = static_cast<size_t>((total_ + 15) / 16 + 1);
        if (words.size() < need) words.resize(need, 0);
        return words.data();
    }

private:
    uint64_t total_{};
};

// parseFasta that packs while parsing: no byte-per-symbol copy of the batch ever exists on the host
template <size_t Sigma>
auto parseFastaPacked(std::string_view text, bool reverse, bool convertUnknownChar) {
    PackedQueries packed;
    std::vector<Query> infos;
    if (text.empty()) return std::make_tuple(std::move(packed), infos);
    if (text[0] != '>') throw std::runtime_error("can't read fasta file");
    std::vector<uint8_t> cur;                   // the record being read (needed for its reverse complement)
    bool open = false;
    auto flush = [&] {
        if (!open) return;
        for (auto c : cur) packed.push(c);
        packed.endQuery();
        if (reverse) {
            for (size_t i = cur.size(); i-- > 0;) packed.push(cur[i] < 6 ? kDnaComplement[cur[i]] : cur[i]);
            packed.endQuery();
        }
        cur.clear();
    };
    size_t i = 0;
    while (i < text.size()) {
        size_t eol = text.find('\n', i);
        if (eol == std::string_view::npos) eol = text.size();
        std::string_view line = text.substr(i, eol - i);
        if (!line.empty() && line.back() == '\r') line.remove_suffix(1);
        if (!line.empty() && line[0] == '>') {
            flush();
            open = true;
            line.remove_prefix(1);
            if (!line.empty() && line[0] == ' ') line.remove_prefix(1);
            infos.push_back(Query{std::string{line}, false});
            if (reverse) infos.push_back(Query{std::string{line}, true});
        } else {
            for (char ch : line) cur.push_back(rankOf<Sigma>(ch, convertUnknownChar));
        }
        i = eol + 1;
    }
    flush();
    return std::make_tuple(std::move(packed), infos);
}

template <size_t Sigma>
auto loadQueriesPacked(std::string const& path, bool reverse, bool convertUnknownChar) {
    if (path.empty() || !std::filesystem::exists(path)) return std::make_tuple(PackedQueries{}, std::vector<Query>{});
    std::ifstream ifs(path, std::ios::binary);
    std::string text((std::istreambuf_iterator<char>(ifs)), std::istreambuf_iterator<char>());
    return parseFastaPacked<Sigma>(text, reverse, convertUnknownChar);
}

// device batch of a packed host batch (fmb_queries_upload_packed)
template <typename index_t>
auto uploadQueries(index_t const& index, PackedQueries& packed) {
    fmb_queries* q{};
    check(fmb_queries_upload_packed(&q, index.handle(), packed.data(), packed.offsets.data(), packed.size(), packed.exc_pos.data(), packed.exc_sym.data(),
                                    packed.exc_pos.size()));
    return detail::QueriesHandle{q};
}

// search_and_locate_bulk on a packed host batch (fmb_search_and_locate_packed); scheme == nullptr selects exact search
template <typename index_t, detail::SchemeLike scheme_t = search_scheme::Scheme>
std::vector<fmb_loc32> search_and_locate_bulk(index_t const& index, PackedQueries& packed, bool edit = false, scheme_t const* scheme = nullptr,
                                              std::vector<size_t> const* partition = nullptr, fmb_stats* stats = nullptr) {
    detail::FlatScheme fs;
    if (scheme) fs = detail::flatten(*scheme, *partition);
    std::vector<fmb_loc32> out(std::max<size_t>(packed.size() * 2, 1024));
    for (;;) {
        uint64_t n_out = 0;
        int rc = fmb_search_and_locate_packed(index.handle(), packed.data(), packed.offsets.data(), packed.size(), packed.exc_pos.data(), packed.exc_sym.data(),
                                              packed.exc_pos.size(), edit ? 1 : 0, fs.n_searches, fs.n_parts, fs.pi.data(), fs.l.data(), fs.u.data(),
                                              fs.partition.data(), out.data(), out.size(), &n_out, stats);
        if (rc == FMB_EOVERFLOW && n_out > out.size()) {
            out.resize(n_out);
            continue;
        }
        check(rc);
        out.resize(n_out);
        return out;
    }
}

// Device batch of `queries`; with reverse == true it holds every read followed by its reverse complement (query ids 2i, 2i+1 as
// loadQueries(…, reverse = true) numbers them), the complements being generated on the device.
template <typename index_t, Sequences queries_t>
auto uploadQueries(index_t const& index, queries_t const& queries, bool reverse) {
    auto flat = flatten(queries);
    fmb_queries* q{};
    if (reverse) {
        std::array<uint8_t, 32> comp{};
        for (size_t c = 0; c < comp.size(); ++c) comp[c] = c < 6 ? kDnaComplement[c] : static_cast<uint8_t>(c);
        check(fmb_queries_upload_revcomp(&q, index.handle(), flat.symbols.data(), flat.offsets.data(), flat.size(), comp.data()));
    } else {
        check(fmb_queries_upload(&q, index.handle(), flat.symbols.data(), flat.offsets.data(), flat.size()));
    }
    return detail::QueriesHandle{q};
}

// example/main.cpp:260-266
template <typename results_t>
void saveResults(std::filesystem::path const& path, results_t const& results) {
    auto* ofs = std::fopen(path.c_str(), "w");
    if (!ofs) throw std::runtime_error("cannot write " + path.string());
    for (auto const& r : results) {
        if constexpr (requires { r.qidx; r.seq; r.pos; }) {
            std::fprintf(ofs, "%llu %llu %llu\n", (unsigned long long)r.qidx, (unsigned long long)r.seq, (unsigned long long)r.pos);
        } else {
            auto const& [queryId, seqId, pos, e] = r;
            (void)e;
            std::fprintf(ofs, "%llu %llu %llu\n", (unsigned long long)queryId, (unsigned long long)seqId, (unsigned long long)pos);
        }
    }
    std::fclose(ofs);
}

}  // namespace fmb200::io
