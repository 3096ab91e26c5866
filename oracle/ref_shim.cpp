// ref_shim.cpp -- C interface over the reference's OWN headers (compiled from /root/reference by
// oracle/build_ref.sh into oracle/_ref/libfmref.so).  TEST INFRASTRUCTURE ONLY: it validates the C port in
// fm_oracle.c, generates golden fixtures and serves as the "reference" CPU baseline of bench.py.
// Nothing of the product links against it.
//
// Index type = fmc::BiFMIndex<Sigma, fmc::string::InterleavedBitvector16> (the example's choice,
// src/example/utils.h:262-265) or fmc::FMIndex<Sigma, ...> when no reverse BWT is given.
// Searches = search_no_errors::search (batched, search/SearchNoErrors.h:29), search_ng26::search
// (search/SearchNg26.h:427), search_backtracking::search (search/Backtracking.h:86), LocateLinear (locate.h:15).
// The reference search is single threaded; `threads` > 1 shards the queries into contiguous slices, one
// std::thread each, over the shared const index.
#include "fm_oracle.h"

#include <fmindex-collection/fmindex/BiFMIndex.h>
#include <fmindex-collection/fmindex/FMIndex.h>
#include <fmindex-collection/locate.h>
#include <fmindex-collection/search/Backtracking.h>
#include <fmindex-collection/search/SearchNg26.h>
#include <fmindex-collection/search/SearchNoErrors.h>
#include <fmindex-collection/search/SearchPseudo.h>
#include <fmindex-collection/search/search.h>
#include <fmindex-collection/search_scheme/expand.h>
#include <fmindex-collection/search_scheme/generator/all.h>
#include <fmindex-collection/string/InterleavedBitvector.h>

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <optional>
#include <ranges>
#include <span>
#include <thread>
#include <tuple>
#include <vector>

namespace {

using Entry  = std::tuple<uint32_t, uint32_t>;
using Sparse = fmc::suffixarray::SparseArray<Entry>;
template <size_t S> using Bi  = fmc::BiFMIndex<S, fmc::string::InterleavedBitvector16>;
template <size_t S> using Uni = fmc::FMIndex<S, fmc::string::InterleavedBitvector16>;

using Query   = std::span<uint8_t const>;
using Queries = std::vector<Query>;

struct Timer {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    double seconds() const { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};

}  // namespace

struct fmr_index {
    uint32_t sigma{};
    uint64_t n{};
    std::unique_ptr<Bi<5>> bi5;
    std::unique_ptr<Bi<21>> bi21;
    std::unique_ptr<Uni<5>> uni5;
    std::unique_ptr<Uni<21>> uni21;
};

namespace {

template <typename F>
auto dispatch(fmr_index const* ix, F&& f) {
    if (ix->bi5) return f(*ix->bi5);
    if (ix->bi21) return f(*ix->bi21);
    if (ix->uni5) return f(*ix->uni5);
    return f(*ix->uni21);
}
template <typename F>
auto dispatch_bi(fmr_index const* ix, F&& f) {
    if (ix->bi5) return f(*ix->bi5);
    if (ix->bi21) return f(*ix->bi21);
    std::abort();
}

Queries make_queries(uint8_t const* qsym, uint64_t const* qoff, uint64_t nq) {
    Queries qs(nq);
    for (uint64_t i = 0; i < nq; ++i) qs[i] = Query{qsym + qoff[i], qoff[i + 1] - qoff[i]};
    return qs;
}

// run f(slice_begin, slice_queries, out_vector) on `threads` contiguous shards, concatenate in shard order
template <typename T, typename F>
uint64_t sharded(Queries const& qs, int threads, T** out, double* seconds, F&& f) {
    if (threads < 1) threads = 1;
    std::vector<std::vector<T>> parts(threads);
    Timer t;
    if (threads == 1) {
        f(uint64_t{0}, qs, parts[0]);
    } else {
        std::vector<std::thread> pool;
        uint64_t per = (qs.size() + threads - 1) / threads;
        for (int w = 0; w < threads; ++w) {
            pool.emplace_back([&, w] {
                uint64_t b = std::min<uint64_t>(qs.size(), per * w), e = std::min<uint64_t>(qs.size(), b + per);
                Queries sub(qs.begin() + b, qs.begin() + e);
                f(b, sub, parts[w]);
            });
        }
        for (auto& th : pool) th.join();
    }
    if (seconds) *seconds = t.seconds();
    uint64_t total = 0;
    for (auto& p : parts) total += p.size();
    T* v = static_cast<T*>(std::malloc((total ? total : 1) * sizeof(T)));
    uint64_t k = 0;
    for (auto& p : parts) { std::memcpy(v + k, p.data(), p.size() * sizeof(T)); k += p.size(); }
    *out = v;
    return total;
}

Sparse make_sparse(uint64_t n, uint64_t const* bitmap, uint32_t const* seq, uint32_t const* pos) {
    // SparseArray's range constructor (suffixarray/SparseArray.h:44-58) walks the range twice, in order;
    // a counter-carrying lazy view would be wrong on the second pass, so precompute the sample index per word.
    std::vector<uint64_t> before(n / 64 + 2, 0);
    for (uint64_t w = 0; w + 1 < before.size(); ++w)
        before[w + 1] = before[w] + (w < (n + 63) / 64 ? (uint64_t)__builtin_popcountll(bitmap[w]) : 0);
    auto view = std::views::iota(uint64_t{0}, n) | std::views::transform([&](uint64_t i) -> std::optional<Entry> {
        uint64_t w = bitmap[i >> 6];
        if (!((w >> (i & 63)) & 1)) return std::nullopt;
        uint64_t r = before[i >> 6] + (uint64_t)__builtin_popcountll(w & ((uint64_t{1} << (i & 63)) - 1));
        return Entry{seq[r], pos[r]};
    });
    return Sparse{view};
}

}  // namespace

extern "C" {

fmr_index* fmr_index_from_bwt(uint32_t sigma, uint64_t n, uint8_t const* bwt, uint8_t const* bwt_rev,
                              uint64_t const* sample_bitmap, uint32_t const* sample_seq, uint32_t const* sample_pos,
                              uint64_t /*n_samples*/) {
    if (sigma != 5 && sigma != 21) return nullptr;
    auto ix = new fmr_index{};
    ix->sigma = sigma;
    ix->n = n;
    auto sparse = make_sparse(n, sample_bitmap, sample_seq, sample_pos);
    std::span<uint8_t const> b{bwt, n};
    if (bwt_rev) {
        std::span<uint8_t const> br{bwt_rev, n};
        if (sigma == 5) ix->bi5 = std::make_unique<Bi<5>>(b, br, std::move(sparse));        // BiFMIndex.h:40-51
        else ix->bi21 = std::make_unique<Bi<21>>(b, br, std::move(sparse));
    } else {
        if (sigma == 5) ix->uni5 = std::make_unique<Uni<5>>(b, std::move(sparse));           // FMIndex.h:28-32
        else ix->uni21 = std::make_unique<Uni<21>>(b, std::move(sparse));
    }
    return ix;
}

// Full construction through the reference's own constructor BiFMIndex(Sequences, samplingRate, threads)
// (BiFMIndex.h:107) / FMIndex(Sequences, ...) (FMIndex.h:58).  `text` = s0 0 s1 0 ... ; it is split at the
// delimiters back into sequences first.  Suffix sorting goes through the libsais stand-in (oracle/shim).
fmr_index* fmr_index_build(uint8_t const* text, uint64_t n, uint32_t sigma, uint32_t rate, int bidirectional) {
    if (sigma != 5 && sigma != 21) return nullptr;
    std::vector<std::vector<uint8_t>> seqs;
    std::vector<uint8_t> cur;
    for (uint64_t i = 0; i < n; ++i) {
        if (text[i] == 0) { seqs.push_back(cur); cur.clear(); }
        else cur.push_back(text[i]);
    }
    if (!cur.empty()) return nullptr;   // text must end with a delimiter
    auto ix = new fmr_index{};
    ix->sigma = sigma;
    ix->n = n;
    if (bidirectional) {
        if (sigma == 5) ix->bi5 = std::make_unique<Bi<5>>(seqs, rate, 1);
        else ix->bi21 = std::make_unique<Bi<21>>(seqs, rate, 1);
    } else {
        if (sigma == 5) ix->uni5 = std::make_unique<Uni<5>>(seqs, rate, 1);
        else ix->uni21 = std::make_unique<Uni<21>>(seqs, rate, 1);
    }
    return ix;
}

void fmr_index_free(fmr_index* ix) { delete ix; }

uint64_t fmr_size(fmr_index const* ix) { return dispatch(ix, [](auto const& i) -> uint64_t { return i.size(); }); }
uint64_t fmr_C(fmr_index const* ix, uint32_t s) { return dispatch(ix, [&](auto const& i) -> uint64_t { return i.C[s]; }); }

// String_c accessors (string/concepts.h:26-87); dir 1 = bwtRev (bidirectional only)
uint64_t fmr_symbol(fmr_index const* ix, int dir, uint64_t idx) {
    if (dir) return dispatch_bi(ix, [&](auto const& i) -> uint64_t { return i.bwtRev.symbol(idx); });
    return dispatch(ix, [&](auto const& i) -> uint64_t { return i.bwt.symbol(idx); });
}
uint64_t fmr_rank(fmr_index const* ix, int dir, uint64_t idx, uint64_t s) {
    if (dir) return dispatch_bi(ix, [&](auto const& i) -> uint64_t { return i.bwtRev.rank(idx, s); });
    return dispatch(ix, [&](auto const& i) -> uint64_t { return i.bwt.rank(idx, s); });
}
uint64_t fmr_prefix_rank(fmr_index const* ix, int dir, uint64_t idx, uint64_t s) {
    if (dir) return dispatch_bi(ix, [&](auto const& i) -> uint64_t { return i.bwtRev.prefix_rank(idx, s); });
    return dispatch(ix, [&](auto const& i) -> uint64_t { return i.bwt.prefix_rank(idx, s); });
}
void fmr_all_ranks_and_prefix_ranks(fmr_index const* ix, int dir, uint64_t idx, uint64_t* rs, uint64_t* prs) {
    auto f = [&](auto const& str) {
        auto [a, b] = str.all_ranks_and_prefix_ranks(idx);
        for (size_t k = 0; k < a.size(); ++k) { rs[k] = a[k]; prs[k] = b[k]; }
    };
    if (dir) dispatch_bi(ix, [&](auto const& i) { f(i.bwtRev); });
    else dispatch(ix, [&](auto const& i) { f(i.bwt); });
}

// BiFMIndexCursor::extendLeft/Right(symb) (fmindex/BiFMIndexCursor.h:113-128); cur = {lb, lbRev, len, steps}
void fmr_extend(fmr_index const* ix, int right, uint64_t const cur[4], uint64_t symb, uint64_t out[4]) {
    dispatch_bi(ix, [&](auto const& i) {
        using Index = std::decay_t<decltype(i)>;
        auto c = fmc::BiFMIndexCursor<Index>{i, cur[0], cur[1], cur[2], cur[3]};
        auto r = right ? c.extendRight(symb) : c.extendLeft(symb);
        out[0] = r.lb; out[1] = r.lbRev; out[2] = r.len; out[3] = r.steps;
    });
}

// index.locate(row) -> (seq, pos, steps)  (BiFMIndex.h:177-202, FMIndex.h:114-124)
void fmr_locate_row(fmr_index const* ix, uint64_t row, uint64_t out[3]) {
    dispatch(ix, [&](auto const& i) {
        auto [seq, pos, off] = i.locate(row);
        out[0] = seq; out[1] = pos; out[2] = off;
    });
}

uint64_t fmr_search_exact(fmr_index const* ix, uint8_t const* qsym, uint64_t const* qoff, uint64_t nq,
                          fmo_hit** out, int threads, double* seconds) {
    auto qs = make_queries(qsym, qoff, nq);
    return dispatch(ix, [&](auto const& index) {
        return sharded<fmo_hit>(qs, threads, out, seconds, [&](uint64_t base, Queries const& sub, std::vector<fmo_hit>& res) {
            if constexpr (requires { index.bwtRev; }) {
                fmc::search_no_errors::search(index, sub, [&](size_t qidx, auto cur) {       // SearchNoErrors.h:29
                    res.push_back(fmo_hit{base + qidx, cur.lb, 0, cur.len, cur.steps, 0});
                });
            } else {
                // FMIndexCursor has no `steps` (FMIndexCursor.h:17-19): only the single-query form compiles
                for (size_t qidx = 0; qidx < sub.size(); ++qidx) {
                    auto cur = fmc::search_no_errors::search(index, sub[qidx]);               // SearchNoErrors.h:13
                    if (!cur.empty()) res.push_back(fmo_hit{base + qidx, cur.lb, 0, cur.len, sub[qidx].size(), 0});
                }
            }
        });
    });
}

uint64_t fmr_search_ng26(fmr_index const* ix, uint8_t const* qsym, uint64_t const* qoff, uint64_t nq, int edit,
                         uint32_t n_searches, uint32_t n_parts, uint32_t const* pi, uint32_t const* l, uint32_t const* u,
                         uint32_t const* partition, uint64_t max_hits, fmo_hit** out, int threads, double* seconds) {
    auto qs = make_queries(qsym, qoff, nq);
    fmc::search_scheme::Scheme scheme;
    for (uint32_t s = 0; s < n_searches; ++s) {
        fmc::search_scheme::Search se;
        for (uint32_t p = 0; p < n_parts; ++p) {
            se.pi.push_back(pi[s * n_parts + p]);
            se.l.push_back(l[s * n_parts + p]);
            se.u.push_back(u[s * n_parts + p]);
        }
        scheme.push_back(se);
    }
    std::vector<size_t> part(partition, partition + n_parts);
    size_t n = max_hits == UINT64_MAX ? std::numeric_limits<size_t>::max() : max_hits;
    return dispatch_bi(ix, [&](auto const& index) {
        return sharded<fmo_hit>(qs, threads, out, seconds, [&](uint64_t base, Queries const& sub, std::vector<fmo_hit>& res) {
            auto cb = [&](size_t qidx, auto cur, size_t e) {
                res.push_back(fmo_hit{base + qidx, cur.lb, cur.lbRev, cur.len, cur.steps, e});
            };
            if (edit) fmc::search_ng26::search<true>(index, sub, scheme, part, cb, n);           // SearchNg26.h:427
            else fmc::search_ng26::search<false>(index, sub, scheme, part, cb, n);
        });
    });
}

// fmc::search<Edit>(index, queries, errors, cb) -- the facade (search/search.h:26): k==0 -> no_errors, else ng26+h2
uint64_t fmr_search_facade(fmr_index const* ix, uint8_t const* qsym, uint64_t const* qoff, uint64_t nq, int edit,
                           uint32_t errors, fmo_hit** out, int threads, double* seconds) {
    auto qs = make_queries(qsym, qoff, nq);
    return dispatch_bi(ix, [&](auto const& index) {
        return sharded<fmo_hit>(qs, threads, out, seconds, [&](uint64_t base, Queries const& sub, std::vector<fmo_hit>& res) {
            auto cb = [&](size_t qidx, auto cur, size_t e) {
                uint64_t lbRev = 0;
                if constexpr (requires { cur.lbRev; }) lbRev = cur.lbRev;
                res.push_back(fmo_hit{base + qidx, cur.lb, lbRev, cur.len, cur.steps, e});
            };
            if (edit) fmc::search<true>(index, sub, errors, cb);
            else fmc::search<false>(index, sub, errors, cb);
        });
    });
}

// fmc::search_pseudo::search<Edit>(index, queries, expandedScheme, cb) (search/SearchPseudo.h:171-186); pi / l / u = n_searches x L
uint64_t fmr_search_pseudo(fmr_index const* ix, uint8_t const* qsym, uint64_t const* qoff, uint64_t nq, int edit,
                           uint32_t n_searches, uint32_t L, uint32_t const* pi, uint32_t const* l, uint32_t const* u, fmo_hit** out, int threads, double* seconds) {
    auto qs = make_queries(qsym, qoff, nq);
    fmc::search_scheme::Scheme scheme;
    for (uint32_t s = 0; s < n_searches; ++s) {
        fmc::search_scheme::Search se;
        for (uint32_t p = 0; p < L; ++p) {
            se.pi.push_back(pi[s * L + p]);
            se.l.push_back(l[s * L + p]);
            se.u.push_back(u[s * L + p]);
        }
        scheme.push_back(se);
    }
    return dispatch_bi(ix, [&](auto const& index) {
        return sharded<fmo_hit>(qs, threads, out, seconds, [&](uint64_t base, Queries const& sub, std::vector<fmo_hit>& res) {
            auto cb = [&](size_t qidx, auto cur, size_t e) { res.push_back(fmo_hit{base + qidx, cur.lb, cur.lbRev, cur.len, cur.steps, e}); };
            if (edit) fmc::search_pseudo::search<true>(index, sub, scheme, cb);
            else fmc::search_pseudo::search<false>(index, sub, scheme, cb);
        });
    });
}

uint64_t fmr_search_backtracking(fmr_index const* ix, uint8_t const* qsym, uint64_t const* qoff, uint64_t nq,
                                 uint32_t max_errors, fmo_hit** out, int threads, double* seconds) {
    auto qs = make_queries(qsym, qoff, nq);
    return dispatch(ix, [&](auto const& index) {
        return sharded<fmo_hit>(qs, threads, out, seconds, [&](uint64_t base, Queries const& sub, std::vector<fmo_hit>& res) {
            fmc::search_backtracking::search(index, sub, max_errors, [&](size_t qidx, auto cur, size_t e) {   // Backtracking.h:86
                uint64_t lbRev = 0, steps = 0;
                if constexpr (requires { cur.lbRev; }) lbRev = cur.lbRev;
                if constexpr (requires { cur.steps; }) steps = cur.steps;
                res.push_back(fmo_hit{base + qidx, cur.lb, lbRev, cur.len, steps, e});
            });
        });
    });
}

// LocateLinear over all rows of all hits (locate.h:15-57), reporting (qidx, seq, pos+offset, e) like
// fmc::Search::operator() (search/search.h:55-60).  Hits are sharded over threads.
uint64_t fmr_locate(fmr_index const* ix, fmo_hit const* hits, uint64_t nhits, fmo_loc** out, int threads, double* seconds) {
    if (threads < 1) threads = 1;
    std::vector<uint64_t> start(nhits + 1, 0);
    for (uint64_t i = 0; i < nhits; ++i) start[i + 1] = start[i] + hits[i].len;
    uint64_t total = start[nhits];
    fmo_loc* v = static_cast<fmo_loc*>(std::malloc((total ? total : 1) * sizeof(fmo_loc)));
    Timer t;
    dispatch(ix, [&](auto const& index) {
        using Index = std::decay_t<decltype(index)>;
        auto work = [&](uint64_t b, uint64_t e) {
            for (uint64_t i = b; i < e; ++i) {
                uint64_t k = start[i];
                if constexpr (requires { index.bwtRev; }) {
                    auto cur = fmc::BiFMIndexCursor<Index>{index, hits[i].lb, hits[i].lb_rev, hits[i].len, hits[i].steps};
                    for (auto [seq, pos, off] : fmc::LocateLinear{index, cur})
                        v[k++] = fmo_loc{hits[i].qidx, seq, pos + off, hits[i].e};
                } else {
                    auto cur = fmc::FMIndexCursor<Index>{index, hits[i].lb, hits[i].len};
                    for (auto [seq, pos, off] : fmc::LocateLinear{index, cur})
                        v[k++] = fmo_loc{hits[i].qidx, seq, pos + off, hits[i].e};
                }
            }
        };
        if (threads == 1) { work(0, nhits); return; }
        std::vector<std::thread> pool;
        uint64_t per = (nhits + threads - 1) / threads;
        for (int w = 0; w < threads; ++w) {
            uint64_t b = std::min<uint64_t>(nhits, per * w), e = std::min<uint64_t>(nhits, b + per);
            pool.emplace_back(work, b, e);
        }
        for (auto& th : pool) th.join();
    });
    if (seconds) *seconds = t.seconds();
    *out = v;
    return total;
}

// search_scheme::generator::all[name](minK, K, 0, 0) (search_scheme/generator/all.h:35-118) -> flat arrays.
// Returns the number of searches (0 = unknown generator or buffers too small); *n_parts receives |pi|.
uint32_t fmr_scheme_generate(char const* name, uint32_t min_k, uint32_t max_k, uint32_t cap, uint32_t* n_parts,
                             uint32_t* pi, uint32_t* l, uint32_t* u) {
    auto it = fmc::search_scheme::generator::all.find(name);
    if (it == fmc::search_scheme::generator::all.end()) return 0;
    auto scheme = it->second.generator(min_k, max_k, 0, 0);
    if (scheme.empty()) return 0;
    uint32_t np = scheme[0].pi.size();
    if (scheme.size() * np > cap) return 0;
    *n_parts = np;
    for (size_t s = 0; s < scheme.size(); ++s)
        for (uint32_t p = 0; p < np; ++p) {
            pi[s * np + p] = scheme[s].pi[p];
            l[s * np + p] = scheme[s].l[p];
            u[s * np + p] = scheme[s].u[p];
        }
    return scheme.size();
}

// createUniformPartition(parts, totalSum)  (search_scheme/expand.h:324-336)
void fmr_uniform_partition(uint32_t parts, uint32_t total, uint32_t* out) {
    auto p = fmc::search_scheme::createUniformPartition(parts, total);
    for (uint32_t i = 0; i < parts; ++i) out[i] = p[i];
}

void fmr_free(void* p) { std::free(p); }

}  // extern "C"
