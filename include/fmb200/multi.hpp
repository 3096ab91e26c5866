// fmb200/multi.hpp -- one index replica per GPU, queries sharded contiguously, results concatenated on the host.
// queries are independent, the index is replicated, there is NO collective on the search path;
// every GPU is driven by its own host thread.
#pragma once
#include <exception>
#include <thread>

#include "search.hpp"

namespace fmb200 {

// contiguous shard [begin, end) of `count` items for `rank` of `world` (sizes differ by at most one)
inline std::pair<size_t, size_t> shard_range(size_t count, size_t rank, size_t world) {
    size_t const base = count / world, rest = count % world;
    size_t const begin = rank * base + std::min(rank, rest);
    return {begin, begin + base + (rank < rest ? 1 : 0)};
}

template <size_t Sigma>
struct ReplicatedBiFMIndex {
    std::vector<BiFMIndex<Sigma>> replicas;     // replicas[g] lives on device g

    ReplicatedBiFMIndex(std::span<uint8_t const> bwt, std::span<uint8_t const> bwtRev, SparseArray const& sa, int n_devices = fmb_device_count()) {
        if (n_devices < 1) throw std::runtime_error("fmb200: no CUDA device (libfmb200 has no CPU fallback)");
        for (int g = 0; g < n_devices; ++g) replicas.emplace_back(bwt, bwtRev, sa, g);
    }
    size_t world() const { return replicas.size(); }

    // runs `fn(replica, shard_of_queries, first_qidx) -> std::vector<fmb_hit>` on every GPU and concatenates
    template <typename queries_t, typename Fn>
    std::vector<fmb_hit> run_sharded(queries_t const& queries, Fn&& fn) const {
        size_t const G = world(), Q = std::ranges::size(queries);
        std::vector<std::vector<fmb_hit>> parts(G);
        std::vector<std::exception_ptr> errs(G);
        std::vector<std::thread> threads;
        for (size_t g = 0; g < G; ++g) {
            threads.emplace_back([&, g] {
                try {
                    auto [b, e] = shard_range(Q, g, G);
                    auto shard = std::ranges::subrange(std::ranges::begin(queries) + b, std::ranges::begin(queries) + e);
                    parts[g] = fn(replicas[g], shard);
                    for (auto& h : parts[g]) h.qidx += b;
                } catch (...) { errs[g] = std::current_exception(); }
            });
        }
        for (auto& t : threads) t.join();
        for (auto& e : errs)
            if (e) std::rethrow_exception(e);
        std::vector<fmb_hit> all;
        for (auto& p : parts) all.insert(all.end(), p.begin(), p.end());
        return all;
    }

    template <Sequences queries_t>
    std::vector<fmb_hit> search_exact(queries_t const& queries) const {
        return run_sharded(queries, [](auto const& ix, auto const& shard) { return search_no_errors::search_bulk(ix, shard); });
    }
    template <bool Edit, Sequences queries_t>
    std::vector<fmb_hit> search_scheme(queries_t const& queries, search_scheme::Scheme const& scheme, std::vector<size_t> const& partition) const {
        return run_sharded(queries, [&](auto const& ix, auto const& shard) { return search_ng26::search_bulk<Edit>(ix, shard, scheme, partition); });
    }
};

}  // namespace fmb200
