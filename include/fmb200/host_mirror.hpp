// fmb200/host_mirror.hpp -- the device occurrence-table layouts as a host-side String_c (seam 2).
//
// fmb200::HostMirror<Sigma> satisfies the reference's String_c concept (string/concepts.h:26-87: constructible from a span of
// symbols, size / symbol / rank / prefix_rank / all_ranks / all_ranks_and_prefix_ranks), so it plugs into the reference's own index
// types -- fmc::BiFMIndex<Sigma, fmb200::HostMirror>, fmc::FMIndex<Sigma, fmb200::HostMirror> -- and every algorithm of the
// reference then runs, on the CPU, over exactly the blocks the GPU kernels read:
//   Sigma <= 5   one 32-byte block per 64 rows: u32 cnt[4] (absolute counts of symbols 1..4) + two bit planes of (symbol - 1);
//                delimiter rows are coded like symbol 1 and kept in a sorted side list          (csrc/fmb_device.cuh OccDna)
//   Sigma  > 5   per 64 rows ceil(log2 Sigma) bit planes + Sigma - 1 exclusive prefix counts, padded to 32 bytes  (OccGen)
// blocks() returns the raw bytes; fmb_index_export_blocks returns the device's: the tests compare them byte for byte.
// This is a correctness harness and an API-surface statement ("the OccTable/String concept with rank, prefix_rank and all_ranks"),
// not a product path: nothing in libfmb200.so uses it.
#pragma once
#include <algorithm>
#include <array>
#include <bit>
#include <cstdint>
#include <cstring>
#include <span>
#include <stdexcept>
#include <tuple>
#include <vector>

namespace fmb200 {

template <size_t TSigma>
struct HostMirror {
    static constexpr size_t Sigma = TSigma;
    static constexpr bool kDna = TSigma <= 5;

    HostMirror() = default;
    HostMirror(HostMirror&&) noexcept = default;
    HostMirror(HostMirror const&) = default;
    auto operator=(HostMirror&&) noexcept -> HostMirror& = default;
    auto operator=(HostMirror const&) -> HostMirror& = default;

    HostMirror(std::span<uint8_t const> symbols) : n_{symbols.size()} {
        if (n_ >= 0xFFFFFFFFull - 64) throw std::runtime_error("fmb200::HostMirror: n < 2^32 - 64 rows (32-bit counters, like the device image)");
        size_t const nblocks = n_ / 64 + 1;
        if constexpr (kDna) {
            stride_ = 32;
            bytes_.assign(nblocks * 32, 0);
            uint32_t acc[4] = {0, 0, 0, 0};
            for (size_t b = 0; b < nblocks; ++b) {
                uint64_t p0 = 0, p1 = 0;
                uint32_t cnt[4] = {acc[0], acc[1], acc[2], acc[3]};
                for (size_t r = 0; r < 64 && b * 64 + r < n_; ++r) {
                    uint32_t const s = symbols[b * 64 + r];
                    if (s >= Sigma) throw std::runtime_error("fmb200::HostMirror: symbol >= Sigma");
                    uint32_t const k = s ? s - 1 : 0;
                    if (s == 0) delim_rows_.push_back(static_cast<uint32_t>(b * 64 + r));
                    else acc[k] += 1;
                    p0 |= uint64_t(k & 1) << r;
                    p1 |= uint64_t((k >> 1) & 1) << r;
                }
                uint8_t* out = bytes_.data() + b * 32;
                std::memcpy(out, cnt, 16);
                std::memcpy(out + 16, &p0, 8);
                std::memcpy(out + 24, &p1, 8);
            }
            delim_rows_.push_back(0xFFFFFFFFu);
        } else {
            planes_ = 1;
            while ((size_t{1} << planes_) < Sigma) ++planes_;
            stride_ = (8 * planes_ + 4 * (static_cast<uint32_t>(Sigma) - 1) + 31) / 32 * 32;
            bytes_.assign(nblocks * stride_, 0);
            std::vector<uint32_t> acc(Sigma, 0);          // acc[s] = # symbols == s before the block
            for (size_t b = 0; b < nblocks; ++b) {
                uint8_t* out = bytes_.data() + b * stride_;
                uint32_t run = 0;                         // pc[s] = # symbols <= s = # symbols < s + 1 before the block
                for (size_t s = 0; s + 1 < Sigma; ++s) {
                    run += acc[s];
                    std::memcpy(out + 8 * planes_ + 4 * s, &run, 4);
                }
                uint64_t pl[8] = {};
                for (size_t r = 0; r < 64 && b * 64 + r < n_; ++r) {
                    uint32_t const s = symbols[b * 64 + r];
                    if (s >= Sigma) throw std::runtime_error("fmb200::HostMirror: symbol >= Sigma");
                    acc[s] += 1;
                    for (uint32_t j = 0; j < planes_; ++j) pl[j] |= uint64_t((s >> j) & 1) << r;
                }
                std::memcpy(out, pl, 8 * planes_);
            }
        }
    }

    size_t size() const { return n_; }
    std::span<uint8_t const> blocks() const { return bytes_; }
    size_t block_bytes() const { return stride_; }

    uint8_t symbol(size_t idx) const {
        uint32_t const off = idx & 63;
        if constexpr (kDna) {
            auto const [p0, p1] = planes2(idx >> 6);
            uint32_t const k = uint32_t((p0 >> off) & 1) | (uint32_t((p1 >> off) & 1) << 1);
            if (k == 0 && std::binary_search(delim_rows_.begin(), delim_rows_.end() - 1, static_cast<uint32_t>(idx))) return 0;
            return static_cast<uint8_t>(k + 1);
        } else {
            uint32_t s = 0;
            for (uint32_t j = 0; j < planes_; ++j) s |= uint32_t((plane(idx >> 6, j) >> off) & 1) << j;
            return static_cast<uint8_t>(s);
        }
    }

    uint64_t rank(size_t idx, uint8_t symb) const {
        size_t const blk = idx >> 6;
        uint64_t const low = low_mask(idx & 63);
        if constexpr (kDna) {
            if (symb == 0) return delims_below(idx);
            uint32_t const k = symb - 1u;
            auto const [p0, p1] = planes2(blk);
            uint64_t const m = ((k & 1) ? p0 : ~p0) & ((k & 2) ? p1 : ~p1);
            uint64_t r = cnt(blk, k) + std::popcount(m & low);
            if (k == 0) r -= delims_below(idx) - delims_below(blk * 64);       // delimiter rows are coded like symbol 1
            return r;
        } else {
            uint64_t m = ~uint64_t{0};
            for (uint32_t j = 0; j < planes_; ++j) m &= ((symb >> j) & 1) ? plane(blk, j) : ~plane(blk, j);
            return (pc_hi(blk, symb) - pc_lo(blk, symb)) + std::popcount(m & low);
        }
    }

    // # rows < idx holding a symbol < symb (exclusive, like every String_c of the reference)
    uint64_t prefix_rank(size_t idx, uint8_t symb) const {
        size_t const blk = idx >> 6;
        uint64_t const low = low_mask(idx & 63);
        if constexpr (kDna) {
            if (symb == 0) return 0;
            if (symb == 1) return delims_below(idx);
            uint32_t const k = symb - 1u;                         // k >= 1: delimiter rows alias code 0 < k, the mask counts them
            auto const [p0, p1] = planes2(blk);
            uint64_t const less = k == 1 ? (~p1 & ~p0) : (k == 2 ? ~p1 : (k == 3 ? ~(p1 & p0) : ~uint64_t{0}));
            uint64_t r = delims_below(blk * 64);
            for (uint32_t j = 0; j < 4 && j < k; ++j) r += cnt(blk, j);
            return r + std::popcount(less & low);
        } else {
            uint64_t lt = 0, eq = ~uint64_t{0};
            for (int j = static_cast<int>(planes_) - 1; j >= 0; --j) {
                uint64_t const pj = plane(blk, static_cast<uint32_t>(j));
                if ((symb >> j) & 1) { lt |= eq & ~pj; eq &= pj; }
                else { eq &= ~pj; }
            }
            return pc_lo(blk, symb) + std::popcount(lt & low);
        }
    }

    auto all_ranks(size_t idx) const -> std::array<uint64_t, Sigma> {
        std::array<uint64_t, Sigma> rs{};
        for (size_t s = 0; s < Sigma; ++s) rs[s] = rank(idx, static_cast<uint8_t>(s));
        return rs;
    }
    auto all_ranks_and_prefix_ranks(size_t idx) const -> std::tuple<std::array<uint64_t, Sigma>, std::array<uint64_t, Sigma>> {
        std::array<uint64_t, Sigma> rs = all_ranks(idx), prs{};
        for (size_t s = 1; s < Sigma; ++s) prs[s] = prs[s - 1] + rs[s - 1];
        return {rs, prs};
    }

    template <typename Archive>
    void serialize(Archive& ar) { ar(n_, stride_, planes_, bytes_, delim_rows_); }

private:
    size_t n_{};
    uint32_t stride_{}, planes_{};
    std::vector<uint8_t> bytes_;
    std::vector<uint32_t> delim_rows_;      // Sigma <= 5: sorted rows holding symbol 0, one 0xFFFFFFFF behind

    static uint64_t low_mask(uint32_t off) { return (uint64_t{1} << off) - 1; }
    uint32_t cnt(size_t blk, uint32_t k) const {
        uint32_t v;
        std::memcpy(&v, bytes_.data() + blk * 32 + 4 * k, 4);
        return v;
    }
    std::pair<uint64_t, uint64_t> planes2(size_t blk) const {
        uint64_t p0, p1;
        std::memcpy(&p0, bytes_.data() + blk * 32 + 16, 8);
        std::memcpy(&p1, bytes_.data() + blk * 32 + 24, 8);
        return {p0, p1};
    }
    uint64_t delims_below(size_t row) const {
        return static_cast<uint64_t>(std::lower_bound(delim_rows_.begin(), delim_rows_.end() - 1, static_cast<uint32_t>(row)) - delim_rows_.begin());
    }
    uint64_t plane(size_t blk, uint32_t j) const {
        uint64_t v;
        std::memcpy(&v, bytes_.data() + blk * stride_ + 8 * j, 8);
        return v;
    }
    uint32_t pc(size_t blk, size_t s) const {     // # symbols <= s before the block, s < Sigma - 1
        uint32_t v;
        std::memcpy(&v, bytes_.data() + blk * stride_ + 8 * planes_ + 4 * s, 4);
        return v;
    }
    uint64_t pc_lo(size_t blk, uint32_t symb) const { return symb == 0 ? 0 : pc(blk, symb - 1); }
    uint64_t pc_hi(size_t blk, uint32_t symb) const { return symb + 1 >= Sigma ? blk * 64 : pc(blk, symb); }
};

}  // namespace fmb200
