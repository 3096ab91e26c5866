// fmb200/adapt.hpp -- attach a device image to an EXISTING reference index object.
//
// `fmb200::attach(refIndex)` reads a fmc::BiFMIndex / fmc::FMIndex through its public members only
// (fmindex/BiFMIndex.h:32-35: bwt, bwtRev, C, annotatedArray; String_c::symbol(i), string/concepts.h:48-50;
// SparseArray::value(i), suffixarray/SparseArray.h:63-70), uploads it through fmb_index_create and returns an
// `Attached<RefIndex>` that every fmb200 search accepts.  Delegates then receive the REFERENCE's own cursor types
// (fmc::BiFMIndexCursor<RefIndex>{index, lb, lbRev, len, steps}, fmindex/BiFMIndexCursor.h:31-37;
// fmc::LeftBiFMIndexCursor, :223) pointing at the caller's index, so user code behind the delegate keeps compiling.
//
// This header does not include any reference header: it is a template over "an index with the reference's public
// members", so it compiles with or without fmindex-collection on the include path.  The cursor templates to hand
// out are passed explicitly:
//
//     auto dev = fmb200::attach<fmc::BiFMIndexCursor, fmc::LeftBiFMIndexCursor>(index);
//     fmb200::search_ng26::search<true>(dev, queries, scheme, partition, [&](size_t qidx, auto cursor, size_t e) {...});
#pragma once
#include <optional>

#include "index.hpp"

namespace fmb200 {

template <typename RefIndex, template <typename> class Cursor, template <typename> class LeftCursor>
struct Attached {
    static constexpr size_t Sigma = RefIndex::Sigma;
    static constexpr size_t FirstSymb = [] { if constexpr (requires { RefIndex::FirstSymb; }) return size_t{RefIndex::FirstSymb}; else return size_t{1}; }();
    using cursor_t = Cursor<RefIndex>;
    using left_cursor_t = LeftCursor<RefIndex>;
    using LEntry = std::tuple<uint32_t, uint32_t, size_t>;

    RefIndex const* ref{};
    detail::IndexHandle h;
    // presence of this member marks the index as bidirectional for the search front-ends
    DeviceString<Sigma> bwt, bwtRev;

    fmb_index const* handle() const { return h.get(); }
    size_t size() const { return ref->size(); }
    auto locate(size_t idx) const -> LEntry { return locate_rows_raw(h.get(), std::array<uint64_t, 1>{idx})[0]; }
    // searches construct `cursor_t{index, ...}`: the conversion makes those cursors point at the reference index
    operator RefIndex const&() const { return *ref; }
};

template <template <typename> class Cursor, template <typename> class LeftCursor, typename RefIndex>
auto attach(RefIndex const& ref, int device = 0) -> Attached<RefIndex, Cursor, LeftCursor> {
    // BiFMIndex<...>::NoDelim / ::ReuseRev (fmindex/BiFMIndex.h:22-28): the variant flags of the reference index carry over
    constexpr bool noDelim = [] { if constexpr (requires { RefIndex::Delim_v; }) return !RefIndex::Delim_v; else return false; }();
    constexpr bool reuseRev = [] { if constexpr (requires { RefIndex::ReuseRev_v; }) return bool{RefIndex::ReuseRev_v}; else return false; }();
    constexpr bool bidirectional = requires { ref.bwtRev.symbol(size_t{}); };          // a stored bwtRev (not with ReuseRev)
    size_t const n = ref.size();
    std::vector<uint8_t> bwt(n), bwtRev;
    for (size_t i = 0; i < n; ++i) bwt[i] = static_cast<uint8_t>(ref.bwt.symbol(i));
    if constexpr (bidirectional) {
        bwtRev.resize(n);
        for (size_t i = 0; i < n; ++i) bwtRev[i] = static_cast<uint8_t>(ref.bwtRev.symbol(i));
    }
    SparseArray sa;
    sa.rows = n;
    sa.bitmap.assign((n + 63) / 64, 0);
    for (size_t i = 0; i < n; ++i) {
        auto v = ref.annotatedArray.value(i);
        if (!v) continue;
        sa.bitmap[i / 64] |= uint64_t{1} << (i % 64);
        sa.seq.push_back(static_cast<uint32_t>(std::get<0>(*v)));
        sa.pos.push_back(static_cast<uint32_t>(std::get<1>(*v)));
    }
    fmb_index* raw{};
    check(fmb_index_create_ex(&raw, device, RefIndex::Sigma, n, bwt.data(), bidirectional ? bwtRev.data() : nullptr, sa.bitmap.data(), sa.seq.data(),
                              sa.pos.data(), sa.seq.size(), (noDelim ? FMB_INDEX_NO_DELIM : 0u) | (reuseRev ? FMB_INDEX_REUSE_REV : 0u)));
    Attached<RefIndex, Cursor, LeftCursor> a;
    a.ref = &ref;
    a.h.reset(raw);
    a.bwt = DeviceString<RefIndex::Sigma>{raw, 0, n};
    a.bwtRev = DeviceString<RefIndex::Sigma>{raw, (bidirectional || reuseRev) ? 1 : 0, n};
    // the device C must equal the reference's (utils.h:200-206)
    uint64_t c[RefIndex::Sigma + 1];
    check(fmb_index_get_C(raw, c));
    for (size_t s = 0; s <= RefIndex::Sigma; ++s)
        if (c[s] != ref.C[s]) throw std::runtime_error("fmb200::attach: C array of the device image differs from the reference index");
    return a;
}

}  // namespace fmb200
