// fmb_kernels.cuh -- sm_100a kernels: occ-table packing (K1), String_c / cursor batch ops, exact backward
// search (K2), LF-walk locate (K4), result compaction (K5).
#pragma once
#include "fmb_device.cuh"
#include "fmb_host.hpp"

namespace fmb {

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    uint64_t z = x + 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

// T[i] = 1 + (splitmix64(seed + i) >> 32) % (sigma-1), T[n-1] = 0   (fmb200.h fmb_synth_text_device)
__global__ void synth_text_kernel(uint8_t* text, uint64_t n, uint32_t sigma, uint64_t seed) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    text[i] = (i == n - 1) ? 0 : (uint8_t)(1 + (splitmix64(seed + i) >> 32) % (sigma - 1));
}

// reads[q*L + j] = text[off(q) + j], off(q) = splitmix64(q * 0x632BE59BD9B4E019 + seed) % (n - L)
__global__ void synth_reads_kernel(const uint8_t* __restrict__ text, uint64_t n, uint64_t nq, uint32_t L, uint64_t seed, uint8_t* __restrict__ out) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= nq * L) return;
    uint64_t q = i / L, j = i % L;
    uint64_t off = splitmix64(q * 0x632BE59BD9B4E019ull + seed) % (n - L);
    out[i] = text[off + j];
}

// ---------------------------------------------------------------------------------------------------------
// K1: pack 64 BWT bytes into one DnaBlock (local counts; absolute counts are added after a scan)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_dna_kernel(const uint8_t* __restrict__ bwt, uint64_t n, uint32_t sigma,
                                                       DnaBlock* __restrict__ blocks, uint4* __restrict__ counts,
                                                       uint32_t* __restrict__ dcount, uint32_t* __restrict__ bad) {
    uint64_t blk = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t nblocks = n / 64 + 1;
    if (blk > nblocks) return;
    if (blk == nblocks) { dcount[blk] = 0; return; }
    uint64_t base = blk * 64;
    uint64_t p0 = 0, p1 = 0;
    uint32_t c[4] = {0, 0, 0, 0};
    uint32_t nd = 0, isbad = 0;
    if (base + 64 <= n && (reinterpret_cast<uintptr_t>(bwt) & 15) == 0) {
        const uint4* src = reinterpret_cast<const uint4*>(bwt + base);
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            uint4 v = __ldg(src + w);
            uint32_t words[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    uint32_t s = (words[j] >> (8 * b)) & 0xFF;
                    uint32_t r = w * 16 + j * 4 + b;
                    isbad |= (s >= sigma);
                    uint32_t k = s ? s - 1 : 0;
                    nd += (s == 0);
                    p0 |= (uint64_t)(k & 1) << r;
                    p1 |= (uint64_t)((k >> 1) & 1) << r;
                    c[0] += (s == 1); c[1] += (s == 2); c[2] += (s == 3); c[3] += (s == 4);
                }
            }
        }
    } else {
        for (uint32_t r = 0; r < 64 && base + r < n; ++r) {
            uint32_t s = bwt[base + r];
            isbad |= (s >= sigma);
            uint32_t k = s ? s - 1 : 0;
            nd += (s == 0);
            p0 |= (uint64_t)(k & 1) << r;
            p1 |= (uint64_t)((k >> 1) & 1) << r;
            c[0] += (s == 1); c[1] += (s == 2); c[2] += (s == 3); c[3] += (s == 4);
        }
    }
    DnaBlock out;
    out.cnt[0] = out.cnt[1] = out.cnt[2] = out.cnt[3] = 0;
    out.p0 = p0;
    out.p1 = p1;
    blocks[blk] = out;
    counts[blk] = make_uint4(c[0], c[1], c[2], c[3]);
    dcount[blk] = nd;
    if (isbad) atomicOr(bad, 1u);
}

__global__ void store_counts_kernel(DnaBlock* __restrict__ blocks, const uint4* __restrict__ counts, uint64_t nblocks) {
    uint64_t blk = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (blk >= nblocks) return;
    uint4 c = counts[blk];
    *reinterpret_cast<uint4*>(blocks + blk) = c;
}

__global__ void scatter_delims_kernel(const uint8_t* __restrict__ bwt, uint64_t n, const uint32_t* __restrict__ dstart,
                                      uint32_t* __restrict__ delim_rows) {
    uint64_t blk = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t nblocks = n / 64 + 1;
    if (blk >= nblocks) return;
    uint32_t a = dstart[blk], b = dstart[blk + 1];
    if (a == b) return;
    uint64_t base = blk * 64;
    for (uint32_t r = 0; r < 64 && base + r < n; ++r)
        if (bwt[base + r] == 0) delim_rows[a++] = (uint32_t)(base + r);
}

template <class OCC>
__global__ void compute_C_kernel(IndexView<OCC> ix, uint64_t* out) {
    uint32_t s = threadIdx.x;
    if (s > ix.sigma) return;
    const OCC& occ = ix.occ[0];
    typename OCC::Block b = occ.load(ix.n >> 6, s < ix.sigma ? s : ix.sigma - 1);
    // prefix_rank(n, s); for s == sigma everything is smaller
    out[s] = (s == ix.sigma) ? ix.n : occ.prefix_rank(b, ix.n, s);
}

__global__ void popcount_words_kernel(const uint64_t* __restrict__ bitmap, uint64_t have, uint64_t count, uint32_t* __restrict__ out) {
    uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (w >= count) return;
    out[w] = w < have ? __popcll(bitmap[w]) : 0;
}
__global__ void build_marks_kernel(const uint64_t* __restrict__ bitmap, uint64_t have, uint64_t words,
                                   const uint32_t* __restrict__ before, uint4* __restrict__ marks) {
    uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (w >= words) return;
    uint64_t bits = w < have ? bitmap[w] : 0;
    marks[w] = make_uint4((uint32_t)bits, (uint32_t)(bits >> 32), before[w], 0);
}
__global__ void export_marks_kernel(const uint4* __restrict__ marks, uint64_t have, uint64_t* __restrict__ out) {
    uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (w >= have) return;
    uint4 m = marks[w];
    out[w] = (uint64_t)m.x | ((uint64_t)m.y << 32);
}
__global__ void zip_samples_kernel(const uint32_t* __restrict__ seq, const uint32_t* __restrict__ pos, uint64_t count, uint2* __restrict__ out) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= count) return;
    out[i] = make_uint2(seq[i], pos[i]);
}

template <class OCC>
__global__ void unpack_bwt_kernel(IndexView<OCC> ix, int dir, uint8_t* __restrict__ out) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= ix.n) return;
    const OCC& occ = ix.occ[dir];
    typename OCC::Block b = occ.load((uint32_t)(i >> 6), 0);
    out[i] = (uint8_t)occ.symbol(b, (row_t)i);
}

// ---------------------------------------------------------------------------------------------------------
// String_c batch ops: op 0 symbol, 1 rank, 2 prefix_rank, 3 all_ranks_and_prefix_ranks
// (string/concepts.h:48-87; all_ranks_and_prefix_ranks: prs[0] = 0, prs[c] = prs[c-1] + rs[c-1],
//  string/InterleavedBitvector.h:141-160)
// ---------------------------------------------------------------------------------------------------------
template <class OCC>
__global__ void string_op_kernel(IndexView<OCC> ix, int dir, int op, const uint64_t* __restrict__ idx,
                                 const uint8_t* __restrict__ symb, uint64_t count, uint64_t* __restrict__ out,
                                 uint64_t* __restrict__ out2) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= count) return;
    const OCC& occ = ix.occ[dir];
    row_t row = (row_t)idx[i];
    if (op == 0) {
        typename OCC::Block b = occ.load(row >> 6, 0);
        out[i] = occ.symbol(b, row);
    } else if (op == 1) {
        typename OCC::Block b = occ.load(row >> 6, symb[i]);
        out[i] = occ.rank(b, row, symb[i]);
    } else if (op == 2) {
        uint32_t s = symb[i];
        if (s >= ix.sigma) { out[i] = row; return; }
        typename OCC::Block b = occ.load(row >> 6, s);
        out[i] = occ.prefix_rank(b, row, s);
    } else {
        uint64_t acc = 0;
        for (uint32_t s = 0; s < ix.sigma; ++s) {
            typename OCC::Block b = occ.load(row >> 6, s);
            uint64_t r = occ.rank(b, row, s);
            out[i * ix.sigma + s] = r;
            out2[i * ix.sigma + s] = acc;
            acc += r;
        }
    }
}

// cursor batch ops (fmindex/BiFMIndexCursor.h:113-128; :58-82 for all symbols)
template <class OCC>
__global__ void cursor_op_kernel(IndexView<OCC> ix, int right, int all, const uint64_t* __restrict__ cur,
                                 const uint8_t* __restrict__ symb, uint64_t count, uint64_t* __restrict__ out) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= count) return;
    Cursor c{(row_t)cur[4 * i], (row_t)cur[4 * i + 1], (row_t)cur[4 * i + 2], (uint32_t)cur[4 * i + 3]};
    uint32_t lookups = 0;
    if (!all) {
        Cursor o = extend_bi(ix, c, symb[i], right, lookups);
        out[4 * i] = o.lb; out[4 * i + 1] = o.lb_rev; out[4 * i + 2] = o.len; out[4 * i + 3] = o.steps;
    } else {
        for (uint32_t s = 0; s < ix.sigma; ++s) {
            Cursor o = extend_bi(ix, c, s, right, lookups);
            uint64_t* p = out + (i * ix.sigma + s) * 4;
            p[0] = o.lb; p[1] = o.lb_rev; p[2] = o.len; p[3] = o.steps;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// K2: exact backward search, one query per thread, full occupancy (search/SearchNoErrors.h:13-26 per query).
// The reference keeps 32 queries in flight per core in software (:29-85); here every resident thread is one
// in-flight query: 2048 per SM, ~300 k per GPU, which is what hides the ~1 us random-sector latency.
// Query symbols are read as aligned 16-byte chunks (register buffered, one load per 16 steps).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t chunk_byte(const uint4& v, uint32_t i) {
    uint32_t w = (i & 8) ? ((i & 4) ? v.w : v.z) : ((i & 4) ? v.y : v.x);
    return (w >> (8 * (i & 3))) & 0xFF;
}

template <class OCC, bool COUNT>
__global__ void __launch_bounds__(256) exact_search_kernel(IndexView<OCC> ix, const uint8_t* __restrict__ qsym,
                                                           const uint64_t* __restrict__ qoff, uint32_t nq,
                                                           uint32_t* __restrict__ out_lb, uint32_t* __restrict__ out_len,
                                                           unsigned long long* __restrict__ counters) {
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t ext = 0, lookups = 0;
    if (q < nq) {
        uint64_t off = qoff[q];
        uint32_t L = (uint32_t)(qoff[q + 1] - off);
        row_t lb = 0, len = ix.n;
        uint64_t cur_chunk = ~uint64_t(0);
        uint4 chunk = make_uint4(0, 0, 0, 0);
        for (uint32_t pos = L; pos-- > 0;) {
            uint64_t a = off + pos;
            if ((a >> 4) != cur_chunk) {
                cur_chunk = a >> 4;
                chunk = __ldg(reinterpret_cast<const uint4*>(qsym) + cur_chunk);
            }
            uint32_t c = chunk_byte(chunk, (uint32_t)(a & 15));
            if (c >= ix.sigma) { len = 0; break; }
            extend_left_uni(ix, lb, len, c, lookups);
            ++ext;
            if (len == 0) break;
        }
        out_lb[q] = lb;
        out_len[q] = len;
    }
    if (COUNT) {
        for (int o = 16; o > 0; o >>= 1) {
            ext += __shfl_xor_sync(0xFFFFFFFFu, ext, o);
            lookups += __shfl_xor_sync(0xFFFFFFFFu, lookups, o);
        }
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(counters + 0, (unsigned long long)ext);
            atomicAdd(counters + 1, (unsigned long long)lookups);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// K5: ordered compaction of the per-query intervals into hit records
// ---------------------------------------------------------------------------------------------------------
__global__ void flag_nonzero_kernel(const uint32_t* __restrict__ len, uint64_t count, uint32_t* __restrict__ flag) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= count) return;
    flag[i] = len[i] ? 1u : 0u;
}
__global__ void compact_exact_hits_kernel(const uint32_t* __restrict__ lb, const uint32_t* __restrict__ len,
                                          const uint32_t* __restrict__ pos, const uint64_t* __restrict__ qoff, uint32_t nq, uint32_t qidx_base,
                                          HitRec* __restrict__ hits) {
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    uint32_t l = len[q];
    if (!l) return;
    HitRec h;
    h.qidx = q + qidx_base; h.lb = lb[q]; h.lb_rev = 0; h.len = l;
    h.steps = (uint32_t)(qoff[q + 1] - qoff[q]);
    h.e = 0;
    hits[pos[q]] = h;
}
__global__ void hit_lengths_kernel(const HitRec* __restrict__ hits, uint64_t nh, uint32_t* __restrict__ out) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i > nh) return;
    out[i] = i < nh ? hits[i].len : 0;
}

// ---------------------------------------------------------------------------------------------------------
// K4: locate.  One SA row per thread; LF walk to the nearest sampled row (fmindex/BiFMIndex.h:177-202):
//     while row not sampled: c = BWT[row]; row = C[c] + rank(row, c); ++steps
// The marker word and the occ block of a row are independent loads and are issued together.
// starts[] = exclusive prefix sum of the hit interval lengths; a thread finds its hit by binary search.
// ---------------------------------------------------------------------------------------------------------
template <class OCC, bool COUNT>
__global__ void __launch_bounds__(256) locate_kernel(IndexView<OCC> ix, const HitRec* __restrict__ hits,
                                                     const uint32_t* __restrict__ starts, uint32_t nh, uint32_t total,
                                                     LocRec* __restrict__ out, unsigned long long* __restrict__ counters) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t steps = 0;
    if (t < total) {
        // largest h with starts[h] <= t
        uint32_t lo = 0, hi = nh;
        while (hi - lo > 1) {
            uint32_t mid = (lo + hi) >> 1;
            if (__ldg(starts + mid) <= t) lo = mid; else hi = mid;
        }
        HitRec h = hits[lo];
        row_t row = h.lb + (t - __ldg(starts + lo));
        const OCC& occ = ix.occ[0];
        uint2 sample;
        for (;;) {
            uint4 m = __ldg(ix.marks + (row >> 6));
            typename OCC::Block b = occ.load(row >> 6, 0);
            uint64_t bits = (uint64_t)m.x | ((uint64_t)m.y << 32);
            uint32_t o = row & 63;
            if ((bits >> o) & 1) {
                sample = __ldg(ix.samples + (m.z + __popcll(bits & low_mask(o))));
                break;
            }
            uint32_t c = occ.symbol(b, row);
            if (OCC::kSymbolLoad) b = occ.load(row >> 6, c);
            row = ix.C[c] + occ.rank(b, row, c);
            ++steps;
        }
        LocRec r;
        r.qidx = h.qidx; r.seq = sample.x; r.pos = sample.y + steps; r.e = h.e;
        out[t] = r;
    }
    if (COUNT) {
        uint32_t s = steps;
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
        if ((threadIdx.x & 31) == 0 && s) {
            atomicAdd(counters + 2, (unsigned long long)s);
            atomicAdd(counters + 1, (unsigned long long)s);
        }
    }
}

// index.locate(row) for arbitrary rows (fmindex/BiFMIndex.h:177-202): same LF walk as locate_kernel, row list input
template <class OCC>
__global__ void __launch_bounds__(256) locate_rows_kernel(IndexView<OCC> ix, const uint64_t* __restrict__ rows, uint64_t count,
                                                          uint32_t* __restrict__ seq, uint32_t* __restrict__ pos, uint64_t* __restrict__ steps_out) {
    uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (t >= count) return;
    row_t row = (row_t)rows[t];
    const OCC& occ = ix.occ[0];
    uint32_t steps = 0;
    for (;;) {
        uint4 m = __ldg(ix.marks + (row >> 6));
        typename OCC::Block b = occ.load(row >> 6, 0);
        uint64_t bits = (uint64_t)m.x | ((uint64_t)m.y << 32);
        uint32_t o = row & 63;
        if ((bits >> o) & 1) {
            uint2 sample = __ldg(ix.samples + (m.z + __popcll(bits & low_mask(o))));
            seq[t] = sample.x;
            pos[t] = sample.y;
            steps_out[t] = steps;
            return;
        }
        uint32_t c = occ.symbol(b, row);
        if (OCC::kSymbolLoad) b = occ.load(row >> 6, c);
        row = ix.C[c] + occ.rank(b, row, c);
        ++steps;
    }
}

// annotatedArray.value(row) (suffixarray/SparseArray.h:63-70)
template <class OCC>
__global__ void sample_value_kernel(IndexView<OCC> ix, const uint64_t* __restrict__ rows, uint64_t count, uint8_t* __restrict__ has,
                                    uint32_t* __restrict__ seq, uint32_t* __restrict__ pos) {
    uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (t >= count) return;
    row_t row = (row_t)rows[t];
    uint4 m = __ldg(ix.marks + (row >> 6));
    uint64_t bits = (uint64_t)m.x | ((uint64_t)m.y << 32);
    uint32_t o = row & 63;
    bool h = (bits >> o) & 1;
    has[t] = h ? 1 : 0;
    uint2 sample = make_uint2(0, 0);
    if (h) sample = __ldg(ix.samples + (m.z + __popcll(bits & low_mask(o))));
    seq[t] = sample.x;
    pos[t] = sample.y;
}

}  // namespace fmb
