/* fm_oracle.h -- CPU restatement ("port") of fmindex-collection's search hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker.  The product path (libfmb200.so) never links or calls this.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this port against the golden
 * vectors of the reference's own test-suite (search/checkSearches.cpp, checkSearchBacktracking.cpp,
 * fmindex/checkBiFMIndex.cpp, checkBiFMIndexCursor.cpp, string/unittest.cpp) and, when
 * oracle/_ref/libfmref.so exists, differentially against the reference itself compiled from
 * /root/reference (oracle/ref_shim.cpp).
 *
 * All reference citations are relative to /root/reference/src/fmindex-collection/.
 */
#ifndef FM_ORACLE_H
#define FM_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fmo_index fmo_index;

/* one reported cursor: (qidx, BiFMIndexCursor{lb, lbRev, len, steps}, e)  -- fmindex/BiFMIndexCursor.h:22-37 */
typedef struct {
    uint64_t qidx, lb, lb_rev, len, steps, e;
} fmo_hit;

/* one located row: fmc::Search::operator() reports (qidx, seqId, pos+offset, e) -- search/search.h:55-60 */
typedef struct {
    uint64_t qidx, seq, pos, e;
} fmo_loc;

/* counters of the algorithmic work unit of SURVEY.md §8(d) */
typedef struct {
    uint64_t extensions;    /* cursor extensions (a6/a7/a8)                                     */
    uint64_t occ_lookups;   /* 1 per extension if lb>>6 == (lb+len)>>6 else 2                    */
    uint64_t lf_steps;      /* LF steps taken by locate                                          */
    uint64_t locate_lookups;/* per located row: 1 marker test + per LF step (1 occ + 1 marker) + 1 sample fetch */
} fmo_counters;

/* ---- construction ------------------------------------------------------------------------- */
/* text = s0 0 s1 0 ... (utils.h:413-464 createSequences), n symbols in [0,sigma).  Builds SA by suffix
 * sorting (stands in for libsais, utils.h:97-129), BWT[i]=T[(SA[i]+n-1)%n] (utils.h:145-163), the BWT of
 * reverse(T) (BiFMIndex.h:82-91) when bidirectional, C (utils.h:200-206) and the text-space sampled SA
 * (BiFMIndex.h:121-135: position sampled iff offset_in_sequence % rate == 0, delimiter counts). */
fmo_index* fmo_index_build(const uint8_t* text, uint64_t n, uint32_t sigma, uint32_t sampling_rate, int bidirectional);

/* mirrors BiFMIndex(bwt, bwtRev, SparseArray) (BiFMIndex.h:40-51) / FMIndex(bwt, SparseArray) (FMIndex.h:28-32).
 * sample_bitmap: n bits (LSB-first in u64 words), bit i set iff row i carries a sample; sample_seq/pos: the
 * samples in row order. bwt_rev may be NULL (unidirectional). */
fmo_index* fmo_index_from_bwt(uint32_t sigma, uint64_t n, const uint8_t* bwt, const uint8_t* bwt_rev,
                              const uint64_t* sample_bitmap, const uint32_t* sample_seq, const uint32_t* sample_pos,
                              uint64_t n_samples);
void fmo_index_free(fmo_index* ix);

uint64_t        fmo_size(const fmo_index* ix);
uint32_t        fmo_sigma(const fmo_index* ix);
const uint8_t*  fmo_bwt(const fmo_index* ix);
const uint8_t*  fmo_bwt_rev(const fmo_index* ix);
const uint64_t* fmo_sa(const fmo_index* ix);             /* NULL when built from BWT */
const uint64_t* fmo_C(const fmo_index* ix);              /* sigma+1 entries */
uint64_t        fmo_n_samples(const fmo_index* ix);
const uint64_t* fmo_sample_bitmap(const fmo_index* ix);
const uint32_t* fmo_sample_seq(const fmo_index* ix);
const uint32_t* fmo_sample_pos(const fmo_index* ix);

/* ---- String_c concept (string/concepts.h:26-87), dir 0 = bwt, 1 = bwtRev ------------------ */
uint64_t fmo_symbol(const fmo_index* ix, int dir, uint64_t idx);
uint64_t fmo_rank(const fmo_index* ix, int dir, uint64_t idx, uint64_t symb);
uint64_t fmo_prefix_rank(const fmo_index* ix, int dir, uint64_t idx, uint64_t symb);
void     fmo_all_ranks(const fmo_index* ix, int dir, uint64_t idx, uint64_t* rs /*sigma*/);
void     fmo_all_ranks_and_prefix_ranks(const fmo_index* ix, int dir, uint64_t idx, uint64_t* rs, uint64_t* prs);

/* ---- cursor steps (fmindex/BiFMIndexCursor.h:113-128, :58-82); cur = {lb, lbRev, len, steps} ---- */
void fmo_extend_left(const fmo_index* ix, const uint64_t cur[4], uint64_t symb, uint64_t out[4]);
void fmo_extend_right(const fmo_index* ix, const uint64_t cur[4], uint64_t symb, uint64_t out[4]);
void fmo_extend_left_all(const fmo_index* ix, const uint64_t cur[4], uint64_t* out /*sigma*4*/);
void fmo_extend_right_all(const fmo_index* ix, const uint64_t cur[4], uint64_t* out /*sigma*4*/);

/* ---- searches.  Queries: concatenated symbols + offsets[Q+1].  Each returns the number of hits and
 *      stores a malloc'ed array in *out (free with fmo_free). ------------------------------------ */
/* search/SearchNoErrors.h:29-85 (batched) == :13-26 applied per query; empty intervals never reported. */
uint64_t fmo_search_exact(const fmo_index* ix, const uint8_t* qsym, const uint64_t* qoff, uint64_t nq,
                          fmo_hit** out, fmo_counters* ctr);
/* search/SearchNg26.h:427-433 with explicit scheme (n_searches x n_parts arrays pi,l,u) and partition;
 * max_hits = the `n` parameter (UINT64_MAX for "all"), clipping as in :414-421. */
uint64_t fmo_search_ng26(const fmo_index* ix, const uint8_t* qsym, const uint64_t* qoff, uint64_t nq, int edit,
                         uint32_t n_searches, uint32_t n_parts, const uint32_t* pi, const uint32_t* l, const uint32_t* u,
                         const uint32_t* partition, uint64_t max_hits, fmo_hit** out, fmo_counters* ctr);
/* The same search; additionally *keys_out receives one discovery-order key per hit (see ng26_key_edge in fm_oracle.c):
 * a test aid for the device's hit-limited search, whose post-sort by this key must reproduce the order of this
 * depth-first search.  fmo_ng26_key_layout returns the field widths (0 = the keys do not fit 64 bits). */
uint64_t fmo_search_ng26_keys(const fmo_index* ix, const uint8_t* qsym, const uint64_t* qoff, uint64_t nq, int edit,
                              uint32_t n_searches, uint32_t n_parts, const uint32_t* pi, const uint32_t* l, const uint32_t* u,
                              const uint32_t* partition, uint64_t max_hits, fmo_hit** out, uint64_t** keys_out, fmo_counters* ctr);
int fmo_ng26_key_layout(uint32_t sigma, uint32_t n_searches, uint32_t n_parts, const uint32_t* u, const uint32_t* partition,
                        uint32_t* slots, uint32_t* bits, uint32_t* maxd, uint32_t* ords);
/* search/SearchPseudo.h:171-186: expanded scheme = n_searches x L arrays pi, l, u (one entry per query symbol, search_scheme/expand.h);
 * edit != 0: search_distance (no redundancy filter), else search_hm.  Queries of another length than L are skipped. */
uint64_t fmo_search_pseudo(const fmo_index* ix, const uint8_t* qsym, const uint64_t* qoff, uint64_t nq, int edit,
                           uint32_t n_searches, uint32_t L, const uint32_t* pi, const uint32_t* l, const uint32_t* u, fmo_hit** out, fmo_counters* ctr);
/* search/Backtracking.h:85-88 (Hamming, works on unidirectional indices too) */
uint64_t fmo_search_backtracking(const fmo_index* ix, const uint8_t* qsym, const uint64_t* qoff, uint64_t nq,
                                 uint32_t max_errors, fmo_hit** out, fmo_counters* ctr);
/* locate.h:15-57 LocateLinear over every row of every hit + BiFMIndex.h:177-202 locate(); pos = pos+steps */
uint64_t fmo_locate(const fmo_index* ix, const fmo_hit* hits, uint64_t nhits, fmo_loc** out, fmo_counters* ctr);
/* BiFMIndex::locate(row) -> (seq, pos, steps) */
void fmo_locate_row(const fmo_index* ix, uint64_t row, uint64_t out[3]);
/* single_locate_step(row): returns 1 and fills (seq,pos) when the row is sampled (BiFMIndex.h:204-206) */
int  fmo_single_locate_step(const fmo_index* ix, uint64_t row, uint64_t out[2]);

void fmo_free(void* p);

#ifdef __cplusplus
}
#endif
#endif
