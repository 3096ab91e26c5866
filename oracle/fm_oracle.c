/* fm_oracle.c -- plain-C restatement of fmindex-collection's search hot path.
 * TEST INFRASTRUCTURE ONLY (see fm_oracle.h).  Citations are relative to
 * /root/reference/src/fmindex-collection/.
 *
 * The occurrence table here is deliberately naive (per-64-row checkpoints + byte scan): it
 * implements the String_c *semantics* (string/concepts.h:26-87), not any particular layout.
 */
#define _GNU_SOURCE
#include "fm_oracle.h"

#include <stdlib.h>
#include <string.h>

struct fmo_index {
    uint32_t sigma;
    uint64_t n;
    uint8_t* bwt[2];      /* [0] = bwt, [1] = bwtRev (NULL if unidirectional) */
    uint64_t* cp[2];      /* checkpoints: cp[dir][(i>>6)*sigma + c] = #c in bwt[0, (i>>6)<<6) */
    uint64_t* C;          /* sigma+1 */
    uint64_t* sa;         /* full SA when built from text, else NULL */
    uint64_t* sample_bitmap;
    uint64_t* sample_rank; /* per 64-bit word: #set bits before the word */
    uint32_t* sample_seq;
    uint32_t* sample_pos;
    uint64_t n_samples;
};

/* ------------------------------------------------------------------------------------------ */
/* suffix sorting: prefix doubling.  Stands in for libsais (utils.h:97-129).  Order = ordinary  */
/* lexicographic order of the non-cyclic suffixes (a proper prefix sorts first).               */
/* ------------------------------------------------------------------------------------------ */
typedef struct { uint64_t k1, k2, idx; } sa_item;

static int sa_item_cmp(const void* a, const void* b) {
    const sa_item* x = (const sa_item*)a;
    const sa_item* y = (const sa_item*)b;
    if (x->k1 != y->k1) return x->k1 < y->k1 ? -1 : 1;
    if (x->k2 != y->k2) return x->k2 < y->k2 ? -1 : 1;
    return 0;
}

static uint64_t* build_sa(const uint8_t* text, uint64_t n, uint32_t sigma) {
    uint64_t* sa = (uint64_t*)malloc((n ? n : 1) * sizeof(uint64_t));
    if (n == 0) return sa;
    uint64_t* rank = (uint64_t*)malloc(n * sizeof(uint64_t));
    sa_item* items = (sa_item*)malloc(n * sizeof(sa_item));
    /* initial key: the first K symbols, (symbol+1) packed big-endian, 0 = past the end */
    uint32_t bits = 1;
    while ((1u << bits) < sigma + 1) ++bits;
    uint32_t K = 64 / bits;
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t key = 0;
        for (uint32_t j = 0; j < K; ++j) {
            uint64_t v = (i + j < n) ? (uint64_t)text[i + j] + 1 : 0;
            key = (key << bits) | v;
        }
        items[i].k1 = key; items[i].k2 = 0; items[i].idx = i;
    }
    uint64_t h = K;
    for (;;) {
        qsort(items, n, sizeof(sa_item), sa_item_cmp);
        /* ranks are 1-based group starts so that 0 can mean "past the end" */
        uint64_t groups = 0;
        for (uint64_t i = 0; i < n; ++i) {
            if (i == 0 || sa_item_cmp(&items[i - 1], &items[i]) != 0) ++groups, rank[items[i].idx] = i + 1;
            else rank[items[i].idx] = rank[items[i - 1].idx];
        }
        if (groups == n) break;
        for (uint64_t i = 0; i < n; ++i) {
            uint64_t idx = items[i].idx;
            items[i].k1 = rank[idx];
            items[i].k2 = (idx + h < n) ? rank[idx + h] : 0;
        }
        h *= 2;
    }
    for (uint64_t i = 0; i < n; ++i) sa[i] = items[i].idx;
    free(items);
    free(rank);
    return sa;
}

/* ------------------------------------------------------------------------------------------ */
static uint64_t* build_checkpoints(const uint8_t* bwt, uint64_t n, uint32_t sigma) {
    uint64_t nblocks = n / 64 + 1;
    uint64_t* cp = (uint64_t*)calloc(nblocks * sigma, sizeof(uint64_t));
    uint64_t* acc = (uint64_t*)calloc(sigma, sizeof(uint64_t));
    for (uint64_t b = 0; b < nblocks; ++b) {
        memcpy(cp + b * sigma, acc, sigma * sizeof(uint64_t));
        uint64_t end = (b + 1) * 64 < n ? (b + 1) * 64 : n;
        for (uint64_t i = b * 64; i < end; ++i) acc[bwt[i]] += 1;
    }
    free(acc);
    return cp;
}

static void finish_index(fmo_index* ix) {
    for (int d = 0; d < 2; ++d)
        ix->cp[d] = ix->bwt[d] ? build_checkpoints(ix->bwt[d], ix->n, ix->sigma) : NULL;
    /* computeC, utils.h:200-206: C[s] = prefix_rank(n, s), s = 0..sigma */
    ix->C = (uint64_t*)calloc(ix->sigma + 1, sizeof(uint64_t));
    for (uint64_t i = 0; i < ix->n; ++i) ix->C[ix->bwt[0][i] + 1] += 1;
    for (uint32_t s = 1; s <= ix->sigma; ++s) ix->C[s] += ix->C[s - 1];
    uint64_t words = ix->n / 64 + 1;
    ix->sample_rank = (uint64_t*)calloc(words, sizeof(uint64_t));
    uint64_t acc = 0;
    for (uint64_t w = 0; w < words; ++w) {
        ix->sample_rank[w] = acc;
        acc += (uint64_t)__builtin_popcountll(ix->sample_bitmap[w]);
    }
}

fmo_index* fmo_index_build(const uint8_t* text, uint64_t n, uint32_t sigma, uint32_t rate, int bidirectional) {
    fmo_index* ix = (fmo_index*)calloc(1, sizeof(fmo_index));
    ix->sigma = sigma;
    ix->n = n;
    ix->sa = build_sa(text, n, sigma);
    ix->bwt[0] = (uint8_t*)malloc(n ? n : 1);
    for (uint64_t i = 0; i < n; ++i) ix->bwt[0][i] = text[(ix->sa[i] + n - 1) % n];   /* utils.h:145-163 */
    if (bidirectional) {
        uint8_t* rev = (uint8_t*)malloc(n ? n : 1);
        for (uint64_t i = 0; i < n; ++i) rev[i] = text[n - 1 - i];                       /* BiFMIndex.h:82-91 */
        uint64_t* sar = build_sa(rev, n, sigma);
        ix->bwt[1] = (uint8_t*)malloc(n ? n : 1);
        for (uint64_t i = 0; i < n; ++i) ix->bwt[1][i] = rev[(sar[i] + n - 1) % n];
        free(sar);
        free(rev);
    }
    /* text-space sampling, BiFMIndex.h:121-135: walk the text, (refId,pos) restart after each delimiter */
    uint32_t* tseq = (uint32_t*)malloc((n ? n : 1) * sizeof(uint32_t));
    uint32_t* tpos = (uint32_t*)malloc((n ? n : 1) * sizeof(uint32_t));
    {
        uint32_t seq = 0, pos = 0;
        for (uint64_t i = 0; i < n; ++i) {
            tseq[i] = seq; tpos[i] = pos;
            ++pos;
            if (text[i] == 0) { ++seq; pos = 0; }
        }
    }
    uint64_t words = n / 64 + 1;
    ix->sample_bitmap = (uint64_t*)calloc(words, sizeof(uint64_t));
    uint64_t ns = 0;
    for (uint64_t i = 0; i < n; ++i) if (tpos[ix->sa[i]] % rate == 0) ++ns;
    ix->sample_seq = (uint32_t*)malloc((ns ? ns : 1) * sizeof(uint32_t));
    ix->sample_pos = (uint32_t*)malloc((ns ? ns : 1) * sizeof(uint32_t));
    ns = 0;
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t p = ix->sa[i];
        if (tpos[p] % rate == 0) {
            ix->sample_bitmap[i >> 6] |= 1ull << (i & 63);
            ix->sample_seq[ns] = tseq[p];
            ix->sample_pos[ns] = tpos[p];
            ++ns;
        }
    }
    ix->n_samples = ns;
    free(tseq);
    free(tpos);
    finish_index(ix);
    return ix;
}

fmo_index* fmo_index_from_bwt(uint32_t sigma, uint64_t n, const uint8_t* bwt, const uint8_t* bwt_rev,
                              const uint64_t* sample_bitmap, const uint32_t* sample_seq, const uint32_t* sample_pos,
                              uint64_t n_samples) {
    fmo_index* ix = (fmo_index*)calloc(1, sizeof(fmo_index));
    ix->sigma = sigma;
    ix->n = n;
    ix->bwt[0] = (uint8_t*)malloc(n ? n : 1);
    memcpy(ix->bwt[0], bwt, n);
    if (bwt_rev) {
        ix->bwt[1] = (uint8_t*)malloc(n ? n : 1);
        memcpy(ix->bwt[1], bwt_rev, n);
    }
    uint64_t words = n / 64 + 1;
    ix->sample_bitmap = (uint64_t*)calloc(words, sizeof(uint64_t));
    memcpy(ix->sample_bitmap, sample_bitmap, ((n + 63) / 64) * sizeof(uint64_t));
    ix->sample_seq = (uint32_t*)malloc((n_samples ? n_samples : 1) * sizeof(uint32_t));
    ix->sample_pos = (uint32_t*)malloc((n_samples ? n_samples : 1) * sizeof(uint32_t));
    memcpy(ix->sample_seq, sample_seq, n_samples * sizeof(uint32_t));
    memcpy(ix->sample_pos, sample_pos, n_samples * sizeof(uint32_t));
    ix->n_samples = n_samples;
    finish_index(ix);
    return ix;
}

void fmo_index_free(fmo_index* ix) {
    if (!ix) return;
    for (int d = 0; d < 2; ++d) { free(ix->bwt[d]); free(ix->cp[d]); }
    free(ix->C); free(ix->sa); free(ix->sample_bitmap); free(ix->sample_rank);
    free(ix->sample_seq); free(ix->sample_pos);
    free(ix);
}

uint64_t        fmo_size(const fmo_index* ix) { return ix->n; }
uint32_t        fmo_sigma(const fmo_index* ix) { return ix->sigma; }
const uint8_t*  fmo_bwt(const fmo_index* ix) { return ix->bwt[0]; }
const uint8_t*  fmo_bwt_rev(const fmo_index* ix) { return ix->bwt[1]; }
const uint64_t* fmo_sa(const fmo_index* ix) { return ix->sa; }
const uint64_t* fmo_C(const fmo_index* ix) { return ix->C; }
uint64_t        fmo_n_samples(const fmo_index* ix) { return ix->n_samples; }
const uint64_t* fmo_sample_bitmap(const fmo_index* ix) { return ix->sample_bitmap; }
const uint32_t* fmo_sample_seq(const fmo_index* ix) { return ix->sample_seq; }
const uint32_t* fmo_sample_pos(const fmo_index* ix) { return ix->sample_pos; }
void fmo_free(void* p) { free(p); }

/* ------------------------------------------------------------------------------------------ */
/* String_c semantics: rank(i,c) = #{j<i : bwt[j]==c}; prefix_rank(i,c) = #{j<i : bwt[j] < c}  */
/* (exclusive, c may equal sigma) -- string/concepts.h:52-64, InterleavedBitvector.h:112-128.  */
/* ------------------------------------------------------------------------------------------ */
uint64_t fmo_symbol(const fmo_index* ix, int dir, uint64_t idx) { return ix->bwt[dir][idx]; }

uint64_t fmo_rank(const fmo_index* ix, int dir, uint64_t idx, uint64_t symb) {
    uint64_t b = idx >> 6;
    uint64_t r = ix->cp[dir][b * ix->sigma + symb];
    const uint8_t* p = ix->bwt[dir];
    for (uint64_t i = b << 6; i < idx; ++i) r += (p[i] == symb);
    return r;
}

uint64_t fmo_prefix_rank(const fmo_index* ix, int dir, uint64_t idx, uint64_t symb) {
    uint64_t b = idx >> 6;
    uint64_t r = 0;
    for (uint64_t s = 0; s < symb; ++s) r += ix->cp[dir][b * ix->sigma + s];
    const uint8_t* p = ix->bwt[dir];
    for (uint64_t i = b << 6; i < idx; ++i) r += (p[i] < symb);
    return r;
}

void fmo_all_ranks(const fmo_index* ix, int dir, uint64_t idx, uint64_t* rs) {
    uint64_t b = idx >> 6;
    memcpy(rs, ix->cp[dir] + b * ix->sigma, ix->sigma * sizeof(uint64_t));
    const uint8_t* p = ix->bwt[dir];
    for (uint64_t i = b << 6; i < idx; ++i) rs[p[i]] += 1;
}

/* InterleavedBitvector.h:141-160: prs[0] = 0, prs[c] = prs[c-1] + rs[c-1] */
void fmo_all_ranks_and_prefix_ranks(const fmo_index* ix, int dir, uint64_t idx, uint64_t* rs, uint64_t* prs) {
    fmo_all_ranks(ix, dir, idx, rs);
    prs[0] = 0;
    for (uint32_t c = 1; c < ix->sigma; ++c) prs[c] = prs[c - 1] + rs[c - 1];
}

/* ------------------------------------------------------------------------------------------ */
/* cursors                                                                                      */
/* ------------------------------------------------------------------------------------------ */
typedef struct { uint64_t lb, lb_rev, len, steps; } cursor_t;

static void count_ext(fmo_counters* ctr, uint64_t lo, uint64_t len) {
    if (!ctr) return;
    ctr->extensions += 1;
    ctr->occ_lookups += ((lo >> 6) == ((lo + len) >> 6)) ? 1 : 2;
}

/* BiFMIndexCursor::extendLeft(symb), BiFMIndexCursor.h:113-120 */
static cursor_t ext_left(const fmo_index* ix, cursor_t c, uint64_t s, fmo_counters* ctr) {
    count_ext(ctr, c.lb, c.len);
    cursor_t r;
    uint64_t newLb = fmo_rank(ix, 0, c.lb, s);
    r.lb_rev = c.lb_rev + fmo_prefix_rank(ix, 0, c.lb + c.len, s) - fmo_prefix_rank(ix, 0, c.lb, s);
    r.len = fmo_rank(ix, 0, c.lb + c.len, s) - newLb;
    r.lb = newLb + ix->C[s];
    r.steps = c.steps + 1;
    return r;
}

/* BiFMIndexCursor::extendRight(symb), BiFMIndexCursor.h:121-128 */
static cursor_t ext_right(const fmo_index* ix, cursor_t c, uint64_t s, fmo_counters* ctr) {
    count_ext(ctr, c.lb_rev, c.len);
    cursor_t r;
    r.lb = c.lb + fmo_prefix_rank(ix, 1, c.lb_rev + c.len, s) - fmo_prefix_rank(ix, 1, c.lb_rev, s);
    uint64_t newLbRev = fmo_rank(ix, 1, c.lb_rev, s);
    r.len = fmo_rank(ix, 1, c.lb_rev + c.len, s) - newLbRev;
    r.lb_rev = newLbRev + ix->C[s];
    r.steps = c.steps + 1;
    return r;
}

/* BiFMIndexCursor::extendLeft() / extendRight(), BiFMIndexCursor.h:58-82 (all sigma children) */
static void ext_all(const fmo_index* ix, cursor_t c, int right, cursor_t* out, fmo_counters* ctr) {
    uint32_t sg = ix->sigma;
    uint64_t rs1[256], prs1[256], rs2[256], prs2[256];
    if (!right) {
        count_ext(ctr, c.lb, c.len);
        fmo_all_ranks_and_prefix_ranks(ix, 0, c.lb, rs1, prs1);
        fmo_all_ranks_and_prefix_ranks(ix, 0, c.lb + c.len, rs2, prs2);
        for (uint32_t i = 0; i < sg; ++i) {
            out[i].lb = rs1[i] + ix->C[i];
            out[i].lb_rev = c.lb_rev + prs2[i] - prs1[i];
            out[i].len = rs2[i] - rs1[i];
            out[i].steps = c.steps + 1;
        }
    } else {
        count_ext(ctr, c.lb_rev, c.len);
        fmo_all_ranks_and_prefix_ranks(ix, 1, c.lb_rev, rs1, prs1);
        fmo_all_ranks_and_prefix_ranks(ix, 1, c.lb_rev + c.len, rs2, prs2);
        for (uint32_t i = 0; i < sg; ++i) {
            out[i].lb = c.lb + prs2[i] - prs1[i];
            out[i].lb_rev = rs1[i] + ix->C[i];
            out[i].len = rs2[i] - rs1[i];
            out[i].steps = c.steps + 1;
        }
    }
}

static cursor_t cur_from(const uint64_t c[4]) { cursor_t r = {c[0], c[1], c[2], c[3]}; return r; }
static void cur_to(cursor_t c, uint64_t* o) { o[0] = c.lb; o[1] = c.lb_rev; o[2] = c.len; o[3] = c.steps; }

void fmo_extend_left(const fmo_index* ix, const uint64_t cur[4], uint64_t symb, uint64_t out[4]) {
    cur_to(ext_left(ix, cur_from(cur), symb, NULL), out);
}
void fmo_extend_right(const fmo_index* ix, const uint64_t cur[4], uint64_t symb, uint64_t out[4]) {
    cur_to(ext_right(ix, cur_from(cur), symb, NULL), out);
}
void fmo_extend_left_all(const fmo_index* ix, const uint64_t cur[4], uint64_t* out) {
    cursor_t cs[256];
    ext_all(ix, cur_from(cur), 0, cs, NULL);
    for (uint32_t i = 0; i < ix->sigma; ++i) cur_to(cs[i], out + 4 * i);
}
void fmo_extend_right_all(const fmo_index* ix, const uint64_t cur[4], uint64_t* out) {
    cursor_t cs[256];
    ext_all(ix, cur_from(cur), 1, cs, NULL);
    for (uint32_t i = 0; i < ix->sigma; ++i) cur_to(cs[i], out + 4 * i);
}

/* ------------------------------------------------------------------------------------------ */
/* growable hit list                                                                           */
/* ------------------------------------------------------------------------------------------ */
typedef struct { fmo_hit* v; uint64_t n, cap; } hitvec;

static void hit_push(hitvec* hv, uint64_t qidx, cursor_t c, uint64_t e) {
    if (hv->n == hv->cap) {
        hv->cap = hv->cap ? hv->cap * 2 : 1024;
        hv->v = (fmo_hit*)realloc(hv->v, hv->cap * sizeof(fmo_hit));
    }
    fmo_hit h = {qidx, c.lb, c.lb_rev, c.len, c.steps, e};
    hv->v[hv->n++] = h;
}

/* ------------------------------------------------------------------------------------------ */
/* exact search -- search/SearchNoErrors.h:13-26 per query (the batched form :29-85 yields the  */
/* same set; it only interleaves queries).  Unidirectional cursor: lbRev is not maintained by   */
/* LeftBiFMIndexCursor (BiFMIndexCursor.h:203-256), reported here as 0.                        */
/* ------------------------------------------------------------------------------------------ */
uint64_t fmo_search_exact(const fmo_index* ix, const uint8_t* qsym, const uint64_t* qoff, uint64_t nq,
                          fmo_hit** out, fmo_counters* ctr) {
    hitvec hv = {0, 0, 0};
    for (uint64_t q = 0; q < nq; ++q) {
        const uint8_t* query = qsym + qoff[q];
        uint64_t L = qoff[q + 1] - qoff[q];
        uint64_t lb = 0, len = ix->n, steps = 0;
        int dead = 0;
        for (uint64_t i = 0; i < L; ++i) {
            uint64_t s = query[L - i - 1];
            count_ext(ctr, lb, len);
            uint64_t newLb = fmo_rank(ix, 0, lb, s);                 /* BiFMIndexCursor.h:248-255 */
            len = fmo_rank(ix, 0, lb + len, s) - newLb;
            lb = newLb + ix->C[s];
            ++steps;
            if (len == 0) { dead = 1; break; }
        }
        if (!dead && len > 0) {                                        /* SearchNoErrors.h:72-76 */
            cursor_t c = {lb, 0, len, steps};
            hit_push(&hv, q, c, 0);
        }
    }
    *out = hv.v;
    return hv.n;
}

/* ------------------------------------------------------------------------------------------ */
/* search_ng26 -- search/SearchNg26.h:18-366, function for function.                            */
/* ------------------------------------------------------------------------------------------ */
typedef struct { uint8_t lastRank, lastQRank; } side_t;

typedef struct {                   /* SearchNg26.h:41-52 */
    cursor_t cur;
    side_t side[2];
    uint64_t e, part, partitionEntryValue, queryPosL, queryPosR;
    char LInfo, RInfo;
    int Right, NextPos;
    uint64_t key;                  /* discovery-order key of the path (test aid, see ng26_key_edge) */
} state_t;

typedef struct {
    const fmo_index* ix;
    int edit;
    const uint8_t* query;
    uint32_t n_parts;
    const uint32_t *pi, *l, *u;     /* current search */
    const uint32_t* partition;
    uint32_t first_symb;            /* BiFMIndex.h:26 FirstSymb = 1 for delimited indices */
    /* delegate state (search_n_impl, SearchNg26.h:411-421) */
    uint64_t qidx, ct, max_hits;
    hitvec* hv;
    fmo_counters* ctr;
    /* discovery-order keys (fmo_search_ng26_keys) */
    uint32_t search_idx, key_slots, key_bits, key_maxd, key_ords;
    uint64_t *keys, keys_n, keys_cap;
    int want_keys;
} ng26_ctx;

/* Discovery-order key -- TEST AID for the device's hit-limited search (fmb_search_scheme_n).
 * The device enumerates the search tree in an arbitrary order and must afterwards put the hits of a query into
 * the order this depth-first search finds them in.  Two paths of one search split at ONE node, so their order is
 * decided by the order in which that node visits its children:
 *   search_next_dir        (:170-218)  match, then for every symbol c: deletion(c), substitution(c), then insertion
 *   search_next_dir_single (:286-363)  insertion, then match / deletion  or  substitution / deletion
 * A path is therefore identified by its error edges.  Slot `e` of the key (e = errors before the edge) holds a code
 * for the edge taken at depth d = steps + e (strictly increasing along a path):
 *   insertion at a single-row node ("before the match child")   code = d
 *   no (further) error edge                                        code = MID = maxd
 *   every other error edge ("after the match child")               code = MID + 1 + (maxd-1-d) * ords + ord
 * with ord = 0 substitution / 1 deletion at a single-row node and 2c deletion(c) / 2c+1 substitution(c) /
 * 2 sigma + 1 insertion at a wide node.  Comparing the slots from e = 0 upwards (after the search number in the top
 * byte) reproduces the depth-first order; fmo_search_ng26_keys returns the keys so that tests can check that they
 * ascend in the order the hits are reported. */
static uint64_t ng26_key_edge(const ng26_ctx* cx, const state_t* st, int single, int before, uint32_t ord) {
    if (!cx->want_keys) return st->key;
    uint64_t d = st->cur.steps + st->e;
    uint64_t code = before ? d : (uint64_t)cx->key_maxd + 1 + ((uint64_t)cx->key_maxd - 1 - d) * cx->key_ords + ord;
    (void)single;
    uint32_t sh = 56 - (uint32_t)(st->e + 1) * cx->key_bits;
    uint64_t mask = ((1ull << cx->key_bits) - 1) << sh;
    return (st->key & ~mask) | (code << sh);
}

static int ng26_delegate(ng26_ctx* cx, cursor_t cur, uint64_t e, uint64_t key) {
    if (cur.len + cx->ct > cx->max_hits) cur.len = cx->max_hits - cx->ct;    /* :415-417 */
    cx->ct += cur.len;
    hit_push(cx->hv, cx->qidx, cur, e);
    if (cx->want_keys) {
        if (cx->keys_n == cx->keys_cap) {
            cx->keys_cap = cx->keys_cap ? cx->keys_cap * 2 : 1024;
            cx->keys = (uint64_t*)realloc(cx->keys, cx->keys_cap * sizeof(uint64_t));
        }
        cx->keys[cx->keys_n++] = key;
    }
    return cx->ct == cx->max_hits;                                             /* :420 */
}

static int ng26_search_next(ng26_ctx* cx, const state_t* state);
static int ng26_search_next_dir(ng26_ctx* cx, const state_t* state);
static int ng26_search_next_dir_single(ng26_ctx* cx, const state_t* state);

static cursor_t ng26_extend(ng26_ctx* cx, const state_t* st, uint64_t symb) {   /* :82-88 */
    return st->Right ? ext_right(cx->ix, st->cur, symb, cx->ctr) : ext_left(cx->ix, st->cur, symb, cx->ctr);
}

static int ng26_search_next_pos(ng26_ctx* cx, state_t state) {                  /* :119-141 */
    if (state.cur.len == 0) return 0;
    if (state.NextPos) {
        if (state.Right) state.queryPosR += 1;
        else state.queryPosL -= 1;
        state.partitionEntryValue -= 1;
        if (state.partitionEntryValue == 0) {
            state.part += 1;
            if (state.part != cx->n_parts) state.partitionEntryValue = cx->partition[cx->pi[state.part]];
            return ng26_search_next(cx, &state);
        }
    }
    if (state.cur.len > 1) return ng26_search_next_dir(cx, &state);
    return ng26_search_next_dir_single(cx, &state);
}

static int ng26_search_next(ng26_ctx* cx, const state_t* state) {              /* :98-117 */
    if (state->cur.len == 0) return 0;
    if (state->part == cx->n_parts) {
        if (!cx->edit || ((state->LInfo == 'M' || state->LInfo == 'I') && (state->RInfo == 'M' || state->RInfo == 'I'))) {
            if (cx->l[cx->n_parts - 1] <= state->e && state->e <= cx->u[cx->n_parts - 1])
                return ng26_delegate(cx, state->cur, state->e, state->key);
        }
        return 0;
    }
    state_t ns = *state;
    ns.Right = (state->part == 0) || (cx->pi[state->part - 1] < cx->pi[state->part]);
    if (state->cur.len > 1) return ng26_search_next_dir(cx, &ns);
    return ng26_search_next_dir_single(cx, &ns);
}

static int ng26_search_next_dir_no_errors(ng26_ctx* cx, state_t state) {       /* :225-250 */
    uint64_t loops = state.partitionEntryValue;
    uint8_t nextSymb = 0;
    for (uint64_t i = 0; i < loops; ++i) {
        nextSymb = cx->query[state.Right ? (state.queryPosR + i) : (state.queryPosL - i)];
        state.cur = ng26_extend(cx, &state, nextSymb);
        if (state.cur.len == 0) return 0;
    }
    state.side[state.Right].lastRank = nextSymb;
    state.side[state.Right].lastQRank = nextSymb;
    state.part += 1;
    state.partitionEntryValue = 0;
    if (state.part != cx->n_parts) state.partitionEntryValue = cx->partition[cx->pi[state.part]];
    if (state.Right) { state.queryPosR += loops; state.RInfo = 'M'; }
    else { state.queryPosL -= loops; state.LInfo = 'M'; }
    return ng26_search_next(cx, &state);
}

static int ng26_search_next_dir(ng26_ctx* cx, const state_t* state) {          /* :143-224 */
    const int R = state->Right;
    const int Edit = cx->edit;
    char TInfo = R ? state->RInfo : state->LInfo;
    int Deletion = (TInfo != 'S' && TInfo != 'I') && Edit;
    int Insertion = (TInfo != 'S' && TInfo != 'D') && Edit;
    char OnMatchL = R ? state->LInfo : 'M',       OnMatchR = R ? 'M' : state->RInfo;
    char OnSubstituteL = R ? state->LInfo : 'S',  OnSubstituteR = R ? 'S' : state->RInfo;
    char OnDeletionL = R ? state->LInfo : 'D',    OnDeletionR = R ? 'D' : state->RInfo;
    char OnInsertionL = R ? state->LInfo : 'I',   OnInsertionR = R ? 'I' : state->RInfo;

    uint8_t nextSymb = cx->query[R ? state->queryPosR : state->queryPosL];
    uint64_t lpart = cx->l[state->part], upart = cx->u[state->part];

    int matchAllowed = (state->partitionEntryValue > 1 || lpart <= state->e)
                       && state->e <= upart
                       && (TInfo != 'I' || nextSymb != state->side[R].lastQRank)
                       && (TInfo != 'D' || nextSymb != state->side[R].lastRank);
    int insertionAllowed = (state->partitionEntryValue > 1 || lpart <= state->e + 1) && state->e + 1 <= upart;
    int substitutionAllowed = insertionAllowed;
    int mismatchAllowed = state->e + 1 <= upart;

    if (mismatchAllowed) {
        cursor_t cursors[256];
        ext_all(cx->ix, state->cur, R, cursors, cx->ctr);                       /* :170 */
        if (matchAllowed) {
            state_t ns = *state;
            ns.cur = cursors[nextSymb];
            ns.side[R].lastRank = nextSymb;
            ns.side[R].lastQRank = nextSymb;
            ns.LInfo = OnMatchL; ns.RInfo = OnMatchR;
            ns.NextPos = 1;
            if (ng26_search_next_pos(cx, ns)) return 1;
        }
        for (uint64_t i = cx->first_symb; i < cx->ix->sigma; ++i) {             /* :184-205 */
            state_t ns = *state;
            ns.e = state->e + 1;
            ns.cur = cursors[i];
            ns.side[R].lastRank = (uint8_t)i;
            if (Deletion) {
                ns.LInfo = OnDeletionL; ns.RInfo = OnDeletionR;
                ns.NextPos = 0;
                ns.key = ng26_key_edge(cx, state, 0, 0, (uint32_t)(2 * i));
                if (ng26_search_next_pos(cx, ns)) return 1;
            }
            if (!substitutionAllowed) continue;
            if (i == nextSymb) continue;
            ns.side[R].lastQRank = nextSymb;
            ns.LInfo = OnSubstituteL; ns.RInfo = OnSubstituteR;
            ns.NextPos = 1;
            ns.key = ng26_key_edge(cx, state, 0, 0, (uint32_t)(2 * i + 1));
            if (ng26_search_next_pos(cx, ns)) return 1;
        }
        if (Insertion && insertionAllowed) {                                    /* :207-218 */
            state_t ns = *state;
            ns.e = state->e + 1;
            ns.side[R].lastQRank = nextSymb;
            ns.LInfo = OnInsertionL; ns.RInfo = OnInsertionR;
            ns.NextPos = 1;
            ns.key = ng26_key_edge(cx, state, 0, 0, 2 * cx->ix->sigma + 1);
            if (ng26_search_next_pos(cx, ns)) return 1;
        }
    } else if (matchAllowed) {
        if (ng26_search_next_dir_no_errors(cx, *state)) return 1;
    }
    return 0;
}

static int ng26_search_next_dir_single(ng26_ctx* cx, const state_t* state) {   /* :251-365 */
    const int R = state->Right;
    const int Edit = cx->edit;
    char TInfo = R ? state->RInfo : state->LInfo;
    int Deletion = (TInfo != 'S' && TInfo != 'I') && Edit;
    int Insertion = (TInfo != 'S' && TInfo != 'D') && Edit;
    char OnMatchL = R ? state->LInfo : 'M',       OnMatchR = R ? 'M' : state->RInfo;
    char OnSubstituteL = R ? state->LInfo : 'S',  OnSubstituteR = R ? 'S' : state->RInfo;
    char OnDeletionL = R ? state->LInfo : 'D',    OnDeletionR = R ? 'D' : state->RInfo;
    char OnInsertionL = R ? state->LInfo : 'I',   OnInsertionR = R ? 'I' : state->RInfo;

    /* :267-277 symbolRight/Left (BiFMIndexCursor.h:180-190) + single-symbol extend */
    uint64_t curISymb = R ? fmo_symbol(cx->ix, 1, state->cur.lb_rev) : fmo_symbol(cx->ix, 0, state->cur.lb);
    cursor_t icursorNext = ng26_extend(cx, state, curISymb);

    uint8_t curQSymb = cx->query[R ? state->queryPosR : state->queryPosL];
    uint64_t lpart = cx->l[state->part], upart = cx->u[state->part];
    int insertionAllowed = (state->partitionEntryValue > 1 || lpart <= state->e + 1) && state->e + 1 <= upart;
    int substitutionAllowed = insertionAllowed;
    int mismatchAllowed = state->e + 1 <= upart;

    if (Insertion && insertionAllowed) {                                        /* :286-297 */
        state_t ns = *state;
        ns.e = state->e + 1;
        ns.side[R].lastQRank = curQSymb;
        ns.LInfo = OnInsertionL; ns.RInfo = OnInsertionR;
        ns.NextPos = 1;
        ns.key = ng26_key_edge(cx, state, 1, 1, 0);
        if (ng26_search_next_pos(cx, ns)) return 1;
    }
    if (curISymb < cx->first_symb) return 0;                                    /* :300-302 */

    int matchAllowed = (state->partitionEntryValue > 1 || lpart <= state->e)
                       && state->e <= upart
                       && (TInfo != 'I' || curQSymb != state->side[R].lastQRank)
                       && (TInfo != 'D' || curQSymb != state->side[R].lastRank);

    if (curISymb == curQSymb) {
        if (matchAllowed) {
            if (!mismatchAllowed) return ng26_search_next_dir_no_errors(cx, *state) ? 1 : 0;   /* :311-315 */
            state_t ns = *state;
            ns.side[R].lastRank = curQSymb;
            ns.side[R].lastQRank = curQSymb;
            ns.cur = icursorNext;
            ns.LInfo = OnMatchL; ns.RInfo = OnMatchR;
            ns.NextPos = 1;
            if (ng26_search_next_pos(cx, ns)) return 1;
        }
        if (Deletion && mismatchAllowed) {                                      /* :326-338 */
            state_t ns = *state;
            ns.e = state->e + 1;
            ns.side[R].lastRank = (uint8_t)curISymb;
            ns.cur = icursorNext;
            ns.LInfo = OnDeletionL; ns.RInfo = OnDeletionR;
            ns.NextPos = 0;
            ns.key = ng26_key_edge(cx, state, 1, 0, 1);
            if (ng26_search_next_pos(cx, ns)) return 1;
        }
    } else if (mismatchAllowed) {                                               /* :339-363 */
        state_t ns = *state;
        ns.e = state->e + 1;
        ns.side[R].lastRank = (uint8_t)curISymb;
        ns.cur = icursorNext;
        if (substitutionAllowed) {
            uint8_t saved = ns.side[R].lastQRank;      /* Restore{lastQRank, curQSymb}, Restore.h:8-27 */
            ns.side[R].lastQRank = curQSymb;
            ns.LInfo = OnSubstituteL; ns.RInfo = OnSubstituteR;
            ns.NextPos = 1;
            ns.key = ng26_key_edge(cx, state, 1, 0, 0);
            int f = ng26_search_next_pos(cx, ns);
            ns.side[R].lastQRank = saved;
            if (f) return 1;
        }
        if (Deletion) {
            ns.LInfo = OnDeletionL; ns.RInfo = OnDeletionR;
            ns.NextPos = 0;
            ns.key = ng26_key_edge(cx, state, 1, 0, 1);
            if (ng26_search_next_pos(cx, ns)) return 1;
        }
    }
    return 0;
}

static int ng26_run(ng26_ctx* cx) {                                             /* :62-79 */
    state_t st;
    memset(&st, 0, sizeof st);
    for (uint32_t i = 0; i < cx->pi[0]; ++i) {
        st.queryPosL += cx->partition[i];
        st.queryPosR += cx->partition[i];
    }
    st.queryPosL -= 1;                       /* may wrap, as in the reference (:69-72) */
    st.partitionEntryValue = cx->partition[cx->pi[0]];
    st.cur.lb = 0; st.cur.lb_rev = 0; st.cur.len = cx->ix->n; st.cur.steps = 0;   /* BiFMIndexCursor.h:28-30 */
    st.LInfo = 'M'; st.RInfo = 'M';
    if (cx->want_keys) {                     /* search number, then every slot = "no error edge" */
        st.key = (uint64_t)cx->search_idx << 56;
        for (uint32_t i = 0; i < cx->key_slots; ++i) st.key |= (uint64_t)cx->key_maxd << (56 - (i + 1) * cx->key_bits);
    }
    return ng26_search_next(cx, &st);
}

int fmo_ng26_key_layout(uint32_t sigma, uint32_t n_searches, uint32_t n_parts, const uint32_t* u, const uint32_t* partition,
                        uint32_t* slots, uint32_t* bits, uint32_t* maxd, uint32_t* ords) {
    uint32_t K = 0, total = 0;
    for (uint32_t i = 0; i < n_searches * n_parts; ++i) K = u[i] > K ? u[i] : K;
    for (uint32_t p = 0; p < n_parts; ++p) total += partition[p];
    *slots = K;
    *maxd = total + 2 * K + 2;
    *ords = 2 * sigma + 2;
    uint64_t range = (uint64_t)*maxd + 1 + (uint64_t)*maxd * *ords;
    uint32_t b = 1;
    while ((1ull << b) < range) ++b;
    *bits = b;
    return K * b <= 56;
}

uint64_t fmo_search_ng26_keys(const fmo_index* ix, const uint8_t* qsym, const uint64_t* qoff, uint64_t nq, int edit,
                              uint32_t n_searches, uint32_t n_parts, const uint32_t* pi, const uint32_t* l, const uint32_t* u,
                              const uint32_t* partition, uint64_t max_hits, fmo_hit** out, uint64_t** keys_out, fmo_counters* ctr) {
    hitvec hv = {0, 0, 0};
    uint64_t *keys = NULL, keys_n = 0, keys_cap = 0;
    *out = NULL;
    if (keys_out) *keys_out = NULL;
    if (nq == 0 || max_hits == 0) return 0;                                     /* :409-410 */
    uint32_t slots = 0, bits = 0, maxd = 0, ords = 0;
    if (keys_out && !fmo_ng26_key_layout(ix->sigma, n_searches, n_parts, u, partition, &slots, &bits, &maxd, &ords)) return 0;
    for (uint64_t q = 0; q < nq; ++q) {                                         /* :411-422 */
        ng26_ctx cx;
        memset(&cx, 0, sizeof cx);
        cx.ix = ix; cx.edit = edit; cx.query = qsym + qoff[q];
        cx.n_parts = n_parts; cx.partition = partition; cx.first_symb = 1;
        cx.qidx = q; cx.ct = 0; cx.max_hits = max_hits; cx.hv = &hv; cx.ctr = ctr;
        cx.want_keys = keys_out != NULL;
        cx.key_slots = slots; cx.key_bits = bits; cx.key_maxd = maxd; cx.key_ords = ords;
        cx.keys = keys; cx.keys_n = keys_n; cx.keys_cap = keys_cap;
        for (uint32_t s = 0; s < n_searches; ++s) {                             /* search_impl :385-390 */
            cx.pi = pi + (size_t)s * n_parts;
            cx.l = l + (size_t)s * n_parts;
            cx.u = u + (size_t)s * n_parts;
            cx.search_idx = s;
            if (ng26_run(&cx)) break;
        }
        keys = cx.keys; keys_n = cx.keys_n; keys_cap = cx.keys_cap;
    }
    *out = hv.v;
    if (keys_out) *keys_out = keys;
    return hv.n;
}

uint64_t fmo_search_ng26(const fmo_index* ix, const uint8_t* qsym, const uint64_t* qoff, uint64_t nq, int edit,
                         uint32_t n_searches, uint32_t n_parts, const uint32_t* pi, const uint32_t* l, const uint32_t* u,
                         const uint32_t* partition, uint64_t max_hits, fmo_hit** out, fmo_counters* ctr) {
    return fmo_search_ng26_keys(ix, qsym, qoff, nq, edit, n_searches, n_parts, pi, l, u, partition, max_hits, out, NULL, ctr);
}

/* ------------------------------------------------------------------------------------------ */
/* search_pseudo -- search/SearchPseudo.h:13-186: the plain recursive search over an EXPANDED     */
/* scheme (one pi / l / u entry per query symbol), Hamming (search_hm :60-98) or edit distance    */
/* without any redundancy filter (search_distance :100-165).                                      */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    const fmo_index* ix; const uint8_t* query; uint64_t L, qidx; const uint32_t *pi, *l, *u; hitvec* hv; fmo_counters* ctr;
} pseudo_ctx;

static int pseudo_right(const pseudo_ctx* cx, uint64_t pos) { return pos == 0 || cx->pi[pos - 1] < cx->pi[pos]; }   /* :43-56 */

static void pseudo_search(pseudo_ctx* cx, int edit, cursor_t cur, uint64_t e, uint64_t pos) {
    if (cur.len == 0) return;
    if (pos == cx->L) {                                                        /* :66-71, :106-111 */
        if (cx->l[pos - 1] <= e && e <= cx->u[pos - 1]) hit_push(cx->hv, cx->qidx, cur, e);
        return;
    }
    if (e > cx->u[pos]) return;
    uint8_t rank = cx->query[cx->pi[pos]];
    const int right = pseudo_right(cx, pos);
    cursor_t cursors[256];
    memset(cursors, 0, sizeof(cursor_t) * cx->ix->sigma);
    if (e + 1 <= cx->u[pos]) ext_all(cx->ix, cur, right, cursors, cx->ctr);     /* :82-86, :122-128 */
    else cursors[rank] = right ? ext_right(cx->ix, cur, rank, cx->ctr) : ext_left(cx->ix, cur, rank, cx->ctr);
    if (cx->l[pos] <= e) pseudo_search(cx, edit, cursors[rank], e, pos + 1);    /* match */
    if (cx->l[pos] <= e + 1 && e + 1 <= cx->u[pos])                             /* substitution */
        for (uint64_t i = 1; i < cx->ix->sigma; ++i)
            if (i != rank) pseudo_search(cx, edit, cursors[i], e + 1, pos + 1);
    if (!edit) return;
    if (e + 1 <= cx->u[pos])                                                    /* deletion :150-155 */
        for (uint64_t i = 1; i < cx->ix->sigma; ++i) pseudo_search(cx, edit, cursors[i], e + 1, pos);
    if (cx->l[pos] <= e + 1 && e + 1 <= cx->u[pos]) pseudo_search(cx, edit, cur, e + 1, pos + 1);   /* insertion :158-160 */
}

uint64_t fmo_search_pseudo(const fmo_index* ix, const uint8_t* qsym, const uint64_t* qoff, uint64_t nq, int edit,
                           uint32_t n_searches, uint32_t L, const uint32_t* pi, const uint32_t* l, const uint32_t* u, fmo_hit** out, fmo_counters* ctr) {
    hitvec hv = {0, 0, 0};
    *out = NULL;
    for (uint64_t q = 0; q < nq; ++q) {                                         /* :171-186 */
        if (qoff[q + 1] - qoff[q] != L) continue;                              /* the reference asserts equal lengths */
        for (uint32_t s = 0; s < n_searches; ++s) {
            pseudo_ctx cx = {ix, qsym + qoff[q], L, q, pi + (size_t)s * L, l + (size_t)s * L, u + (size_t)s * L, &hv, ctr};
            cursor_t root = {0, 0, ix->n, 0};
            pseudo_search(&cx, edit, root, 0, 0);
        }
    }
    *out = hv.v;
    return hv.n;
}

/* ------------------------------------------------------------------------------------------ */
/* search_backtracking -- search/Backtracking.h:15-98                                          */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    const fmo_index* ix; const uint8_t* query; uint64_t L, qidx, maxErrors; hitvec* hv; fmo_counters* ctr;
} bt_ctx;

static void bt_no_errors(bt_ctx* cx, cursor_t cur, uint64_t i) {               /* :66-77 */
    if (cur.len == 0) return;
    for (; i < cx->L; ++i) {
        uint8_t r = cx->query[cx->L - i - 1];
        cur = ext_left(cx->ix, cur, r, cx->ctr);
        if (cur.len == 0) return;
    }
    hit_push(cx->hv, cx->qidx, cur, cx->maxErrors);
}

static void bt_with_errors(bt_ctx* cx, uint64_t e, cursor_t cur, uint64_t i) {  /* :42-64 */
    if (cur.len == 0) return;
    if (e == cx->maxErrors) { bt_no_errors(cx, cur, i); return; }
    for (; i < cx->L; ++i) {
        uint8_t r = cx->query[cx->L - i - 1];
        cursor_t next[256];
        ext_all(cx->ix, cur, 0, next, cx->ctr);
        for (uint64_t s = 1; s < cx->ix->sigma; ++s)
            if (r != s) bt_with_errors(cx, e + 1, next[s], i + 1);
        cur = next[r];
        if (cur.len == 0) return;
    }
    hit_push(cx->hv, cx->qidx, cur, e);
}

uint64_t fmo_search_backtracking(const fmo_index* ix, const uint8_t* qsym, const uint64_t* qoff, uint64_t nq,
                                 uint32_t max_errors, fmo_hit** out, fmo_counters* ctr) {
    hitvec hv = {0, 0, 0};
    for (uint64_t q = 0; q < nq; ++q) {
        bt_ctx cx = {ix, qsym + qoff[q], qoff[q + 1] - qoff[q], q, max_errors, &hv, ctr};
        cursor_t cur = {0, 0, ix->n, 0};
        bt_with_errors(&cx, 0, cur, 0);
    }
    /* a unidirectional index has no lbRev: normalise to 0 so both index kinds compare equal */
    if (!ix->bwt[1]) for (uint64_t i = 0; i < hv.n; ++i) hv.v[i].lb_rev = 0;
    *out = hv.v;
    return hv.n;
}

/* ------------------------------------------------------------------------------------------ */
/* locate -- fmindex/BiFMIndex.h:177-206, suffixarray/SparseArray.h:63-70, locate.h:15-57       */
/* ------------------------------------------------------------------------------------------ */
int fmo_single_locate_step(const fmo_index* ix, uint64_t row, uint64_t out[2]) {
    uint64_t w = ix->sample_bitmap[row >> 6];
    if (!((w >> (row & 63)) & 1)) return 0;
    uint64_t r = ix->sample_rank[row >> 6] + (uint64_t)__builtin_popcountll(w & ((1ull << (row & 63)) - 1));
    out[0] = ix->sample_seq[r];
    out[1] = ix->sample_pos[r];
    return 1;
}

static void locate_row(const fmo_index* ix, uint64_t idx, uint64_t out[3], fmo_counters* ctr) {
    uint64_t sp[2];
    uint64_t steps = 0;
    if (ctr) ctr->locate_lookups += 1;                      /* initial marker test */
    while (!fmo_single_locate_step(ix, idx, sp)) {
        uint64_t symb = ix->bwt[0][idx];
        idx = fmo_rank(ix, 0, idx, symb) + ix->C[symb];     /* BiFMIndex.h:196-197 */
        ++steps;
        if (ctr) { ctr->lf_steps += 1; ctr->locate_lookups += 2; }
    }
    if (ctr) ctr->locate_lookups += 1;                      /* sample fetch */
    out[0] = sp[0]; out[1] = sp[1]; out[2] = steps;
}

void fmo_locate_row(const fmo_index* ix, uint64_t row, uint64_t out[3]) { locate_row(ix, row, out, NULL); }

uint64_t fmo_locate(const fmo_index* ix, const fmo_hit* hits, uint64_t nhits, fmo_loc** out, fmo_counters* ctr) {
    uint64_t total = 0;
    for (uint64_t i = 0; i < nhits; ++i) total += hits[i].len;
    fmo_loc* v = (fmo_loc*)malloc((total ? total : 1) * sizeof(fmo_loc));
    uint64_t k = 0;
    for (uint64_t i = 0; i < nhits; ++i) {
        for (uint64_t row = hits[i].lb; row < hits[i].lb + hits[i].len; ++row) {   /* locate.h:46-56 */
            uint64_t r[3];
            locate_row(ix, row, r, ctr);
            fmo_loc o = {hits[i].qidx, r[0], r[1] + r[2], hits[i].e};                /* search.h:56-58 */
            v[k++] = o;
        }
    }
    *out = v;
    return total;
}
