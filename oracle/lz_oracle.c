/*
 * oracle/lz_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into, imported by, or called from the product).
 *
 * A plain-C, single-threaded CPU restatement of the LZ-ANI pairwise parse: the algorithm that
 * `lz-ani all2all` runs for one (reference, query) pair.  It follows the reference sources cited below
 * (paths relative to /root/reference/3rd_party/lz-ani/src/), but is written from the algorithm's
 * description, not transcribed: containers are flat C arrays, the long-anchor index is exposed as
 * "all positions of an 11-mer in ascending order", and the factor list is an explicit growable array.
 *
 * Parity status: PINNED.  tests/test_oracle_lz.py checks this file against
 *   - example/output/ani.tsv        (132 directed pairs, golden committed under tests/golden/),
 *   - 3rd_party/lz-ani/test/vir61.ani.tsv (3660 directed pairs, LZ-ANI's own CI gate),
 *   - outputs of the unmodified reference binary oracle/_ref/lz-ani on seeded synthetic genomes.
 *
 * Symbol codes (defs.h:24-30): A0 C1 G2 T3; a stored "other" symbol is 5 (code_N_seq); inside the
 * reference text every N becomes 4 (code_N_ref) so that an N never equals an N.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

typedef struct {
    int mal, msl, mrd, mqd, reg, aw, am, ar;   /* params.h:38-45 defaults 11 7 40 40 35 15 7 3 */
} lzo_params;

enum { F_MATCH_CLOSE = 1, F_MATCH_DISTANT = 2, F_RUN_LITERALS = 4 };

typedef struct { int data_pos, flag, offset, len; } lzo_factor;   /* defs.h:32-46 */

typedef struct {
    int ref_start, ref_end, seq_start, seq_end, num_matches, num_mismatches;   /* defs.h:67-148 */
} lzo_region;

typedef struct {
    lzo_params p;
    /* reference side (parser.cpp:16-34) */
    uint8_t *R; int nR;
    int oob;                                 /* the last parse read outside R (undefined in the reference) */
    int64_t *kR_long, *kR_short;           /* k-mer code at every position or -1 */
    int *ht_long; uint32_t ht_long_size, ht_long_mask;
    int *hs_start, *hs_cnt, *hs_pos;       /* CSR over 4^msl short seeds */
    /* query side (parser.cpp:37-50) */
    uint8_t *Q; int nQ, capQ;
    int64_t *kQ_long, *kQ_short;
    /* factor list */
    lzo_factor *f; int nf, capf;
    /* scratch */
    int *left_cnt, *right_cnt; uint8_t *left_is, *right_is; int cap_side;
    int *window;
} lzo_ctx;

/* parser.h:98-107 */
static uint64_t hash_mm(uint64_t x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

/* parser.cpp:53-103.  out[p] = packed code of seq[p..p+len) when that window holds only ACGT, else -1. */
static void kmers_of(const uint8_t *seq, int n, int len, int64_t *out)
{
    uint64_t mask = (~0ULL) >> (64 - 2 * len);
    uint64_t k = 0;
    int valid = 0;                      /* number of consecutive ACGT symbols ending here */
    for (int i = 0; i < n; ++i) {
        uint8_t c = seq[i];
        k = ((k << 2) + (c & 3)) & mask;
        valid = (c >= 4) ? 0 : valid + 1;
        if (i + 1 >= len)
            out[i + 1 - len] = (valid >= len) ? (int64_t)k : -1;
    }
    for (int p = (n - len + 1 > 0 ? n - len + 1 : 0); p < n; ++p)
        out[p] = -1;
}

static void push_factor(lzo_ctx *c, int data_pos, int flag, int offset, int len)
{
    if (c->nf == c->capf) {
        c->capf = c->capf ? 2 * c->capf : 1024;
        c->f = (lzo_factor *)realloc(c->f, sizeof(lzo_factor) * (size_t)c->capf);
    }
    lzo_factor x = { data_pos, flag, offset, len };
    c->f[c->nf++] = x;
}

lzo_ctx *lzo_create(const lzo_params *p)
{
    lzo_ctx *c = (lzo_ctx *)calloc(1, sizeof(lzo_ctx));
    c->p = *p;
    c->window = (int *)calloc((size_t)(p->aw > 0 ? p->aw : 1), sizeof(int));
    return c;
}

static void free_ref(lzo_ctx *c)
{
    free(c->R); free(c->kR_long); free(c->kR_short); free(c->ht_long);
    free(c->hs_start); free(c->hs_cnt); free(c->hs_pos);
    c->R = NULL; c->kR_long = c->kR_short = NULL; c->ht_long = NULL;
    c->hs_start = c->hs_cnt = c->hs_pos = NULL;
}

void lzo_destroy(lzo_ctx *c)
{
    if (!c) return;
    free_ref(c);
    free(c->Q); free(c->kQ_long); free(c->kQ_short); free(c->f);
    free(c->left_cnt); free(c->right_cnt); free(c->left_is); free(c->right_is); free(c->window);
    free(c);
}

/* parser.cpp:16-34 -- codes[] uses the reservoir convention (0..3, 5 = other). */
void lzo_set_reference(lzo_ctx *c, const uint8_t *codes, int len)
{
    const lzo_params *p = &c->p;
    free_ref(c);
    int n = 2 * len + 3 * p->mrd;
    c->nR = n;
    c->R = (uint8_t *)malloc((size_t)n + 1);
    uint8_t *w = c->R;
    for (int i = 0; i < len; ++i) *w++ = codes[i] >= 4 ? 4 : codes[i];
    for (int i = 0; i < 2 * p->mrd; ++i) *w++ = 4;
    for (int i = len - 1; i >= 0; --i) *w++ = codes[i] >= 4 ? 4 : (uint8_t)(3 - codes[i]);
    for (int i = 0; i < p->mrd; ++i) *w++ = 4;

    c->kR_long = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + 1));
    c->kR_short = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + 1));
    kmers_of(c->R, n, p->mal, c->kR_long);
    kmers_of(c->R, n, p->msl, c->kR_short);

    /* parser.cpp:146-189: open addressing, size 2*floor_pow2(n/0.1), linear probing, ascending insertion */
    uint32_t x = (uint32_t)((double)n / 0.1);
    while (x & (x - 1)) x &= x - 1;
    c->ht_long_size = 2 * x;
    c->ht_long_mask = c->ht_long_size - 1;
    c->ht_long = (int *)malloc(sizeof(int) * (size_t)c->ht_long_size);
    for (uint32_t i = 0; i < c->ht_long_size; ++i) c->ht_long[i] = -1;
    for (int i = 0; i < n; ++i) {
        if (c->kR_long[i] < 0) continue;
        uint32_t h = (uint32_t)(hash_mm((uint64_t)c->kR_long[i]) & c->ht_long_mask);
        while (c->ht_long[h] != -1) h = (h + 1) & c->ht_long_mask;
        c->ht_long[h] = i;
    }

    /* parser.cpp:106-143: counting sort of positions by short seed */
    int nb = 1 << (2 * p->msl);
    c->hs_start = (int *)calloc((size_t)nb + 1, sizeof(int));
    c->hs_cnt = (int *)calloc((size_t)nb, sizeof(int));
    c->hs_pos = (int *)malloc(sizeof(int) * (size_t)(n + 1));
    for (int i = 0; i < n; ++i) if (c->kR_short[i] >= 0) c->hs_cnt[c->kR_short[i]]++;
    for (int b = 0; b < nb; ++b) c->hs_start[b + 1] = c->hs_start[b] + c->hs_cnt[b];
    int *fill = (int *)malloc(sizeof(int) * (size_t)nb);
    memcpy(fill, c->hs_start, sizeof(int) * (size_t)nb);
    for (int i = 0; i < n; ++i) if (c->kR_short[i] >= 0) c->hs_pos[fill[c->kR_short[i]]++] = i;
    free(fill);
}

/* parser.cpp:37-50 */
static void set_query(lzo_ctx *c, const uint8_t *codes, int len)
{
    int n = len + c->p.mrd;
    if (n + 1 > c->capQ) {
        c->capQ = n + 1;
        c->Q = (uint8_t *)realloc(c->Q, (size_t)c->capQ);
        c->kQ_long = (int64_t *)realloc(c->kQ_long, sizeof(int64_t) * (size_t)c->capQ);
        c->kQ_short = (int64_t *)realloc(c->kQ_short, sizeof(int64_t) * (size_t)c->capQ);
    }
    c->nQ = n;
    for (int i = 0; i < len; ++i) c->Q[i] = codes[i] >= 4 ? 5 : codes[i];
    for (int i = len; i < n; ++i) c->Q[i] = 5;
    kmers_of(c->Q, n, c->p.msl, c->kQ_short);
    kmers_of(c->Q, n, c->p.mal, c->kQ_long);
}

/* parser.cpp:192-207 */
static int equal_len(const lzo_ctx *c, int rp, int qp, int start)
{
    int max_r = c->nR - rp < c->nQ - qp ? c->nR - rp : c->nQ - qp;
    int r = start;
    while (r < max_r && c->R[rp + r] == c->Q[qp + r]) ++r;
    return r;
}

/* The reference's compare_ranges has no bounds check, and with mqd > mrd its tail call (parser.cpp:713) can run past
 * seq_ref.size(): what it then compares against is whatever the vector's capacity holds (undefined; thread-schedule
 * dependent in lz-ani).  The oracle reads such positions as a symbol that matches nothing and raises c->oob, so that a
 * test can tell "defined by the reference" from "undefined in the reference". */
static int ref_sym(lzo_ctx *c, int idx)
{
    if (idx < 0 || idx >= c->nR) { c->oob = 1; return 0xff; }
    return c->R[idx];
}

/* parser.cpp:210-248: run-length encode the match/mismatch string of Q[d..d+len) against R[r..r+len) */
static void compare_ranges(lzo_ctx *c, int d, int r, int len, int backward)
{
    int flag = backward ? F_MATCH_DISTANT : F_MATCH_CLOSE;
    int j = 0;
    while (j < len) {
        int m = ref_sym(c, r + j) == c->Q[d + j];
        int s = j;
        while (j < len && (ref_sym(c, r + j) == c->Q[d + j]) == m) ++j;
        if (m) {
            push_factor(c, d + s, flag, r + s, j - s);
            if (j < len) flag = F_MATCH_CLOSE;
        } else
            push_factor(c, d + s, F_RUN_LITERALS, 0, j - s);
    }
}

static void emit_sym(lzo_ctx *c, int first, int data_p, int is_match, int ref_p)
{
    int fl = is_match ? F_MATCH_CLOSE : F_RUN_LITERALS;
    if (!first && c->f[c->nf - 1].flag == fl)
        c->f[c->nf - 1].len++;
    else
        push_factor(c, data_p, fl, is_match ? ref_p : 0, 1);
}

/* parser.cpp:251-374: fill the literal gap before a close match by the best left/right split */
static void compare_ranges_both_ways(lzo_ctx *c, int d, int r_left, int r_end_right, int len)
{
    int to_scan = (r_end_right < r_left) ? len : (r_end_right - r_left < len ? r_end_right - r_left : len);
    if (to_scan + 2 > c->cap_side) {
        c->cap_side = to_scan + 64;
        c->left_cnt = (int *)realloc(c->left_cnt, sizeof(int) * (size_t)c->cap_side);
        c->right_cnt = (int *)realloc(c->right_cnt, sizeof(int) * (size_t)c->cap_side);
        c->left_is = (uint8_t *)realloc(c->left_is, (size_t)c->cap_side);
        c->right_is = (uint8_t *)realloc(c->right_is, (size_t)c->cap_side);
    }
    int *L = c->left_cnt, *Rr = c->right_cnt; uint8_t *Li = c->left_is, *Ri = c->right_is;
    L[0] = 0; Li[0] = 0;
    for (int i = 0; i < to_scan; ++i) {
        Li[i + 1] = c->R[r_left + i] == c->Q[d + i];
        L[i + 1] = L[i] + Li[i + 1];
    }
    for (int i = 0; i <= to_scan; ++i) { Rr[i] = 0; Ri[i] = 0; }          /* :300 resize(...,(0,false)) */
    int lim = to_scan < r_end_right ? to_scan : r_end_right;
    for (int i = 1; i <= lim; ++i) {
        Ri[i] = c->R[r_end_right - i] == c->Q[d + len - i];
        Rr[i] = Rr[i - 1] + Ri[i];
    }
    int best = 0, best_split = 0;
    for (int i = 0; i <= to_scan; ++i) {
        int v = L[i] + Rr[to_scan - i];
        if (v >= best) { best = v; best_split = i; }                     /* :308 ties -> largest split */
    }
    int nf0 = c->nf;                                                      /* "first factor of this call" marker */
    /* left part */
    for (int i = 1; i <= best_split; ++i)
        emit_sym(c, i == 1, d + i - 1, Li[i], r_left + i - 1);
    /* middle */
    if (to_scan < len) {
        if (best_split > 0 && c->f[c->nf - 1].flag == F_RUN_LITERALS)
            c->f[c->nf - 1].len += len - to_scan;
        else
            push_factor(c, d + best_split, F_RUN_LITERALS, 0, len - to_scan);
    }
    /* right part */
    if (best_split < to_scan) {
        int shift = len - to_scan;
        int from_right = to_scan - best_split;
        int data_p = d + best_split + shift;
        int m = Ri[from_right];
        if (!m && (best_split > 0 || shift > 0) && c->f[c->nf - 1].flag == F_RUN_LITERALS)
            c->f[c->nf - 1].len++;                                       /* :355-356 (data_p not advanced) */
        else {
            push_factor(c, data_p, m ? F_MATCH_CLOSE : F_RUN_LITERALS, m ? r_end_right - from_right : 0, 1);
            data_p++;
        }
        for (int i = from_right - 1; i > 0; --i, ++data_p)
            emit_sym(c, 0, data_p, Ri[i], r_end_right - i);
    }
    (void)nf0;
}

/* parser.cpp:377-409 */
static int try_extend_forward(lzo_ctx *c, int d, int r)
{
    const lzo_params *p = &c->p;
    int nmm = 0, last = 0, run = p->ar, e;
    memset(c->window, 0, sizeof(int) * (size_t)p->aw);
    for (e = 0; d + e < c->nQ && r + e < c->nR; ++e) {
        int mm = c->Q[d + e] != c->R[r + e];
        nmm -= c->window[e % p->aw];
        c->window[e % p->aw] = mm;
        nmm += mm;
        if (!mm) { if (++run >= p->ar) last = e + 1; }
        else run = 0;
        if (nmm > p->am) break;
    }
    return last;
}

/* parser.cpp:412-441 */
static int try_extend_backward(lzo_ctx *c, int d, int r, int max_len)
{
    const lzo_params *p = &c->p;
    int nmm = 0, last = 0, run = p->ar, e;
    memset(c->window, 0, sizeof(int) * (size_t)p->aw);
    for (e = 0; d - e > 0 && r - e > 0 && e < max_len; ++e) {
        int mm = c->Q[d - e - 1] != c->R[r - e - 1];
        nmm -= c->window[e % p->aw];
        c->window[e % p->aw] = mm;
        nmm += mm;
        if (!mm) { if (++run >= p->ar) last = e + 1; }
        else run = 0;
        if (nmm > p->am) break;
    }
    return last;
}

/* parser.h:134-172 */
static double prob_len(int len)
{
    if (len < 30) return ldexp(1.0, -2 * len);      /* the table holds exact powers of 4 */
    return pow(4, -len);
}

/* parser.h:174-188 -- note the unsigned exponent (negative ints wrap, quirk 4 of SURVEY 8(a)) */
static double ipow(double base, uint32_t e)
{
    double r = 1.0;
    while (e) { if (e & 1) r *= base; base *= base; e /= 2; }
    return r;
}

/* Long-anchor lookup, parser.cpp:514-531 / :585-602: walk the probe chain; strict '>' keeps the first. */
static void anchor_search(const lzo_ctx *c, int i, int *best_len, int *best_pos)
{
    *best_len = 0; *best_pos = 0;
    if (c->kQ_long[i] < 0) return;
    uint32_t h = (uint32_t)(hash_mm((uint64_t)c->kQ_long[i]) & c->ht_long_mask);
    for (; c->ht_long[h] != -1; h = (h + 1) & c->ht_long_mask) {
        int ml = equal_len(c, c->ht_long[h], i, 0);
        if (ml < c->p.mal) continue;
        if (ml > *best_len) { *best_len = ml; *best_pos = c->ht_long[h]; }
    }
}

static int iabs(int x) { return x < 0 ? -x : x; }

/* parser.cpp:482-716 */
static void parse(lzo_ctx *c)
{
    const lzo_params *p = &c->p;
    int nQ = c->nQ;
    int pred = -nQ;             /* ref_pred_pos; <0 == lost */
    int lit = 0;                /* cur_lit_run_len */
    int prev_start = -1, prev_end = 0;
    int i = 0;
    c->nf = 0;

    while (i + p->msl < nQ) {
        int best_pos = 0, best_len = 0;
        if (pred < 0)
            anchor_search(c, i, &best_len, &best_pos);
        else {
            int64_t h = c->kQ_short[i];
            if (h >= 0) {
                const int *bucket = c->hs_pos + c->hs_start[h];
                int bs = c->hs_cnt[h];
                int lo = 0, hi = bs, key = pred - lit;                   /* lower_bound :555 */
                while (lo < hi) { int mid = (lo + hi) >> 1; if (bucket[mid] < key) lo = mid + 1; else hi = mid; }
                for (int j = lo; j < bs && bucket[j] < pred + p->mrd; ++j) {
                    int pos = bucket[j];
                    int ml = equal_len(c, pos, i, p->msl);
                    if (ml >= best_len) {
                        if (ml == best_len) {
                            if (iabs(pos - pred) < iabs(best_pos - pred)) best_pos = pos;
                        } else { best_len = ml; best_pos = pos; }
                    }
                }
            }
            int a_len, a_pos;
            anchor_search(c, i, &a_len, &a_pos);
            if (a_pos) {                                                  /* position used as boolean :604 */
                if (!best_pos) { best_pos = a_pos; best_len = a_len; }
                else {
                    double anchor_prob = ipow(1 - prob_len(a_len), (uint32_t)(int)(2 * ((size_t)c->nR + 1 - (size_t)a_len)));
                    double close_prob = ipow(1 - prob_len(best_len), (uint32_t)(lit + p->mrd + 1 - best_len));
                    if (anchor_prob > close_prob) { best_pos = a_pos; best_len = a_len; }
                }
            }
        }

        if (best_len >= p->msl) {
            int flag = F_MATCH_DISTANT;
            if (pred >= 0 && iabs(best_pos - pred) <= p->mrd) {
                compare_ranges_both_ways(c, i - lit, pred - lit, best_pos + best_len, lit);
                push_factor(c, i, F_MATCH_CLOSE, best_pos, best_len);
            } else {
                if (lit) push_factor(c, i - lit, F_RUN_LITERALS, 0, lit);
                if (prev_start >= 0 && !(prev_end - prev_start >= p->reg)) {     /* eval_region :446-449 */
                    while (c->nf && c->f[c->nf - 1].data_pos >= prev_start) c->nf--;
                    int run_len = i - prev_start;
                    while (c->nf && c->f[c->nf - 1].flag == F_RUN_LITERALS) { run_len += c->f[c->nf - 1].len; c->nf--; }
                    push_factor(c, i - run_len, F_RUN_LITERALS, 0, run_len);
                    prev_start = -1;
                }
                if (c->nf && c->f[c->nf - 1].flag == F_RUN_LITERALS) {
                    int back = try_extend_backward(c, i, best_pos, c->f[c->nf - 1].len);
                    if (back) {
                        c->f[c->nf - 1].len -= back;
                        if (c->f[c->nf - 1].len == 0) c->nf--;
                        compare_ranges(c, i - back, best_pos - back, back, 1);
                        flag = F_MATCH_CLOSE;
                        prev_start = i - back;
                    }
                }
                push_factor(c, i, flag, best_pos, best_len);
                if (flag == F_MATCH_DISTANT) prev_start = i;
                if (prev_start < 0)                                       /* :678-684 */
                    for (int j = c->nf - 1; j >= 0; --j)
                        if (c->f[j].flag == F_MATCH_DISTANT) { prev_start = c->f[j].data_pos; break; }
            }
            i += best_len;
            pred = best_pos + best_len;
            lit = 0;
            int ext = try_extend_forward(c, i, pred);
            compare_ranges(c, i, pred, ext, 0);
            i += ext;
            pred += ext;
            prev_end = i;
        } else {
            ++i; ++pred; ++lit;
        }
        if (lit > p->mqd) pred = -nQ;
    }

    if (pred < 0)
        push_factor(c, i - lit, F_RUN_LITERALS, 0, lit + (nQ - i));
    else
        compare_ranges(c, i - lit, pred - lit - p->msl, lit + (nQ - i), 0);     /* :713 (sic) */
}

/* parser.cpp:734-783 */
static void calc_stats(const lzo_ctx *c, int out[3])
{
    int cur_len = 0, cur_lit = 0, n_lit = 0;
    int sm = 0, sl = 0, nc = 0;
    for (int k = 0; k <= c->nf; ++k) {
        int flush = (k == c->nf) || c->f[k].flag == F_MATCH_DISTANT;
        if (flush) {
            if (cur_len && cur_len + cur_lit >= c->p.reg) { sm += cur_len; sl += cur_lit; ++nc; }
            if (k == c->nf) break;
            cur_len = c->f[k].len; cur_lit = 0; n_lit = 0;
        } else if (c->f[k].flag == F_MATCH_CLOSE) {
            cur_len += c->f[k].len; cur_lit += n_lit; n_lit = 0;
        } else
            n_lit += c->f[k].len;
    }
    out[0] = sm; out[1] = sl; out[2] = nc;
}

static int region_cmp(const void *a, const void *b)
{
    const lzo_region *x = (const lzo_region *)a, *y = (const lzo_region *)b;
    int lx = x->seq_end - x->seq_start, ly = y->seq_end - y->seq_start;
    if (lx != ly) return lx > ly ? -1 : 1;
    return (x->seq_start > y->seq_start) - (x->seq_start < y->seq_start);
}

static void upd_min(int *v, int x) { if (*v < 0 || x < *v) *v = x; }
static void upd_max(int *v, int x) { if (*v < 0 || x > *v) *v = x; }

/* parser.cpp:786-837.  Returns the number of regions written (at most cap); sorted like the reference. */
static int calc_regions(const lzo_ctx *c, lzo_region *out, int cap)
{
    int n = 0, buf_lit = 0;
    lzo_region cur = { -1, -1, -1, -1, 0, 0 };
    for (int k = 0; k < c->nf; ++k) {
        const lzo_factor *x = &c->f[k];
        if (x->flag == F_MATCH_DISTANT) {
            if (cur.seq_end - cur.seq_start >= c->p.reg && n < cap) out[n++] = cur;
            lzo_region z = { -1, -1, -1, -1, 0, 0 };
            cur = z; buf_lit = 0;
        } else if (x->flag == F_MATCH_CLOSE) {
            cur.ref_end += buf_lit; cur.seq_end += buf_lit; cur.num_mismatches += buf_lit; buf_lit = 0;
        } else { buf_lit += x->len; continue; }
        upd_min(&cur.seq_start, x->data_pos); upd_max(&cur.seq_end, x->data_pos + x->len);
        upd_min(&cur.ref_start, x->offset);   upd_max(&cur.ref_end, x->offset + x->len);
        cur.num_matches += x->len;
    }
    if (cur.seq_end - cur.seq_start >= c->p.reg && n < cap) out[n++] = cur;
    qsort(out, (size_t)n, sizeof(lzo_region), region_cmp);
    return n;
}

/* One directed pair: query (reservoir codes) against the reference set by lzo_set_reference. */
void lzo_parse_query(lzo_ctx *c, const uint8_t *qcodes, int qlen, int stats[3])
{
    set_query(c, qcodes, qlen);
    c->oob = 0;
    parse(c);
    calc_stats(c, stats);
}
/* 1 when the last parse compared against positions outside the reference text (undefined in the reference, see ref_sym) */
int lzo_last_oob(const lzo_ctx *c) { return c->oob; }

int lzo_last_regions(const lzo_ctx *c, lzo_region *out, int cap) { return calc_regions(c, out, cap); }
int lzo_last_factors(const lzo_ctx *c, const lzo_factor **out) { *out = c->f; return c->nf; }

/* Batch driver used by tests/bench: pairs (ref id, query id) grouped by reference for index reuse.
 * seqs: concatenated reservoir codes; off[g]..off[g+1] delimit genome g.  stats: 3 ints per pair. */
void lzo_run_pairs_oob(const lzo_params *p, const uint8_t *seqs, const int64_t *off,
                       const int32_t *pair_ref, const int32_t *pair_qry, int64_t n_pairs, int32_t *stats, uint8_t *oob);
void lzo_run_pairs(const lzo_params *p, const uint8_t *seqs, const int64_t *off,
                   const int32_t *pair_ref, const int32_t *pair_qry, int64_t n_pairs, int32_t *stats)
{
    lzo_run_pairs_oob(p, seqs, off, pair_ref, pair_qry, n_pairs, stats, 0);
}
/* oob (optional): per pair, 1 when the parse left the reference text (see ref_sym) */
void lzo_run_pairs_oob(const lzo_params *p, const uint8_t *seqs, const int64_t *off,
                       const int32_t *pair_ref, const int32_t *pair_qry, int64_t n_pairs, int32_t *stats, uint8_t *oob)
{
    lzo_ctx *c = lzo_create(p);
    int cur_ref = -1;
    for (int64_t k = 0; k < n_pairs; ++k) {
        int r = pair_ref[k], q = pair_qry[k];
        if (r != cur_ref) { lzo_set_reference(c, seqs + off[r], (int)(off[r + 1] - off[r])); cur_ref = r; }
        int st[3];
        lzo_parse_query(c, seqs + off[q], (int)(off[q + 1] - off[q]), st);
        stats[3 * k] = st[0]; stats[3 * k + 1] = st[1]; stats[3 * k + 2] = st[2];
        if (oob) oob[k] = (uint8_t)c->oob;
    }
    lzo_destroy(c);
}
