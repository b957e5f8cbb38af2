/*
 * oracle/kmer_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into, imported by, or called from the product).
 *
 * Plain-C CPU restatement of what `kmer-db build` + `kmer-db all2all-sp` + `kmer-db distance ani-shorter`
 * compute for vclust's prefilter (paths relative to /root/reference/3rd_party/kmer-db/src/):
 *   kmo_extract      <- KmerHelper::extract            kmer_extract.h:13-96  (+ MinHashFilter filter.h:33-116)
 *   kmo_sort_unique  <- sort + std::unique              console_build.cpp:94-103 (set size = "total-kmers")
 *   kmo_common       <- PrefixKmerDb + all2all_sp       similarity_calculator.cpp:442-657, stated as what it
 *                       computes: common[i][j] = |distinct k-mers of i  INTERSECT  distinct k-mers of j|, i > j
 *   kmo_ani_shorter  <- metric lambda                   params.cpp:28-32
 * The inverted index / pattern machinery of the reference is an implementation device and is deliberately
 * not restated; a global sort of (k-mer, genome) tuples gives the same integers.
 *
 * Parity status: PINNED by tests/test_oracle_kmer.py against example/output/fltr.txt, the total-kmers /
 * common KATs recorded in SURVEY.md 8(c) (k=25 f=1, k=25 f=0.2, k=15 f=1), kmer-db's own test/synth KAT
 * (k=21) and outputs of the unmodified reference binary oracle/_ref/kmer-db on seeded synthetic genomes.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

/* alphabet.h:80-85 ("A,C,G,TU", case-insensitive); everything else is invalid */
static int nt_code(unsigned char c)
{
    switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': case 'U': case 'u': return 3;
    default: return -1;
    }
}

static uint64_t fmix64(uint64_t k)
{
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}

/* filter.h:96-115 */
uint64_t kmo_minhash(uint64_t kmer, int k)
{
    uint64_t c = (uint64_t)ceil((double)k / 4);
    uint64_t h = kmer * 0x87c37b91114253d5ULL;
    h = (h << 31) | (h >> 33);
    h *= 0x4cf5ad432745937fULL;
    uint64_t h1 = 42 ^ h; h1 ^= c;
    uint64_t h2 = 42 ^ c;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2; h2 += h1;
    return h1 ^ h2;
}

/* filter.h:42-43 with startValue = 0 */
uint64_t kmo_threshold(double fraction)
{
    return (uint64_t)((double)UINT64_MAX * (0.0 + fraction));
}

/*
 * All canonical k-mers of one sequence record (a window holding a non-ACGTU byte yields nothing).
 * For k < 20 the value is re-encoded as (v << s) | (v & ((1<<s)-1)), s = 8 - (2k - 32)  (kmer_extract.h:38-45,88);
 * the hash filter sees the re-encoded value.  fraction >= 1 keeps everything (filter.h:136-146).
 * Returns the number written to out (capacity must be >= n).
 */
size_t kmo_extract(const char *seq, size_t n, int k, double fraction, uint64_t *out)
{
    if (n < (size_t)k) return 0;
    const uint64_t mask = (k == 32) ? ~0ULL : ((1ULL << (2 * k)) - 1);
    const int top = 2 * (k - 1);
    int shift = 0; uint64_t tail = 0;
    if (2 * k - 32 < 8) { shift = 8 - (2 * k - 32); tail = (1ULL << shift) - 1; }
    const int use_filter = fraction < 1.0;
    const uint64_t thr = kmo_threshold(fraction);
    uint64_t fwd = 0, rc = 0;
    size_t cnt = 0, valid = 0;
    for (size_t i = 0; i < n; ++i) {
        int s = nt_code((unsigned char)seq[i]);
        if (s < 0) { valid = 0; s = 0; } else ++valid;
        fwd = ((fwd << 2) | (uint64_t)s) & mask;
        rc = (rc >> 2) | ((uint64_t)(3 - s) << top);
        if (valid >= (size_t)k) {
            uint64_t can = fwd < rc ? fwd : rc;
            can = (can << shift) | (can & tail);
            if (!use_filter || kmo_minhash(can, k) < thr) out[cnt++] = can;
        }
    }
    return cnt;
}

static int cmp_u64(const void *a, const void *b)
{
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return (x > y) - (x < y);
}

size_t kmo_sort_unique(uint64_t *v, size_t n)
{
    if (!n) return 0;
    qsort(v, n, sizeof(uint64_t), cmp_u64);
    size_t w = 1;
    for (size_t i = 1; i < n; ++i) if (v[i] != v[w - 1]) v[w++] = v[i];
    return w;
}

typedef struct { uint64_t kmer; uint32_t gid; } tuple_t;
static int cmp_tuple(const void *a, const void *b)
{
    const tuple_t *x = (const tuple_t *)a, *y = (const tuple_t *)b;
    if (x->kmer != y->kmer) return x->kmer < y->kmer ? -1 : 1;
    return (x->gid > y->gid) - (x->gid < y->gid);
}
static int cmp_pair(const void *a, const void *b) { return cmp_u64(a, b); }

/*
 * Sparse lower-triangular matrix of common distinct k-mers.
 * kmers: concatenation of the per-genome sorted-unique sets, off[g]..off[g+1] delimit genome g.
 * Output (malloc'd, caller frees with kmo_free): rows[], cols[] (row > col), common[]; sorted by (row, col).
 */
int64_t kmo_common(const uint64_t *kmers, const int64_t *off, int n_genomes,
                   uint32_t **rows, uint32_t **cols, uint32_t **common)
{
    int64_t total = off[n_genomes];
    tuple_t *t = (tuple_t *)malloc(sizeof(tuple_t) * (size_t)(total ? total : 1));
    for (int g = 0; g < n_genomes; ++g)
        for (int64_t i = off[g]; i < off[g + 1]; ++i) { t[i].kmer = kmers[i]; t[i].gid = (uint32_t)g; }
    qsort(t, (size_t)total, sizeof(tuple_t), cmp_tuple);
    /* pass 1: count pair increments */
    int64_t n_inc = 0;
    for (int64_t a = 0; a < total;) {
        int64_t b = a; while (b < total && t[b].kmer == t[a].kmer) ++b;
        int64_t m = b - a; n_inc += m * (m - 1) / 2; a = b;
    }
    uint64_t *pairs = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(n_inc ? n_inc : 1));
    int64_t w = 0;
    for (int64_t a = 0; a < total;) {
        int64_t b = a; while (b < total && t[b].kmer == t[a].kmer) ++b;
        for (int64_t i = a; i < b; ++i)
            for (int64_t j = a; j < i; ++j)
                pairs[w++] = ((uint64_t)t[i].gid << 32) | t[j].gid;      /* row = larger id */
        a = b;
    }
    free(t);
    qsort(pairs, (size_t)n_inc, sizeof(uint64_t), cmp_pair);
    int64_t n_out = 0;
    for (int64_t a = 0; a < n_inc;) { int64_t b = a; while (b < n_inc && pairs[b] == pairs[a]) ++b; ++n_out; a = b; }
    *rows = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n_out ? n_out : 1));
    *cols = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n_out ? n_out : 1));
    *common = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n_out ? n_out : 1));
    int64_t o = 0;
    for (int64_t a = 0; a < n_inc;) {
        int64_t b = a; while (b < n_inc && pairs[b] == pairs[a]) ++b;
        (*rows)[o] = (uint32_t)(pairs[a] >> 32); (*cols)[o] = (uint32_t)pairs[a]; (*common)[o] = (uint32_t)(b - a);
        ++o; a = b;
    }
    free(pairs);
    return n_out;
}

void kmo_free(void *p) { free(p); }

/* params.cpp:28-32 */
double kmo_ani_shorter(uint32_t common, uint32_t cnt1, uint32_t cnt2, int k)
{
    double j = (double)common / (cnt1 < cnt2 ? cnt1 : cnt2);
    double d = (j == 0) ? 1.0 : (-1.0 / k) * log((2 * j) / (j + 1));
    return 1.0 - d;
}
