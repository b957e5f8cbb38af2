// vb_prefilter: the all-vs-all shared-k-mer screen (kmer-db build + all2all-sp + the ani-shorter filter) on one B200.
//
// Reference computation (paths under /root/reference/3rd_party/kmer-db/src/):
//   k-mer extraction   kmer_extract.h:13-96, MinHash threshold filter filter.h:33-146
//   per-genome set     console_build.cpp:94-103 (sort + unique; set size = total-kmers, kmer_db.h:129)
//   common counts      prefix_kmer_db.cpp:244-434 + similarity_calculator.cpp:442-657
//   pair filter        sparse_filters.h:12-61 with metric params.cpp:28-32
// Pipeline here (all integer work, HBM-bound, no tensor cores):
//   k1 extract_kernel   packed genomes -> one (canonical k-mer, genome id) tuple per base slot, in genome order
//   k2 rsort::sort_kv   stable LSD radix sort by k-mer  => every k-mer's genome ids ascending, duplicates adjacent
//   k3 segment kernels  runs of equal k-mers -> duplicate counts per genome, pair increments into an HBM hash table
//   k4 emit kernels     table -> (row, col, common) passing the integer filter and a conservative ani test
// The exact IEEE-double ani-shorter test and the text formatting run on the host (libm log() must match glibc's).
#include <algorithm>
#include <chrono>
#include <cmath>

#include "dev_util.cuh"
#include "radix_sort.cuh"

namespace {

constexpr uint64_t KEY_SENTINEL = ~0ULL;
constexpr uint64_t SLOT_EMPTY = ~0ULL;

__device__ __forceinline__ uint64_t fmix64(uint64_t k)
{
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}

// kmer-db filter.h:96-115 (MurmurHash3-style 128-bit finalisation folded to 64 bits); c = ceil(k/4)
__device__ __forceinline__ uint64_t minhash64(uint64_t kmer, uint64_t c)
{
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

// reverse the order of the 32 two-bit digits of x
__device__ __forceinline__ uint64_t reverse_digits(uint64_t x)
{
    x = __brevll(x);
    return ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1);
}

struct ExtractParams {
    int k;
    int shift;              // re-encoding for k < 20 (kmer_extract.h:38-45,88)
    uint64_t tail_mask;
    int use_filter;
    uint64_t max_thr;       // filter.h:42-43
    uint64_t c;             // ceil(k/4)
    uint32_t shard_index;   // multi-GPU: this rank keeps the k-mers with fmix64(kmer) % shard_count == shard_index
    uint32_t shard_count;
};

// k1: one thread per base slot; a warp covers 32 consecutive slots of one 128-slot tile (= one genome).
__global__ void __launch_bounds__(256) extract_kernel(const uint32_t *__restrict__ seq2, const uint32_t *__restrict__ inv,
                                                      const uint32_t *__restrict__ tile_gid, uint64_t n_slots,
                                                      uint64_t n_out, ExtractParams ep, uint64_t *__restrict__ keys,
                                                      uint32_t *__restrict__ vals, uint32_t *__restrict__ valid_cnt)
{
    const int k = ep.k;
    const uint64_t kmask = (~0ULL) >> (64 - 2 * k);
    const uint32_t wmask = (k >= 32) ? 0xffffffffu : ((1u << k) - 1);
    // n_out is a multiple of 32 and so is the thread count: whole warps enter and leave the loop together
    for (uint64_t p = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; p < n_out; p += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t key = KEY_SENTINEL;
        uint32_t gid = 0xffffffffu;
        if (p < n_slots) {
            gid = tile_gid[p >> 7];
            if (gid != 0xffffffffu && (fetch1(inv, p) & wmask) == 0) {
                uint64_t w = fetch2(seq2, p) & kmask;            // digit j = base p+j  (little-endian window)
                uint64_t rc = (~w) & kmask;                      // == reference's kmer_rev as an integer
                uint64_t fw = reverse_digits(w) >> (64 - 2 * k); // == reference's kmer_str (first base most significant)
                uint64_t can = fw < rc ? fw : rc;
                can = (can << ep.shift) | (can & ep.tail_mask);
                bool keep = !ep.use_filter || minhash64(can, ep.c) < ep.max_thr;
                if (keep && ep.shard_count > 1) keep = (fmix64(can) % ep.shard_count) == ep.shard_index;
                if (keep) key = can;
            }
        }
        keys[p] = key;
        vals[p] = gid;
        unsigned ok = __ballot_sync(0xffffffffu, key != KEY_SENTINEL);
        // all active lanes of a warp share the 128-slot tile, hence the genome
        if (ok && (threadIdx.x & 31) == (__ffs(ok) - 1)) atomicAdd(&valid_cnt[gid], (uint32_t)__popc(ok));
    }
}

// k3a (only for very large N): number of pair increments (sum over k-mer runs of m*(m-1)/2), to size the table
__global__ void __launch_bounds__(256) segment_count_kernel(const uint64_t *__restrict__ keys,
                                                            const uint32_t *__restrict__ vals, uint64_t n,
                                                            unsigned long long *__restrict__ n_inc)
{
    unsigned long long local = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t key = keys[i];
        if (key == KEY_SENTINEL) continue;
        uint32_t g = vals[i];
        if (i > 0 && keys[i - 1] == key && vals[i - 1] == g) continue;
        // rank of this genome inside the run = number of distinct genomes before it
        uint32_t prev = g;
        for (uint64_t j = i; j-- > 0 && keys[j] == key;) {
            uint32_t gj = vals[j];
            if (gj != prev) { ++local; prev = gj; }
        }
    }
    for (int o = 16; o; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(n_inc, local);
}

// Accumulator of pair increments.  Two layouts:
//   dense  (N(N-1)/2 <= 2^26): one uint32 counter per pair at index row*(row-1)/2 + col -- an increment is a single
//          fire-and-forget atomic, and an ordered scan of the array yields the pairs already sorted by (row, col);
//   hashed (larger N): open addressing, uint64 key (row << 32 | col) + uint32 count.
struct PairAcc {
    uint32_t *dense;        // non-null selects the dense layout
    uint64_t *tkeys;
    uint32_t *tvals;
    uint64_t cap_mask;
    int *overflow;
};

__device__ __forceinline__ void table_add(const PairAcc &A, uint64_t key, uint32_t inc)
{
    uint64_t h = fmix64(key) & A.cap_mask;
    for (uint64_t probes = 0; probes <= A.cap_mask; ++probes) {
        uint64_t cur = A.tkeys[h];
        if (cur == SLOT_EMPTY) {
            cur = atomicCAS((unsigned long long *)&A.tkeys[h], (unsigned long long)SLOT_EMPTY, (unsigned long long)key);
            if (cur == SLOT_EMPTY) cur = key;
        }
        if (cur == key) { atomicAdd(&A.tvals[h], inc); return; }
        h = (h + 1) & A.cap_mask;
    }
    *A.overflow = 1;
}

// one increment for the pair (hi, lo), hi > lo
__device__ __forceinline__ void pair_add(const PairAcc &A, uint32_t hi, uint32_t lo)
{
    if (A.dense) atomicAdd(&A.dense[(uint64_t)hi * (hi - 1) / 2 + lo], 1u);
    else table_add(A, ((uint64_t)hi << 32) | lo, 1u);
}

// k3b: duplicates per genome; every distinct (k-mer, genome) occurrence pairs with the distinct genomes before it in the run;
// row = the later (larger) genome id, col = the earlier one -- the lower triangle of all2all_sp.
__global__ void __launch_bounds__(256) segment_pairs_kernel(const uint64_t *__restrict__ keys,
                                                            const uint32_t *__restrict__ vals, uint64_t n,
                                                            uint32_t *__restrict__ dup_cnt, PairAcc A)
{
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t key = keys[i];
        if (key == KEY_SENTINEL) continue;
        uint32_t g = vals[i];
        if (i > 0 && keys[i - 1] == key && vals[i - 1] == g) { atomicAdd(&dup_cnt[g], 1u); continue; }   // same k-mer twice in g
        uint32_t prev = g;
        for (uint64_t j = i; j-- > 0 && keys[j] == key;) {
            uint32_t gj = vals[j];
            if (gj != prev) {
                pair_add(A, g, gj);
                prev = gj;
            }
        }
    }
}


// ===============================================================================================================
// MSD path (default): group equal k-mers WITHOUT a full sort.
//   Only the grouping of equal k-mers matters, not their numeric order, so tuples are keyed by h = fmix64(k-mer)
//   (a bijection on 64 bits): the top b1 + b2 bits of h select one of B1*B2 buckets of ~1000 tuples; a bucket is then
//   grouped entirely in shared memory (hash table keyed by h) and the pairs are emitted from there.
//     count_kernel   genomes -> bucket histogram (+ valid k-mers per genome)                       0.375 B/base read
//     scan_kernel    exclusive scan -> bucket offsets, write cursors, tile map
//     part_kernel<A> genomes -> level-1 buckets (extraction fused, coalesced runs via smem staging)  12 B/tuple written
//     part_kernel<B> level-1 -> level-2 buckets                                                     12 B read + 12 B written
//     bucket_kernel  level-2 bucket -> smem grouping -> duplicates per genome + pair increments      12 B read
//   = 48 B of HBM traffic per tuple instead of 7 radix passes x 32 B.
// ===============================================================================================================
constexpr int PART_THREADS = 512;
constexpr int PART_ITEMS = 8;
constexpr int PART_TILE = PART_THREADS * PART_ITEMS;      // 4096 tuples per block
constexpr int MAX_BUCKET_BITS = 10;                       // per level
constexpr int BUCKET_CAP = 2048;                          // tuples a bucket may hold to be grouped in shared memory
constexpr int BUCKET_SLOTS = 4096;                        // hash slots per bucket (load <= 0.5)
constexpr int BUCKET_TARGET = 1280;                       // planned mean bucket size (CAP is 20 sigma above it)

struct MsdPlan {
    int b1, b2;              // bucket bits of level 1 and level 2 (b2 == 0: one level)
    uint32_t B1, B2, NB;     // 1 << b1, 1 << b2, B1 * B2
};

// Singleton screen.  A k-mer that occurs once in the whole input shares nothing and is the common case (92 % of the
// distinct k-mers of c2), so it is dropped before the partition: a table of 2-bit slots indexed by hash bits records
// "seen" (bit 0) and "seen again" (bit 1); only tuples whose slot has bit 1 go on.  Every occurrence of a k-mer maps to
// the same slot, so a k-mer with two or more occurrences (in any genomes, or twice in one genome) always survives;
// a singleton survives only when it collides with another k-mer (harmless).  The table is sized to stay L2-resident.
struct SeenTable {
    uint32_t *words;        // nullptr: screen disabled, everything survives
    uint64_t slot_mask;     // slots - 1 (power of two); 16 slots per word
};

__device__ __forceinline__ bool seen_twice(const SeenTable &T, uint64_t h)
{
    if (!T.words) return true;
    const uint64_t s = (h >> 16) & T.slot_mask;
    return (__ldg(T.words + (s >> 4)) >> (2 * (uint32_t)(s & 15) + 1)) & 1u;
}

constexpr int FINE_BITS_SMALL = 18, FINE_BITS_LARGE = 2 * MAX_BUCKET_BITS;   // resolution of the survivor histogram

// Append this block's surviving tuples (bit r of `keep` selects h[r]; all of one thread's tuples share a genome) to the
// compact list -- order is irrelevant, tuples are grouped by hash later -- and count them in the fine histogram.  One
// global cursor atomic per call and block.  All threads of the block must call; `phase` alternates 0/1 between successive
// calls so that a call never overwrites prefixes another warp is still reading.  The survivors are compacted in shared
// memory first and leave with fully coalesced stores (c3, screen off: 6.8 -> 4.5 ms).  write == 0: only count.
template <int ITEMS>
struct AppendSmem {
    uint32_t warp[2][33];                  // warp totals / prefixes, two phases
    uint32_t base[2];                      // the block's reservation in the list
    uint64_t keys[256 * ITEMS];            // the block's survivors of this call, compacted: written to the list with
    uint32_t vals[256 * ITEMS];            // fully coalesced stores (a thread's own tuples would be 64-byte-strided)
};

template <int ITEMS>
__device__ __forceinline__ void block_append(uint32_t keep, const uint64_t (&h)[ITEMS], uint32_t gid, AppendSmem<ITEMS> &S,
                                             int phase, unsigned long long *__restrict__ cursor, int write,
                                             uint64_t *__restrict__ out_keys, uint32_t *__restrict__ out_vals,
                                             uint32_t *__restrict__ fine_hist, int fine_bits)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    uint32_t *sw = S.warp[phase];
    const uint32_t cnt = (uint32_t)__popc(keep);
    uint32_t x = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) sw[wid] = x;
    __syncthreads();
    if (wid == 0) {
        uint32_t v = lane < nw ? sw[lane] : 0, z = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, z, o); if (lane >= o) z += y; }
        const uint32_t total = __shfl_sync(0xffffffffu, z, 31);
        unsigned long long base = 0;
        if (lane == 0 && total) base = atomicAdd(cursor, (unsigned long long)total);
        if (lane == 0) S.base[phase] = (uint32_t)base;         // the list holds fewer than 2^32 tuples: the low word is enough
        if (lane < nw) sw[lane] = z - v;                       // block-local offset of every warp
        if (lane == 31) sw[32] = total;
    }
    __syncthreads();
    if (!write) return;
    uint32_t o = sw[wid] + x - cnt;
#pragma unroll
    for (int r = 0; r < ITEMS; ++r)
        if ((keep >> r) & 1u) {
            S.keys[o] = h[r];
            S.vals[o] = gid;
            ++o;
            atomicAdd(&fine_hist[(uint32_t)(h[r] >> (64 - fine_bits))], 1u);
        }
    __syncthreads();
    const uint32_t total = sw[32], base = S.base[phase];
    for (uint32_t j = threadIdx.x; j < total; j += blockDim.x) {
        out_keys[base + j] = S.keys[j];
        out_vals[base + j] = S.vals[j];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// k1.  One thread takes KM_ITEMS consecutive base slots (p a multiple of KM_ITEMS): three words of bases and two words
// of validity bits are loaded once and every k-mer window is cut out of them with funnel shifts, so a k-mer costs
// ~50 instructions instead of ~150 with one thread (and five loads) per slot.  h = fmix64(canonical k-mer) is a
// bijection; only the grouping of equal k-mers matters downstream, not their numeric order.
// Nothing is stored per slot: the screen pass only marks the seen table, the collect pass RECOMPUTES the hashes from
// the packed genomes (0.375 B/base) -- 8 B/slot written and read back cost more than hashing twice.
// ---------------------------------------------------------------------------------------------------------------
constexpr int KM_ITEMS = 8;

__device__ __forceinline__ uint32_t kmer_hashes(const uint32_t *__restrict__ seq2, const uint32_t *__restrict__ inv, uint64_t p,
                                                const ExtractParams &ep, uint64_t kmask, uint32_t wmask, uint64_t (&h)[KM_ITEMS])
{
    const uint32_t *q = seq2 + (p >> 4);
    const uint32_t a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    const uint32_t *qi = inv + (p >> 5);
    const uint32_t i0 = __ldg(qi), i1 = __ldg(qi + 1);
    const unsigned ib = (unsigned)(p & 31);                                  // 0, 8, 16 or 24
    const uint32_t iv0 = __funnelshift_r(i0, i1, ib), iv1 = i1 >> ib;        // >= 40 validity bits from slot p on
    const unsigned o = (unsigned)(p & 15) * 2;                               // 0 or 16
    const uint32_t w0 = __funnelshift_r(a, b, o), w1 = __funnelshift_r(b, c, o), w2 = c >> o;   // >= 40 bases from slot p on
    uint32_t ok = 0;
#pragma unroll
    for (int j = 0; j < KM_ITEMS; ++j) {
        const uint32_t bad = __funnelshift_r(iv0, iv1, j) & wmask;
        const uint32_t lo = __funnelshift_r(w0, w1, 2 * j), hi = __funnelshift_r(w1, w2, 2 * j);
        const uint64_t w = (((uint64_t)hi << 32) | lo) & kmask;              // digit t = base p + j + t
        const uint64_t rc = (~w) & kmask;                                    // == reference's kmer_rev as an integer
        const uint64_t fw = reverse_digits(w) >> (64 - 2 * ep.k);            // == reference's kmer_str
        uint64_t can = fw < rc ? fw : rc;
        can = (can << ep.shift) | (can & ep.tail_mask);
        bool keep = bad == 0;
        if (ep.use_filter) keep = keep && minhash64(can, ep.c) < ep.max_thr;
        const uint64_t hh = fmix64(can);
        if (ep.shard_count > 1) keep = keep && (hh % ep.shard_count) == ep.shard_index;
        h[j] = hh;
        ok |= (uint32_t)keep << j;
    }
    return ok;
}

// valid k-mers per genome: a warp covers 32 * KM_ITEMS = 256 consecutive slots = two 128-slot tiles = at most two genomes
__device__ __forceinline__ void count_valid(uint32_t ok, uint32_t gid, int lane, uint32_t *__restrict__ valid_cnt)
{
    const uint32_t c = gid == 0xffffffffu ? 0u : (uint32_t)__popc(ok);
    const uint32_t total = __reduce_add_sync(0xffffffffu, c);
    if (!total) return;
    const uint32_t lo = __reduce_add_sync(0xffffffffu, lane < 16 ? c : 0u);
    const uint32_t g_lo = __shfl_sync(0xffffffffu, gid, 0), g_hi = __shfl_sync(0xffffffffu, gid, 16);
    if (lane == 0 && lo) atomicAdd(&valid_cnt[g_lo], lo);
    if (lane == 16 && total - lo) atomicAdd(&valid_cnt[g_hi], total - lo);
}

// Singleton screen, pass 1 over the slots [p_lo, p_hi) (multiples of 256): mark every k-mer in the seen table; *n_again
// counts the tuples that found their slot already marked (survivors = *n_again + number of slots with bit 1).
__global__ void __launch_bounds__(256) screen_kernel(const uint32_t *__restrict__ seq2, const uint32_t *__restrict__ inv,
                                                     const uint32_t *__restrict__ tile_gid, uint64_t p_lo, uint64_t p_hi,
                                                     ExtractParams ep, SeenTable T, uint32_t *__restrict__ valid_cnt,
                                                     unsigned long long *__restrict__ n_again)
{
    const uint64_t kmask = (~0ULL) >> (64 - 2 * ep.k);
    const uint32_t wmask = (ep.k >= 32) ? 0xffffffffu : ((1u << ep.k) - 1);
    const int lane = threadIdx.x & 31;
    const uint64_t n_groups = (p_hi - p_lo) / KM_ITEMS;
    uint32_t again = 0;
    // whole warps enter and leave the loop together (count_valid synchronises the warp)
    for (uint64_t gi = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; gi - lane < n_groups; gi += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t p = p_lo + gi * KM_ITEMS;
        const uint32_t gid = gi < n_groups ? tile_gid[p >> 7] : 0xffffffffu;
        uint64_t h[KM_ITEMS];
        uint32_t ok = 0;
        if (gid != 0xffffffffu) ok = kmer_hashes(seq2, inv, p, ep, kmask, wmask, h);
        else {
#pragma unroll
            for (int j = 0; j < KM_ITEMS; ++j) h[j] = 0;
        }
        // all eight marks are issued before the first result is looked at (eight atomics in flight per thread instead
        // of one: the kernel is bound by the round trip of the atomic, not by its throughput)
        uint32_t old[KM_ITEMS], widx[KM_ITEMS];
#pragma unroll
        for (int j = 0; j < KM_ITEMS; ++j) {
            const uint64_t s = (h[j] >> 16) & T.slot_mask;
            widx[j] = (uint32_t)(s >> 4);
            old[j] = 0;
            if ((ok >> j) & 1u) old[j] = atomicOr(T.words + widx[j], 1u << (2 * ((uint32_t)(h[j] >> 16) & 15)));
        }
#pragma unroll
        for (int j = 0; j < KM_ITEMS; ++j) {
            const uint32_t b0 = 1u << (2 * ((uint32_t)(h[j] >> 16) & 15));
            if (old[j] & b0) {                                                // (old is 0 for slots without a k-mer)
                ++again;
                if (!(old[j] & (b0 << 1))) atomicOr(T.words + widx[j], b0 << 1);   // only the second occurrence pays for this one
            }
        }
        count_valid(ok, gid, lane, valid_cnt);
    }
    again = __reduce_add_sync(0xffffffffu, again);
    if (lane == 0 && again) atomicAdd(n_again, (unsigned long long)again);
}

// Pass 2 (or the only pass when the screen is off): recompute the hashes, keep the tuples whose seen slot says "twice",
// append them to the compact list (write == 0: only count them).  COUNT_VALID: also count valid k-mers per genome.
template <bool COUNT_VALID>
__global__ void __launch_bounds__(256) collect_kernel(const uint32_t *__restrict__ seq2, const uint32_t *__restrict__ inv,
                                                      const uint32_t *__restrict__ tile_gid, uint64_t p_lo, uint64_t p_hi,
                                                      ExtractParams ep, SeenTable T, uint32_t *__restrict__ valid_cnt,
                                                      unsigned long long *__restrict__ cursor, int write,
                                                      uint64_t *__restrict__ out_keys, uint32_t *__restrict__ out_vals,
                                                      uint32_t *__restrict__ fine_hist, int fine_bits)
{
    __shared__ AppendSmem<KM_ITEMS> s_app;
    const uint64_t kmask = (~0ULL) >> (64 - 2 * ep.k);
    const uint32_t wmask = (ep.k >= 32) ? 0xffffffffu : ((1u << ep.k) - 1);
    const int lane = threadIdx.x & 31;
    const uint64_t n_groups = (p_hi - p_lo) / KM_ITEMS;
    int phase = 0;
    // whole blocks enter and leave the loop together (block_append synchronises the block)
    for (uint64_t base = blockIdx.x * (uint64_t)blockDim.x; base < n_groups; base += (uint64_t)gridDim.x * blockDim.x, phase ^= 1) {
        const uint64_t p = p_lo + (base + threadIdx.x) * KM_ITEMS;
        const uint32_t gid = base + threadIdx.x < n_groups ? tile_gid[p >> 7] : 0xffffffffu;
        uint64_t h[KM_ITEMS];
        uint32_t ok = 0;
        if (gid != 0xffffffffu) ok = kmer_hashes(seq2, inv, p, ep, kmask, wmask, h);
        else {
#pragma unroll
            for (int j = 0; j < KM_ITEMS; ++j) h[j] = 0;
        }
        if (COUNT_VALID) count_valid(ok, gid, lane, valid_cnt);
        if (T.words) {                                                       // eight independent table reads in flight
            uint32_t tw[KM_ITEMS];
#pragma unroll
            for (int j = 0; j < KM_ITEMS; ++j) tw[j] = __ldg(T.words + (((h[j] >> 16) & T.slot_mask) >> 4));
#pragma unroll
            for (int j = 0; j < KM_ITEMS; ++j)
                if (!((tw[j] >> (2 * ((uint32_t)(h[j] >> 16) & 15) + 1)) & 1u)) ok &= ~(1u << j);
        }
        block_append<KM_ITEMS>(ok, h, gid, s_app, phase, cursor, write, out_keys, out_vals, fine_hist, fine_bits);
    }
}

// coarse bucket histogram (NB = 2^total_bits bins) from the fine one
__global__ void __launch_bounds__(256) coarsen_kernel(const uint32_t *__restrict__ fine_hist, int fine_bits, int total_bits,
                                                      uint32_t *__restrict__ hist)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (1u << fine_bits)) return;
    const uint32_t c = fine_hist[i];
    if (c) atomicAdd(&hist[total_bits ? (i >> (fine_bits - total_bits)) : 0u], c);
}

// number of slots with bit 1 set (= distinct slots that hold a repeated k-mer); survivors = *n_again + that number
__global__ void __launch_bounds__(256) seen_popc_kernel(const uint32_t *__restrict__ words, uint64_t n_words,
                                                        unsigned long long *__restrict__ n_slots_again)
{
    uint32_t c = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_words; i += (uint64_t)gridDim.x * blockDim.x)
        c += __popc(words[i] & 0xaaaaaaaau);
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(n_slots_again, (unsigned long long)c);
}

// exclusive scan of one value per thread over a 1024-thread block; returns the prefix, *total gets the sum
__device__ __forceinline__ uint32_t block_exscan_1024(uint32_t v, uint32_t *warp_tot /* [32] shared */, uint32_t *total)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) warp_tot[w] = x;
    __syncthreads();
    if (w == 0) {
        uint32_t t = warp_tot[lane], z = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, z, o); if (lane >= o) z += y; }
        warp_tot[lane] = z - t;
        if (lane == 31) *total = z;
    }
    __syncthreads();
    uint32_t pre = warp_tot[w] + x - v;
    __syncthreads();
    return pre;
}

// one block: off[i] = exclusive prefix of hist (NB + 1 entries); cursor2 = off; cursor1[b] = off[b * B2];
// tile_start[b] = number of PART_TILE tiles in level-1 buckets < b (B1 + 1 entries)
__global__ void __launch_bounds__(1024) scan_kernel(const uint32_t *__restrict__ hist, MsdPlan pl, uint32_t *__restrict__ off,
                                                    uint32_t *__restrict__ cursor1, uint32_t *__restrict__ cursor2,
                                                    uint32_t *__restrict__ tile_start)
{
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t s_total;
    const uint32_t per = (pl.NB + 1023) / 1024;
    const uint32_t lo = min(threadIdx.x * per, pl.NB), hi = min(lo + per, pl.NB);
    uint32_t sum = 0;
    for (uint32_t i = lo; i < hi; ++i) sum += hist[i];
    uint32_t run = block_exscan_1024(sum, warp_tot, &s_total);
    for (uint32_t i = lo; i < hi; ++i) { off[i] = run; cursor2[i] = run; run += hist[i]; }
    if (threadIdx.x == 0) off[pl.NB] = s_total;
    __syncthreads();                    // block-wide visibility of the off[] writes (the same block reads them below)
    uint32_t tiles = 0, beg = 0;
    if (threadIdx.x < pl.B1) {          // B1 <= 1024
        beg = off[threadIdx.x * pl.B2];
        uint32_t end = off[(threadIdx.x + 1) * pl.B2];
        tiles = (end - beg + PART_TILE - 1) / PART_TILE;
    }
    uint32_t tpre = block_exscan_1024(tiles, warp_tot, &s_total);
    if (threadIdx.x < pl.B1) { cursor1[threadIdx.x] = beg; tile_start[threadIdx.x] = tpre; }
    if (threadIdx.x == 0) tile_start[pl.B1] = s_total;
}

// Partition one tile of tuples into buckets.  LEVEL 1: tuples come from the compact survivor list, bucket = top b1
// bits of h.  LEVEL 2: tuples come from a level-1 bucket, bucket = next b2 bits.  Inside the block the tile is first
// grouped by bucket in shared memory, so that every bucket receives one contiguous run per tile.
template <int LEVEL>
__global__ void __launch_bounds__(PART_THREADS, 3) part_kernel(uint64_t n_in, MsdPlan pl, const uint64_t *__restrict__ in_keys,
                                                            const uint32_t *__restrict__ in_vals, const uint32_t *__restrict__ off,
                                                            const uint32_t *__restrict__ tile_start, uint32_t n_tiles,
                                                            uint32_t *__restrict__ cursor, uint64_t *__restrict__ out_keys,
                                                            uint32_t *__restrict__ out_vals)
{
    extern __shared__ unsigned char smem_raw[];
    uint64_t *st_keys = (uint64_t *)smem_raw;                          // PART_TILE
    uint32_t *st_vals = (uint32_t *)(st_keys + PART_TILE);              // PART_TILE
    constexpr int MAXB = 1 << MAX_BUCKET_BITS;
    uint32_t *s_cnt = st_vals + PART_TILE;                              // MAXB each
    uint32_t *s_start = s_cnt + MAXB;
    uint32_t *s_fill = s_start + MAXB;
    uint32_t *s_gbase = s_fill + MAXB;
    __shared__ uint32_t s_total, s_bucket, s_tile_lo;
    const uint32_t NBK = (LEVEL == 1) ? pl.B1 : pl.B2;
    const int shift = (LEVEL == 1) ? (64 - pl.b1) : (64 - pl.b1 - pl.b2);

    if (LEVEL == 2) n_tiles = tile_start[pl.B1];
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (uint32_t b = threadIdx.x; b < NBK; b += PART_THREADS) { s_cnt[b] = 0; s_fill[b] = 0; }
        uint64_t src_lo = 0, src_hi = 0;
        uint32_t parent = 0;
        if (LEVEL == 2) {
            if (threadIdx.x == 0) {                                     // which level-1 bucket does this tile belong to?
                uint32_t lo = 0, hi = pl.B1;
                while (hi - lo > 1) { uint32_t mid = (lo + hi) / 2; if (tile_start[mid] <= tile) lo = mid; else hi = mid; }
                s_bucket = lo; s_tile_lo = tile_start[lo];
            }
        }
        __syncthreads();
        if (LEVEL == 2) {
            parent = s_bucket;
            uint64_t beg = off[parent * pl.B2], end = off[(parent + 1) * pl.B2];
            src_lo = beg + (uint64_t)(tile - s_tile_lo) * PART_TILE;
            src_hi = min(src_lo + PART_TILE, end);
        } else {
            src_lo = (uint64_t)tile * PART_TILE;
            src_hi = min(src_lo + PART_TILE, n_in);
        }
        uint64_t key[PART_ITEMS];
        uint32_t val[PART_ITEMS];
        uint32_t bk[PART_ITEMS];
#pragma unroll
        for (int r = 0; r < PART_ITEMS; ++r) {
            uint64_t p = src_lo + (uint64_t)r * PART_THREADS + threadIdx.x;
            const bool ok = p < src_hi;
            if (ok) { key[r] = in_keys[p]; val[r] = in_vals[p]; }
            bk[r] = 0xffffffffu;
            if (ok) {
                bk[r] = (NBK > 1) ? (uint32_t)((key[r] >> shift) & (NBK - 1)) : 0u;
                atomicAdd(&s_cnt[bk[r]], 1u);
            }
        }
        __syncthreads();
        // exclusive scan of s_cnt (NBK <= 1024) + one global reservation per non-empty bucket
        if (threadIdx.x < 32) {
            uint32_t carry = 0;
            for (uint32_t base = 0; base < NBK; base += 32) {
                uint32_t b = base + threadIdx.x;
                uint32_t v = b < NBK ? s_cnt[b] : 0, x = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (threadIdx.x >= o) x += y; }
                if (b < NBK) s_start[b] = carry + x - v;
                carry += __shfl_sync(0xffffffffu, x, 31);
            }
            if (threadIdx.x == 0) s_total = carry;
        }
        __syncthreads();
        for (uint32_t b = threadIdx.x; b < NBK; b += PART_THREADS) {
            uint32_t c = s_cnt[b];
            if (c) s_gbase[b] = atomicAdd(&cursor[(LEVEL == 1) ? b : parent * pl.B2 + b], c);
        }
#pragma unroll
        for (int r = 0; r < PART_ITEMS; ++r) {
            if (bk[r] != 0xffffffffu) {
                uint32_t pos = s_start[bk[r]] + atomicAdd(&s_fill[bk[r]], 1u);
                st_keys[pos] = key[r];
                st_vals[pos] = val[r];
            }
        }
        __syncthreads();
        const uint32_t total = s_total;
        for (uint32_t s = threadIdx.x; s < total; s += PART_THREADS) {
            uint64_t k = st_keys[s];
            uint32_t b = (NBK > 1) ? (uint32_t)((k >> shift) & (NBK - 1)) : 0u;
            uint32_t dst = s_gbase[b] + (s - s_start[b]);
            out_keys[dst] = k;
            out_vals[dst] = st_vals[s];
        }
        __syncthreads();
    }
}

// shared-memory layout of bucket_kernel (dynamic, 44 KB -> 5 blocks per SM)
constexpr uint32_t CHAIN_END = 0x7fffu;        // prev[]: low 15 bits = previous tuple with the same k-mer
constexpr uint32_t CHAIN_DUP = 0x8000u;        // prev[]: this tuple repeats a genome that is already in its chain
struct BucketSmem {
    uint64_t keys[BUCKET_CAP];                 // h of every tuple
    uint32_t gids[BUCKET_CAP];
    uint32_t table[BUCKET_SLOTS];              // slot -> the most recently inserted tuple with the slot's key
    uint16_t prev[BUCKET_CAP];
};

__device__ __forceinline__ void emit_pair(uint32_t a, uint32_t b, int count_only, unsigned long long &local_inc, const PairAcc &A)
{
    if (count_only) { ++local_inc; return; }
    pair_add(A, a > b ? a : b, a > b ? b : a);
}

// One block per final bucket.  Equal k-mers are linked into chains through a shared-memory hash table: a tuple finds the
// slot of its key and swaps itself in as the slot's newest member, keeping the previous one as its predecessor.  A
// tuple's chain is then exactly the set of tuples with the same k-mer inserted before it, so walking it enumerates every
// unordered pair of the group once -- no sort, no regrouping.  A tuple whose genome already occurs in its chain is a
// within-genome duplicate (counted in dup_cnt, skipped by everybody else).
// count_only: only sum the number of pair increments (sizing pass for very large N).
__global__ void __launch_bounds__(256) bucket_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals,
                                                     const uint32_t *__restrict__ off, uint32_t n_buckets, int count_only,
                                                     uint32_t *__restrict__ dup_cnt, PairAcc A,
                                                     uint32_t *__restrict__ big_list, uint32_t *__restrict__ n_big,
                                                     unsigned long long *__restrict__ n_inc)
{
    extern __shared__ unsigned char smem_raw[];
    BucketSmem &S = *(BucketSmem *)smem_raw;
    const int tid = threadIdx.x, lane = tid & 31;
    unsigned long long local_inc = 0;
    for (uint32_t bkt = blockIdx.x; bkt < n_buckets; bkt += gridDim.x) {
        const uint32_t beg = off[bkt], size = off[bkt + 1] - beg;
        if (size < 2) continue;                                              // uniform for the block
        if (size > BUCKET_CAP) {                                             // too big for shared memory: generic path
            if (tid == 0) big_list[atomicAdd(n_big, 1u)] = bkt;
            continue;
        }
        for (int s = tid; s < BUCKET_SLOTS; s += 256) S.table[s] = 0xffffffffu;
        for (uint32_t i = tid; i < size; i += 256) { S.keys[i] = keys[beg + i]; S.gids[i] = vals[beg + i]; }
        __syncthreads();
        // ---- insert: chain every tuple to the earlier tuples with the same key
        for (uint32_t i = tid; i < size; i += 256) {
            const uint64_t k = S.keys[i];
            uint32_t s = (uint32_t)k & (BUCKET_SLOTS - 1);                    // low bits: independent of the bucket bits
            uint32_t pv = CHAIN_END;
            for (;;) {
                uint32_t cur = S.table[s];
                if (cur == 0xffffffffu) {
                    cur = atomicCAS(&S.table[s], 0xffffffffu, i);
                    if (cur == 0xffffffffu) break;                            // first tuple with this key
                }
                if (S.keys[cur] == k) { pv = atomicExch(&S.table[s], i); break; }
                s = (s + 1) & (BUCKET_SLOTS - 1);
            }
            S.prev[i] = (uint16_t)pv;
        }
        __syncthreads();
        // ---- duplicates: same genome earlier in the chain
        uint32_t dup_mask = 0;                                                // bit t: this thread's t-th tuple is a duplicate
        int t = 0;
        for (uint32_t i = tid; i < size; i += 256, ++t) {                     // size <= BUCKET_CAP = 8 * 256
            const uint32_t g = S.gids[i];
            uint32_t j = S.prev[i];                                           // no flags are set during this phase
            while (j != CHAIN_END) {
                if (S.gids[j] == g) {
                    dup_mask |= 1u << t;
                    if (!count_only) atomicAdd(&dup_cnt[g], 1u);
                    break;
                }
                j = S.prev[j];
            }
        }
        __syncthreads();                                                      // all chain walks done: now flag the duplicates
        t = 0;
        for (uint32_t i = tid; i < size; i += 256, ++t)
            if ((dup_mask >> t) & 1u) S.prev[i] = (uint16_t)(S.prev[i] | CHAIN_DUP);
        __syncthreads();
        // ---- pair increments: every non-duplicate tuple with every non-duplicate tuple before it in its chain
        for (uint32_t i = tid; i < size; i += 256) {
            const uint32_t pi = S.prev[i];
            if (pi & CHAIN_DUP) continue;
            const uint32_t g = S.gids[i];
            uint32_t j = pi;
            while (j != CHAIN_END) {
                const uint32_t pj = S.prev[j];
                if (!(pj & CHAIN_DUP)) emit_pair(g, S.gids[j], count_only, local_inc, A);
                j = pj & CHAIN_END;
            }
        }
        __syncthreads();
    }
    if (count_only) {
        for (int o = 16; o; o >>= 1) local_inc += __shfl_down_sync(0xffffffffu, local_inc, o);
        if (lane == 0 && local_inc) atomicAdd(n_inc, local_inc);
    }
}

// Generic path for buckets that do not fit shared memory (a k-mer shared by thousands of genomes lands here): the
// block sorts the bucket in place in global memory by (h, genome) with the all-ascending bitonic network, then runs
// the same run scan as the LSD path.
__global__ void __launch_bounds__(1024) big_bucket_kernel(uint64_t *__restrict__ keys, uint32_t *__restrict__ vals,
                                                          const uint32_t *__restrict__ off, const uint32_t *__restrict__ big_list,
                                                          const uint32_t *__restrict__ n_big, int count_only,
                                                          uint32_t *__restrict__ dup_cnt, PairAcc A,
                                                          unsigned long long *__restrict__ n_inc)
{
    const uint32_t nb = *n_big;
    unsigned long long local_inc = 0;
    for (uint32_t q = blockIdx.x; q < nb; q += gridDim.x) {
        const uint32_t bkt = big_list[q];
        const uint32_t beg = off[bkt], c = off[bkt + 1] - beg;
        uint64_t *K = keys + beg;
        uint32_t *V = vals + beg;
        uint32_t n2 = 1; while (n2 < c) n2 <<= 1;
        auto cmpx = [&](uint32_t i, uint32_t l) {
            uint64_t ka = K[i], kb = K[l];
            uint32_t va = V[i], vb = V[l];
            if (ka > kb || (ka == kb && va > vb)) { K[i] = kb; K[l] = ka; V[i] = vb; V[l] = va; }
        };
        for (uint32_t k = 2; k <= n2; k <<= 1) {
            for (uint32_t i = threadIdx.x; i < n2; i += blockDim.x) { uint32_t l = i ^ (k - 1); if (l > i && l < c) cmpx(i, l); }
            __syncthreads();
            for (uint32_t j = k >> 2; j > 0; j >>= 1) {
                for (uint32_t i = threadIdx.x; i < n2; i += blockDim.x) { uint32_t l = i ^ j; if (l > i && l < c) cmpx(i, l); }
                __syncthreads();
            }
        }
        for (uint32_t i = threadIdx.x; i < c; i += blockDim.x) {
            uint64_t key = K[i];
            uint32_t g = V[i];
            if (i > 0 && K[i - 1] == key && V[i - 1] == g) { if (!count_only) atomicAdd(&dup_cnt[g], 1u); continue; }
            uint32_t prev = g;
            for (uint32_t j = i; j-- > 0 && K[j] == key;) {
                uint32_t gj = V[j];
                if (gj != prev) { emit_pair(g, gj, count_only, local_inc, A); prev = gj; }
            }
        }
        __syncthreads();
    }
    if (count_only) {
        for (int o = 16; o; o >>= 1) local_inc += __shfl_down_sync(0xffffffffu, local_inc, o);
        if ((threadIdx.x & 31) == 0 && local_inc) atomicAdd(n_inc, local_inc);
    }
}

__global__ void totals_kernel(const uint32_t *__restrict__ valid_cnt, const uint32_t *__restrict__ dup_cnt, uint32_t n,
                              uint32_t *__restrict__ totals)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) totals[i] = valid_cnt[i] - dup_cnt[i];
}

struct EmitParams {
    uint32_t min_kmers;
    double min_ident_slack;     // min_ident minus a safety margin; the exact test is redone on the host
    int k;
    int gbits;
};

__device__ __forceinline__ bool emit_pass(uint64_t key, uint32_t common, const uint32_t *__restrict__ totals,
                                          const EmitParams &ep)
{
    if (key == SLOT_EMPTY || common < ep.min_kmers || common == 0) return false;
    uint32_t r = (uint32_t)(key >> 32), c = (uint32_t)key;
    uint32_t tr = totals[r], tc = totals[c];
    double j = (double)common / (double)(tr < tc ? tr : tc);
    double d = (-1.0 / ep.k) * log((2 * j) / (j + 1));
    return (1.0 - d) >= ep.min_ident_slack;
}

// k4: pass 0 counts, pass 1 writes (compact key = row << gbits | col, value = common)
__global__ void __launch_bounds__(256) emit_kernel(const uint64_t *__restrict__ tkeys, const uint32_t *__restrict__ tvals,
                                                   uint64_t cap, const uint32_t *__restrict__ totals, EmitParams ep,
                                                   int write, unsigned long long *__restrict__ cursor,
                                                   uint64_t *__restrict__ out_keys, uint32_t *__restrict__ out_vals)
{
    // cap is a multiple of 32 (power of two >= 1024): whole warps enter and leave the loop together
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < cap; i += (uint64_t)gridDim.x * blockDim.x) {
        bool ok = false;
        uint64_t key = SLOT_EMPTY;
        uint32_t v = 0;
        if (i < cap) {
            key = tkeys[i];
            if (key != SLOT_EMPTY) { v = tvals[i]; ok = emit_pass(key, v, totals, ep); }
        }
        unsigned m = __ballot_sync(0xffffffffu, ok);
        if (!m) continue;
        int lane = threadIdx.x & 31;
        int leader = __ffs(m) - 1;
        unsigned long long base = 0;
        if (lane == leader) base = atomicAdd(cursor, (unsigned long long)__popc(m));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (ok && write) {
            unsigned long long o = base + __popc(m & ((1u << lane) - 1));
            out_keys[o] = ((uint64_t)(uint32_t)(key >> 32) << ep.gbits) | (uint32_t)key;
            out_vals[o] = v;
        }
    }
}


// k4 (dense layout): ordered compaction of the triangular counter array.  pass 0: passing entries per block;
// pass 1 (after a scan of the block counts): write them in index order = sorted by (row, col).
__device__ __forceinline__ void tri_decode(uint64_t t, uint32_t &row, uint32_t &col)
{
    uint32_t r = (uint32_t)((1.0 + sqrt(1.0 + 8.0 * (double)t)) * 0.5);
    while ((uint64_t)r * (r - 1) / 2 > t) --r;
    while ((uint64_t)(r + 1) * r / 2 <= t) ++r;
    row = r; col = (uint32_t)(t - (uint64_t)r * (r - 1) / 2);
}

__global__ void __launch_bounds__(256) dense_emit_kernel(const uint32_t *__restrict__ dense, uint64_t n_entries, uint64_t per_block,
                                                         const uint32_t *__restrict__ totals, EmitParams ep, int write,
                                                         uint32_t *__restrict__ block_cnt, uint64_t *__restrict__ out_keys,
                                                         uint32_t *__restrict__ out_vals)
{
    __shared__ uint32_t warp_cnt[8];
    __shared__ uint32_t s_run;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint64_t lo = blockIdx.x * per_block, hi = min(n_entries, lo + per_block);
    if (threadIdx.x == 0) s_run = write ? block_cnt[blockIdx.x] : 0;        // pass 1: block_cnt holds the exclusive offsets
    __syncthreads();
    for (uint64_t base = lo; base < hi; base += 256) {
        const uint64_t t = base + threadIdx.x;
        bool ok = false;
        uint32_t v = 0, row = 0, col = 0;
        if (t < hi) {
            v = dense[t];
            if (v >= ep.min_kmers && v > 0) {
                tri_decode(t, row, col);
                ok = emit_pass(((uint64_t)row << 32) | col, v, totals, ep);
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) warp_cnt[wid] = __popc(m);
        __syncthreads();
        uint32_t before = 0, all = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) { uint32_t c = warp_cnt[j]; if (j < wid) before += c; all += c; }
        if (ok && write) {
            const uint32_t o = s_run + before + __popc(m & ((1u << lane) - 1));
            out_keys[o] = ((uint64_t)row << ep.gbits) | col;
            out_vals[o] = v;
        }
        __syncthreads();
        if (threadIdx.x == 0) s_run += all;
        __syncthreads();
    }
    if (!write && threadIdx.x == 0) block_cnt[blockIdx.x] = s_run;
}

// exclusive scan of up to 4096 block counts in one block; total -> *total
__global__ void __launch_bounds__(1024) scan_blocks_kernel(uint32_t *__restrict__ cnt, uint32_t n, unsigned long long *__restrict__ total)
{
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t s_total;
    uint32_t v[4], sum = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) { uint32_t i = threadIdx.x * 4 + j; v[j] = i < n ? cnt[i] : 0; sum += v[j]; }
    uint32_t pre = block_exscan_1024(sum, warp_tot, &s_total);
#pragma unroll
    for (int j = 0; j < 4; ++j) { uint32_t i = threadIdx.x * 4 + j; if (i < n) cnt[i] = pre; pre += v[j]; }
    if (threadIdx.x == 0) *total = s_total;
}

__global__ void fill_u64_kernel(uint64_t *p, uint64_t n, uint64_t v)
{
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = v;
}

int grid_for(uint64_t n, int threads = 256, int max_blocks = 148 * 16)
{
    return (int)std::max<uint64_t>(1, std::min<uint64_t>((n + threads - 1) / threads, (uint64_t)max_blocks));
}

}  // namespace

vb_pairs *vb_pairs_alloc(uint64_t n, uint32_t n_genomes);

// shard_count > 1: partial result of one k-mer shard -- no thresholds, ani = 0, total_kmers = this shard's part
void vb_prefilter_impl(vb_ctx *ctx, const vb_genomes *g, const vb_prefilter_params *p, uint32_t shard_index,
                       uint32_t shard_count, vb_pairs **out_pairs)
{
    const bool partial = shard_count > 1;
    if (shard_count == 0 || shard_index >= shard_count) throw vb_error(VB_ERR_ARG, "bad k-mer shard");
    if (p->k < 10 || p->k > 31) throw vb_error(VB_ERR_ARG, "k must be in [10, 31]");
    if (!(p->kmers_fraction > 0)) throw vb_error(VB_ERR_ARG, "kmers_fraction must be > 0");
    cudaStream_t st = (cudaStream_t)ctx->stream;
    VB_CUDA(cudaSetDevice(ctx->device));
    const uint32_t n = g->count();
    EventTimer t_all(st), t_up(st), t_ext(st), t_sort(st), t_seg(st), t_emit(st);

    t_all.start();
    DevBuf<uint32_t> counters(3 * (size_t)std::max<uint32_t>(n, 1));        // valid | dup | totals
    uint32_t *valid_cnt = counters.p, *dup_cnt = counters.p + n, *totals = counters.p + 2 * (size_t)n;
    VB_CUDA(cudaMemsetAsync(counters.p, 0, counters.bytes(), st));
    ExtractParams ep;
    ep.k = p->k;
    ep.shift = 0; ep.tail_mask = 0;
    if (2 * p->k - 32 < 8) { ep.shift = 8 - (2 * p->k - 32); ep.tail_mask = (1ULL << ep.shift) - 1; }
    ep.use_filter = p->kmers_fraction < 1.0;
    ep.max_thr = (uint64_t)((double)UINT64_MAX * (0.0 + p->kmers_fraction));
    ep.c = (uint64_t)std::ceil((double)p->k / 4);
    ep.shard_index = shard_index;
    ep.shard_count = shard_count;
    DevBuf<unsigned long long> scalars(8);
    VB_CUDA(cudaMemsetAsync(scalars.p, 0, scalars.bytes(), st));
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, ctx->device);
    const bool use_lsd = getenv("VB_PREFILTER_LSD") != nullptr;            // A/B switch: the round-1 LSD radix path

    // ---- singleton screen set-up (MSD path), before the genomes are fetched: when they have to be uploaded, the
    // screen pass over each chunk of genomes is enqueued while the next chunk is still on the PCIe bus
    // seen table: >= 4 slots per expected k-mer, at most 2^28 slots (64 MB, stays in the 126 MB L2); for inputs
    // beyond ~2^27 k-mers the collision rate would let most singletons through, so the screen is switched off
    const char *seen_env = getenv("VB_PREFILTER_SEEN");                    // "0": off, "N": force 2^N slots
    const double est_all = (double)vb_store_slots(g, VB_STORE_PAD) * std::min(1.0, p->kmers_fraction) / shard_count;
    int seen_bits = 20;
    while ((double)(1ULL << seen_bits) < 4.0 * est_all && seen_bits < 28) ++seen_bits;
    bool use_seen = !use_lsd && est_all <= (double)(1ULL << 27);
    if (seen_env && !use_lsd) { int v = atoi(seen_env); use_seen = v > 0; if (v >= 10 && v <= 34) seen_bits = v; }
    // The slot axis is walked in chunks (kernel-side offsets inside a chunk stay small; inputs beyond 2^32 base
    // slots just take more launches).  VB_PREFILTER_CHUNK (slots, test hook) forces many small chunks.
    const uint64_t chunk_env = getenv("VB_PREFILTER_CHUNK") ? strtoull(getenv("VB_PREFILTER_CHUNK"), nullptr, 10) : 0;
    const uint64_t chunk = chunk_env ? std::max<uint64_t>(2048, chunk_env / 2048 * 2048) : (1ULL << 31);
    auto grid_of = [&](uint64_t lo, uint64_t hi, int per_sm) {
        return (int)std::max<uint64_t>(1, std::min<uint64_t>(((hi - lo) / KM_ITEMS + 255) / 256, (uint64_t)n_sm * per_sm));
    };
    DevBuf<uint32_t> seen_words;
    SeenTable seen = {nullptr, 0};
    if (use_seen) {
        seen_words.alloc((1ULL << seen_bits) / 16);
        VB_CUDA(cudaMemsetAsync(seen_words.p, 0, seen_words.bytes(), st));
        seen = {seen_words.p, (1ULL << seen_bits) - 1};
    }
    auto screen_range = [&](const DevGenomes &d, uint64_t lo_all, uint64_t hi_all) {
        for (uint64_t lo = lo_all; lo < hi_all; lo += chunk) {
            const uint64_t hi = std::min(hi_all, lo + chunk);
            screen_kernel<<<grid_of(lo, hi, 8), 256, 0, st>>>(d.seq2.p, d.inv_kdb.p, d.tile_gid.p, lo, hi, ep, seen, valid_cnt,
                                                               scalars.p + 4);
            VB_LAUNCH_CHECK(ctx);
        }
    };
    bool screened = false;
    const vb_chunk_fn hook = [&](const DevGenomes &d, uint64_t lo, uint64_t hi) {
        if (use_seen) { screen_range(d, lo, hi); screened = true; }
    };
    t_up.start();
    const DevGenomes &dg = vb_get_dev_genomes(ctx, g, VB_STORE_PAD, nullptr, &hook);
    t_up.stop();
    const uint64_t n_slots = dg.total_slots;

    const unsigned long long max_pairs = (unsigned long long)n * (n > 0 ? n - 1 : 0) / 2;
    unsigned long long n_inc = max_pairs;
    const bool force_hash = getenv("VB_PREFILTER_HASH") != nullptr;   // test hook: exercise the large-N layout
    const bool count_first = force_hash || max_pairs > (1ULL << 26);   // large N: hashed pair table sized by a counting pass
    DevBuf<uint64_t> tkeys;
    DevBuf<uint32_t> tvals, dense;
    DevBuf<int> overflow(1);
    uint64_t cap = 0;
    PairAcc acc = {nullptr, nullptr, nullptr, 0, overflow.p};
    auto alloc_table = [&]() {
        VB_CUDA(cudaMemsetAsync(overflow.p, 0, sizeof(int), st));
        if (!count_first) {                                  // dense triangular counters
            dense.alloc(std::max<unsigned long long>(max_pairs, 1));
            VB_CUDA(cudaMemsetAsync(dense.p, 0, dense.bytes(), st));
            acc.dense = dense.p;
            cap = max_pairs;
            return;
        }
        unsigned long long distinct_bound = std::min(n_inc, max_pairs);
        cap = 1024;
        while (cap < 2 * distinct_bound) cap <<= 1;
        if (cap > (1ULL << 32)) throw vb_error(VB_ERR_MEM, "pair table would exceed 2^32 slots; split the input (--batch-size)");
        tkeys.alloc(cap);
        tvals.alloc(cap);
        VB_CUDA(cudaMemsetAsync(tkeys.p, 0xff, tkeys.bytes(), st));       // SLOT_EMPTY
        VB_CUDA(cudaMemsetAsync(tvals.p, 0, tvals.bytes(), st));
        acc.tkeys = tkeys.p; acc.tvals = tvals.p; acc.cap_mask = cap - 1;
    };
    auto read_n_inc = [&]() {
        VB_CUDA(cudaMemcpyAsync(&n_inc, scalars.p, sizeof(n_inc), cudaMemcpyDeviceToHost, st));
        VB_CUDA(cudaStreamSynchronize(st));
    };
    rsort::Workspace ws;
    unsigned long long n_survivors = 0;

    if (use_lsd) {
        // ---- k1 + k2 + k3, LSD flavour: full stable sort of (k-mer, genome) tuples, then a run scan
        t_ext.start();
        const uint64_t n_pad = ((n_slots + rsort::TILE - 1) / rsort::TILE) * rsort::TILE;
        DevBuf<uint64_t> keys_a(n_pad), keys_b(n_pad);
        DevBuf<uint32_t> vals_a(n_pad), vals_b(n_pad);
        extract_kernel<<<grid_for(n_pad), 256, 0, st>>>(dg.seq2.p, dg.inv_kdb.p, dg.tile_gid.p, n_slots, n_pad, ep, keys_a.p,
                                                       vals_a.p, valid_cnt);
        VB_LAUNCH_CHECK(ctx);
        t_ext.stop();
        t_sort.start();                                     // bit 2k+shift is set only in the sentinel: it sorts last
        const int key_bits = 2 * p->k + ep.shift + 1;
        bool in_b = rsort::sort_kv<8>(ctx, keys_a.p, vals_a.p, keys_b.p, vals_b.p, n_pad, key_bits, ws);
        const uint64_t *skeys = in_b ? keys_b.p : keys_a.p;
        const uint32_t *svals = in_b ? vals_b.p : vals_a.p;
        t_sort.stop();
        t_seg.start();
        if (count_first) {
            segment_count_kernel<<<grid_for(n_pad), 256, 0, st>>>(skeys, svals, n_pad, scalars.p);
            VB_LAUNCH_CHECK(ctx);
            read_n_inc();
        }
        alloc_table();
        segment_pairs_kernel<<<grid_for(n_pad), 256, 0, st>>>(skeys, svals, n_pad, dup_cnt, acc);
        VB_LAUNCH_CHECK(ctx);
    } else {
        // ---- k1 + k2 + k3, MSD flavour: hash once, screen out singletons, hash-bucket partition, shared-memory grouping
        t_ext.start();
        const int fine_bits = est_all > (double)BUCKET_TARGET * (double)(1u << FINE_BITS_SMALL) ? FINE_BITS_LARGE : FINE_BITS_SMALL;
        DevBuf<uint32_t> fine_hist(1u << fine_bits);
        VB_CUDA(cudaMemsetAsync(fine_hist.p, 0, fine_hist.bytes(), st));
        unsigned long long *d_cursor = scalars.p + 6;
        // small inputs: room for every slot's tuple, no counting pass.  Large inputs: count the survivors first.
        const bool exact_alloc = n_slots >= (1ULL << 31) || 12.0 * (double)n_slots > 0.125 * (double)ctx->mem_total ||
                                 getenv("VB_PREFILTER_EXACT") != nullptr;
        auto for_chunks = [&](auto &&launch) {
            for (uint64_t lo = 0; lo < n_slots; lo += chunk) launch(lo, std::min(n_slots, lo + chunk));
        };
        bool counted_valid = false;
        if (use_seen) {
            if (!screened) screen_range(dg, 0, n_slots);         // (already enqueued chunk by chunk when the genomes were uploaded)
            counted_valid = true;
        }
        uint64_t list_cap = n_slots + 64;
        if (exact_alloc) {
            unsigned long long cnt[2] = {0, 0};
            if (use_seen) {
                seen_popc_kernel<<<n_sm * 4, 256, 0, st>>>(seen_words.p, (1ULL << seen_bits) / 16, scalars.p + 5);
                VB_LAUNCH_CHECK(ctx);
                VB_CUDA(cudaMemcpyAsync(cnt, scalars.p + 4, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
                VB_CUDA(cudaStreamSynchronize(st));
                list_cap = cnt[0] + cnt[1] + 64;
            } else {
                for_chunks([&](uint64_t lo, uint64_t hi) {
                    collect_kernel<true><<<grid_of(lo, hi, 8), 256, 0, st>>>(dg.seq2.p, dg.inv_kdb.p, dg.tile_gid.p, lo, hi, ep, seen,
                                                                              valid_cnt, d_cursor, 0, nullptr, nullptr, nullptr, fine_bits);
                    VB_LAUNCH_CHECK(ctx);
                });
                counted_valid = true;
                VB_CUDA(cudaMemcpyAsync(cnt, d_cursor, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
                VB_CUDA(cudaMemsetAsync(d_cursor, 0, sizeof(unsigned long long), st));
                VB_CUDA(cudaStreamSynchronize(st));
                list_cap = cnt[0] + 64;
            }
        }
        if (list_cap >= (1ULL << 32)) throw vb_error(VB_ERR_ARG, "more than 2^32 k-mer tuples in one prefilter call: split the input or shard the k-mers over more GPUs");
        DevBuf<uint64_t> keys0(list_cap);                    // compact list of surviving (hash, genome) tuples
        DevBuf<uint32_t> vals0(list_cap);
        for_chunks([&](uint64_t lo, uint64_t hi) {
            if (counted_valid)
                collect_kernel<false><<<grid_of(lo, hi, 8), 256, 0, st>>>(dg.seq2.p, dg.inv_kdb.p, dg.tile_gid.p, lo, hi, ep, seen, valid_cnt,
                                                                           d_cursor, 1, keys0.p, vals0.p, fine_hist.p, fine_bits);
            else
                collect_kernel<true><<<grid_of(lo, hi, 8), 256, 0, st>>>(dg.seq2.p, dg.inv_kdb.p, dg.tile_gid.p, lo, hi, ep, seen, valid_cnt,
                                                                          d_cursor, 1, keys0.p, vals0.p, fine_hist.p, fine_bits);
            VB_LAUNCH_CHECK(ctx);
        });
        VB_CUDA(cudaMemcpyAsync(&n_survivors, d_cursor, sizeof(n_survivors), cudaMemcpyDeviceToHost, st));
        VB_CUDA(cudaStreamSynchronize(st));                  // the bucket plan and the grids below depend on the count
        const uint64_t n_keep = n_survivors;
        MsdPlan pl;
        {
            int bits = 0;
            while ((double)n_keep / (double)(1ULL << bits) > BUCKET_TARGET && bits < fine_bits) ++bits;
            pl.b1 = (bits + 1) / 2; pl.b2 = bits - pl.b1;
            pl.B1 = 1u << pl.b1; pl.B2 = 1u << pl.b2; pl.NB = pl.B1 * pl.B2;
        }
        DevBuf<uint32_t> hist(pl.NB), off(pl.NB + 1), cursor1(pl.B1), cursor2(pl.NB), tile_start(pl.B1 + 1), big_list(pl.NB + 1);
        DevBuf<uint32_t> n_big(1);
        VB_CUDA(cudaMemsetAsync(hist.p, 0, hist.bytes(), st));
        VB_CUDA(cudaMemsetAsync(n_big.p, 0, sizeof(uint32_t), st));
        coarsen_kernel<<<(1u << fine_bits) / 256, 256, 0, st>>>(fine_hist.p, fine_bits, pl.b1 + pl.b2, hist.p);
        VB_LAUNCH_CHECK(ctx);
        scan_kernel<<<1, 1024, 0, st>>>(hist.p, pl, off.p, cursor1.p, cursor2.p, tile_start.p);
        VB_LAUNCH_CHECK(ctx);
        t_ext.stop();
        t_sort.start();
        DevBuf<uint64_t> keys1(n_keep + 64), keys2(pl.b2 ? n_keep + 64 : 1);
        DevBuf<uint32_t> vals1(n_keep + 64), vals2(pl.b2 ? n_keep + 64 : 1);
        const size_t part_smem = PART_TILE * 12 + 4 * (1 << MAX_BUCKET_BITS) * sizeof(uint32_t);
        static bool attr_done_dev[64] = {false};             // the opt-in is per device (one process may hold several contexts)
        bool &attr_done = attr_done_dev[ctx->device & 63];
        if (!attr_done) {
            VB_CUDA(cudaFuncSetAttribute(part_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)part_smem));
            VB_CUDA(cudaFuncSetAttribute(part_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)part_smem));
            VB_CUDA(cudaFuncSetAttribute(bucket_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BucketSmem)));
            attr_done = true;
        }
        const uint32_t tiles1 = (uint32_t)((n_keep + PART_TILE - 1) / PART_TILE);
        uint64_t *fkeys = keys1.p;
        uint32_t *fvals = vals1.p;
        if (tiles1) {
            part_kernel<1><<<std::min<uint32_t>(tiles1, 148 * 8), PART_THREADS, part_smem, st>>>(
                n_keep, pl, keys0.p, vals0.p, off.p, tile_start.p, tiles1, cursor1.p, keys1.p, vals1.p);
            VB_LAUNCH_CHECK(ctx);
        }
        if (pl.b2) {
            part_kernel<2><<<std::min<uint32_t>(tiles1 + pl.B1, 148 * 8), PART_THREADS, part_smem, st>>>(
                n_keep, pl, keys1.p, vals1.p, off.p, tile_start.p, 0, cursor2.p, keys2.p, vals2.p);
            VB_LAUNCH_CHECK(ctx);
            fkeys = keys2.p; fvals = vals2.p;
        }
        t_sort.stop();
        t_seg.start();
        const int bgrid = (int)std::min<uint32_t>(pl.NB, 148 * 12);
        if (count_first) {
            bucket_kernel<<<bgrid, 256, sizeof(BucketSmem), st>>>(fkeys, fvals, off.p, pl.NB, 1, dup_cnt, acc, big_list.p, n_big.p,
                                                                  scalars.p);
            VB_LAUNCH_CHECK(ctx);
            big_bucket_kernel<<<64, 1024, 0, st>>>(fkeys, fvals, off.p, big_list.p, n_big.p, 1, dup_cnt, acc, scalars.p);
            VB_LAUNCH_CHECK(ctx);
            read_n_inc();
            VB_CUDA(cudaMemsetAsync(n_big.p, 0, sizeof(uint32_t), st));
        }
        alloc_table();
        bucket_kernel<<<bgrid, 256, sizeof(BucketSmem), st>>>(fkeys, fvals, off.p, pl.NB, 0, dup_cnt, acc, big_list.p, n_big.p,
                                                              scalars.p);
        VB_LAUNCH_CHECK(ctx);
        big_bucket_kernel<<<64, 1024, 0, st>>>(fkeys, fvals, off.p, big_list.p, n_big.p, 0, dup_cnt, acc, scalars.p);
        VB_LAUNCH_CHECK(ctx);
    }
    totals_kernel<<<(n + 255) / 256 + 1, 256, 0, st>>>(valid_cnt, dup_cnt, n, totals);
    VB_LAUNCH_CHECK(ctx);
    t_seg.stop();

    // ---- k4
    t_emit.start();
    EmitParams em;
    em.min_kmers = partial ? 1u : (uint32_t)std::max(p->min_kmers, 0);
    em.min_ident_slack = partial ? -1e300 : p->min_ident - 1e-7;
    em.k = p->k;
    em.gbits = 1;
    while ((1ULL << em.gbits) < n) em.gbits++;
    unsigned long long n_emit = 0;
    int h_overflow = 0;
    std::vector<uint64_t> h_keys;
    std::vector<uint32_t> h_vals;
    if (acc.dense) {
        // dense layout: ordered compaction, the output is born sorted by (row, col)
        const uint32_t n_blocks = (uint32_t)std::min<uint64_t>(4096, (max_pairs + 255) / 256 + 1);
        const uint64_t per_block = ((max_pairs + n_blocks - 1) / n_blocks + 255) / 256 * 256;
        DevBuf<uint32_t> block_cnt(4096);
        dense_emit_kernel<<<n_blocks, 256, 0, st>>>(dense.p, max_pairs, per_block, totals, em, 0, block_cnt.p, nullptr, nullptr);
        VB_LAUNCH_CHECK(ctx);
        scan_blocks_kernel<<<1, 1024, 0, st>>>(block_cnt.p, n_blocks, scalars.p + 1);
        VB_LAUNCH_CHECK(ctx);
        VB_CUDA(cudaMemcpyAsync(&n_emit, scalars.p + 1, sizeof(n_emit), cudaMemcpyDeviceToHost, st));
        VB_CUDA(cudaStreamSynchronize(st));
        h_keys.resize(n_emit); h_vals.resize(n_emit);
        if (n_emit) {
            DevBuf<uint64_t> ek(n_emit);
            DevBuf<uint32_t> ev(n_emit);
            dense_emit_kernel<<<n_blocks, 256, 0, st>>>(dense.p, max_pairs, per_block, totals, em, 1, block_cnt.p, ek.p, ev.p);
            VB_LAUNCH_CHECK(ctx);
            VB_CUDA(cudaMemcpyAsync(h_keys.data(), ek.p, sizeof(uint64_t) * n_emit, cudaMemcpyDeviceToHost, st));
            VB_CUDA(cudaMemcpyAsync(h_vals.data(), ev.p, sizeof(uint32_t) * n_emit, cudaMemcpyDeviceToHost, st));
            VB_CUDA(cudaStreamSynchronize(st));
        }
    } else {
        emit_kernel<<<grid_for(cap), 256, 0, st>>>(tkeys.p, tvals.p, cap, totals, em, 0, scalars.p + 1, nullptr, nullptr);
        VB_LAUNCH_CHECK(ctx);
        VB_CUDA(cudaMemcpyAsync(&n_emit, scalars.p + 1, sizeof(n_emit), cudaMemcpyDeviceToHost, st));
        VB_CUDA(cudaMemcpyAsync(&h_overflow, overflow.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        VB_CUDA(cudaStreamSynchronize(st));
        if (h_overflow) throw vb_error(VB_ERR_INTERNAL, "pair table overflow");
        const uint64_t e_pad = ((n_emit + rsort::TILE - 1) / rsort::TILE) * rsort::TILE;
        h_keys.resize(n_emit); h_vals.resize(n_emit);
        if (n_emit) {
            DevBuf<uint64_t> ek_a(e_pad), ek_b(e_pad);
            DevBuf<uint32_t> ev_a(e_pad), ev_b(e_pad);
            fill_u64_kernel<<<grid_for(e_pad), 256, 0, st>>>(ek_a.p, e_pad, KEY_SENTINEL);
            VB_LAUNCH_CHECK(ctx);
            emit_kernel<<<grid_for(cap), 256, 0, st>>>(tkeys.p, tvals.p, cap, totals, em, 1, scalars.p + 2, ek_a.p, ev_a.p);
            VB_LAUNCH_CHECK(ctx);
            bool eb = rsort::sort_kv<8>(ctx, ek_a.p, ev_a.p, ek_b.p, ev_b.p, e_pad, 2 * em.gbits + 1, ws);
            VB_CUDA(cudaMemcpyAsync(h_keys.data(), eb ? ek_b.p : ek_a.p, sizeof(uint64_t) * n_emit, cudaMemcpyDeviceToHost, st));
            VB_CUDA(cudaMemcpyAsync(h_vals.data(), eb ? ev_b.p : ev_a.p, sizeof(uint32_t) * n_emit, cudaMemcpyDeviceToHost, st));
            VB_CUDA(cudaStreamSynchronize(st));
        }
    }
    std::vector<uint32_t> h_tot(n);
    if (n) VB_CUDA(cudaMemcpyAsync(h_tot.data(), totals, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, st));
    t_emit.stop();
    t_all.stop();
    VB_CUDA(cudaStreamSynchronize(st));

    // ---- host: exact IEEE-double metric (params.cpp:28-32) and the two -min filters (sparse_filters.h:49-61)
    const auto hp0 = std::chrono::steady_clock::now();
    std::vector<uint32_t> o_row, o_col, o_common;
    std::vector<double> o_ani;
    o_row.reserve(n_emit); o_col.reserve(n_emit); o_common.reserve(n_emit); o_ani.reserve(n_emit);
    const uint64_t cmask = (1ULL << em.gbits) - 1;
    for (uint64_t i = 0; i < n_emit; ++i) {
        const uint32_t r = (uint32_t)(h_keys[i] >> em.gbits), c = (uint32_t)(h_keys[i] & cmask);
        const double a = partial ? 0.0 : vb_ani_shorter(h_vals[i], h_tot[r], h_tot[c], p->k);
        if (partial || a >= p->min_ident) { o_row.push_back(r); o_col.push_back(c); o_common.push_back(h_vals[i]); o_ani.push_back(a); }
    }
    // --max-seqs: the per-row sampler needs complete counts, so a k-mer shard leaves it to vb_pairs_merge
    if (!partial && p->max_seqs > 0) vb_sample_rows(n, (uint32_t)p->max_seqs, o_row, o_col, o_common, o_ani);
    vb_pairs *res = vb_pairs_alloc(o_row.size(), n);
    for (uint64_t o = 0; o < o_row.size(); ++o) {
        res->row[o] = o_row[o]; res->col[o] = o_col[o]; res->common[o] = o_common[o]; res->ani[o] = o_ani[o];
    }
    for (uint32_t i = 0; i < n; ++i) res->total_kmers[i] = h_tot[i];
    res->k = p->k;
    res->kmers_fraction = p->kmers_fraction;
    *out_pairs = res;
    ctx->set_timing("prefilter.host_post_ms",
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - hp0).count());

    ctx->set_timing("prefilter.total_ms", t_all.ms());
    ctx->set_timing("prefilter.upload_pack_ms", t_up.ms());
    ctx->set_timing("prefilter.extract_ms", t_ext.ms());
    ctx->set_timing("prefilter.sort_ms", t_sort.ms());
    ctx->set_timing("prefilter.segment_ms", t_seg.ms());
    ctx->set_timing("prefilter.emit_ms", t_emit.ms());
    ctx->set_timing("prefilter.tuples", (double)dg.total_slots);
    ctx->set_timing("prefilter.survivors", (double)n_survivors);
    ctx->set_timing("prefilter.pair_increments", (double)n_inc);
    ctx->set_timing("prefilter.table_slots", (double)cap);
    ctx->set_timing("prefilter.candidates", (double)n_emit);
}
