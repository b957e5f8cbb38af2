// The prefilter: all-vs-all shared-k-mer screen (kmer-db build + all2all-sp + the ani-shorter filter) on B200s.
//
// Reference computation (paths under /root/reference/3rd_party/kmer-db/src/):
//   k-mer extraction   kmer_extract.h:13-96, MinHash threshold filter filter.h:33-146
//   per-genome set     console_build.cpp:94-103 (sort + unique; set size = total-kmers, kmer_db.h:129)
//   common counts      prefix_kmer_db.cpp:244-434 + similarity_calculator.cpp:442-657 (bubbles: bubble_helper.h:79-151)
//   tiled variant      console_all2all_parts.cpp:143-331 + db2db_sp similarity_calculator.cpp:1225-1540
//   pair filter        sparse_filters.h:12-61 with metric params.cpp:28-32
//
// Pipeline (integer work, HBM/L2 bound, no tensor cores).  h = fmix64(canonical k-mer) is a bijection, so equal h <=>
// equal k-mer; only GROUPING matters, never numeric order, so nothing is ever fully sorted:
//   extract   packed genomes -> (h, genome) tuples [+ optional singleton screen] + a fine histogram of bucket sizes
//   level 1   tuples -> B1 "parent" buckets by hash range.  Parents are spread over the ranks of a multi-GPU run:
//             this is the send buffer of all-to-all #1 (one rank: it is simply the next stage's input)
//   level 2   parent -> final buckets of ~1 280 tuples
//   group     one block per bucket: shared-memory hash chains -> duplicates per genome, pair increments into a dense
//             triangular or an open-addressing accumulator; oversized buckets and k-mers shared by thousands of
//             genomes ("bubbles") take dedicated paths
//   finish    one rank: thresholds + ordered compaction.  Several ranks: partial counts travel to the owners of both
//             genomes (all-to-all #2), which sort, sum and threshold them.
// Inputs beyond ~10^9 k-mers per pass run in several passes over disjoint slices of the hash space into the SAME
// accumulator (what --batch-size / all2all-parts is for; the result does not depend on it).
// The exact IEEE-double ani-shorter value (libm log() must match glibc's) is computed on the host for the pairs that
// passed; the device decides pass / fail with a margin and flags the (practically non-existent) borderline cases.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <memory>
#include <thread>

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
    uint32_t shard_index;   // this pass keeps the k-mers with h % shard_count == shard_index (passes, vb_prefilter_partial)
    uint32_t shard_count;
    uint32_t gid_base;      // global id of the store's genome 0 (multi-GPU: this rank's block)
};

// How the hash selects a bucket.  hi = top 32 bits of h.  parent = floor(hi * B1 / 2^32) in [0, B1) -- contiguous hash
// ranges, B1 need not be a power of two (B1 = ranks * parents per rank); the low 32 bits of the same product are the
// position inside the parent's range, whose top bits select the level-2 bucket.  The bucket kernel's shared-memory hash
// uses the LOW word of h, and the pass selection uses h % passes: all (practically) independent.
constexpr int FINE_BITS = 10;                              // resolution of the histogram inside a parent (>= level-2 bits)
__device__ __forceinline__ uint32_t parent_of(uint64_t h, uint32_t B1) { return __umulhi((uint32_t)(h >> 32), B1); }
__device__ __forceinline__ uint32_t frac_of(uint64_t h, uint32_t B1) { return (uint32_t)(h >> 32) * B1; }

constexpr int PART_THREADS = 512;
constexpr int PART_ITEMS = 8;
constexpr int PART_TILE = PART_THREADS * PART_ITEMS;      // 4096 tuples per block
constexpr int MAX_BUCKET_BITS = 10;                       // per level
constexpr int BUCKET_CAP = 2048;                          // tuples a bucket may hold to be grouped in shared memory
constexpr int BUCKET_SLOTS = 4096;                        // hash slots per bucket (load <= 0.5)
constexpr int BUCKET_TARGET = 1280;                       // planned mean bucket size (CAP is 20 sigma above it)
constexpr uint32_t HUGE_MIN = 4096;                       // k-mers shared by at least this many tuples: the bubble path

// Singleton screen.  A k-mer that occurs once in the whole input shares nothing and is the common case (92 % of the
// distinct k-mers of c2), so it is dropped before the partition: a table of 2-bit slots indexed by hash bits records
// "seen" (bit 0) and "seen again" (bit 1); only tuples whose slot has bit 1 go on.  Every occurrence of a k-mer maps to
// the same slot, so a k-mer with two or more occurrences (in any genomes, or twice in one genome) always survives;
// a singleton survives only when it collides with another k-mer (harmless).  The table is sized to stay L2-resident;
// used on one rank only (a rank of a multi-GPU run sees only its own genomes at this point).
struct SeenTable {
    uint32_t *words;        // nullptr: screen disabled, everything survives
    uint64_t slot_mask;     // slots - 1 (power of two); 16 slots per word
};

// Append this block's surviving tuples (bit r of `keep` selects h[r]; all of one thread's tuples share a genome) to the
// compact list -- order is irrelevant, tuples are grouped by hash later -- and count them in the fine histogram.  One
// global cursor atomic per call and block.  All threads of the block must call; `phase` alternates 0/1 between successive
// calls so that a call never overwrites prefixes another warp is still reading.  The survivors are compacted in shared
// memory first and leave with fully coalesced stores.  write == 0: only count.
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
                                             uint32_t *__restrict__ fine_hist, uint32_t B1)
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
            atomicAdd(&fine_hist[(parent_of(h[r], B1) << FINE_BITS) | (frac_of(h[r], B1) >> (32 - FINE_BITS))], 1u);
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

__device__ __forceinline__ uint32_t global_gid(const uint32_t *__restrict__ tile_gid, uint64_t p, bool in_range, uint32_t gid_base)
{
    const uint32_t l = in_range ? tile_gid[p >> 7] : 0xffffffffu;
    return l == 0xffffffffu ? l : l + gid_base;
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
        const uint32_t gid = global_gid(tile_gid, p, gi < n_groups, ep.gid_base);
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
                                                      uint32_t *__restrict__ fine_hist, uint32_t B1)
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
        const uint32_t gid = global_gid(tile_gid, p, base + threadIdx.x < n_groups, ep.gid_base);
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
        block_append<KM_ITEMS>(ok, h, gid, s_app, phase, cursor, write, out_keys, out_vals, fine_hist, B1);
    }
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

// ---------------------------------------------------------------------------------------------------------------
// histograms and scans
// ---------------------------------------------------------------------------------------------------------------
// Level-2 bucket of a tuple inside its parent: sub = (fine bin * B2) >> FINE_BITS, B2 <= 2^FINE_BITS buckets per parent (any
// number, so that the mean bucket size can be set exactly).  Coarse histogram: out[parent * B2 + sub] = sum over the
// `n_src` source histograms (stride src_stride) of the fine bins that map to sub.
__device__ __forceinline__ uint32_t sub_of(uint64_t h, uint32_t B1, uint32_t B2)
{
    return ((frac_of(h, B1) >> (32 - FINE_BITS)) * B2) >> FINE_BITS;
}
__global__ void __launch_bounds__(256) coarsen_kernel(const uint32_t *__restrict__ fine, uint32_t n_src, uint64_t src_stride,
                                                      uint32_t B2, uint32_t n_out, uint32_t *__restrict__ out)
{
    // one warp per coarse bin (a bin of the level-1 histogram sums 1 024 fine bins per source)
    const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= n_out) return;
    const uint32_t parent = i / B2, sub = i % B2;
    const uint32_t f_lo = ((sub << FINE_BITS) + B2 - 1) / B2, f_hi = (((sub + 1) << FINE_BITS) + B2 - 1) / B2;
    uint32_t s = 0;
    for (uint32_t src = 0; src < n_src; ++src) {
        const uint32_t *f = fine + src * src_stride + ((uint64_t)parent << FINE_BITS);
        for (uint32_t j = f_lo + lane; j < f_hi; j += 32) s += f[j];
    }
    s = __reduce_add_sync(0xffffffffu, s);
    if (lane == 0) out[i] = s;
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

// Exclusive scan of n <= 2^20 counters in three small launches: 1024-element chunks scanned independently (coalesced),
// the chunk totals scanned by one block, the chunk bases added.  out has n + 1 entries (out[n] = total); out2 (optional)
// receives a copy of the first n (the partition kernels' write cursors).
__global__ void __launch_bounds__(1024) exscan_chunks_kernel(const uint32_t *__restrict__ in, uint32_t n, uint32_t *__restrict__ out,
                                                             uint32_t *__restrict__ chunk_tot)
{
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t s_total;
    const uint32_t i = blockIdx.x * 1024u + threadIdx.x;
    const uint32_t v = i < n ? in[i] : 0u;
    const uint32_t pre = block_exscan_1024(v, warp_tot, &s_total);
    if (i < n) out[i] = pre;
    if (threadIdx.x == 0) chunk_tot[blockIdx.x] = s_total;
}
__global__ void __launch_bounds__(1024) exscan_totals_kernel(uint32_t *__restrict__ chunk_tot, uint32_t n_chunks, uint32_t *__restrict__ grand)
{
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t s_total;
    const uint32_t v = threadIdx.x < n_chunks ? chunk_tot[threadIdx.x] : 0u;
    const uint32_t pre = block_exscan_1024(v, warp_tot, &s_total);
    if (threadIdx.x < n_chunks) chunk_tot[threadIdx.x] = pre;
    if (threadIdx.x == 0) *grand = s_total;
}
__global__ void __launch_bounds__(1024) exscan_add_kernel(uint32_t *__restrict__ out, uint32_t n, const uint32_t *__restrict__ chunk_base,
                                                          uint32_t *__restrict__ out2)
{
    const uint32_t i = blockIdx.x * 1024u + threadIdx.x;
    if (i >= n) return;
    const uint32_t v = out[i] + chunk_base[blockIdx.x];
    out[i] = v;
    if (out2) out2[i] = v;
}

// ---------------------------------------------------------------------------------------------------------------
// partition
// ---------------------------------------------------------------------------------------------------------------
// A source segment of the level-2 partition: a run of tuples that all belong to one parent (one rank: the parent's
// level-1 bucket; several ranks: the part of it that one peer sent).
struct Segment {
    uint32_t beg, len;       // in the level-1 / receive buffer
    uint32_t parent;         // local parent index
    uint32_t tile_start;     // number of PART_TILE tiles in the segments before this one
};

// Partition one tile of tuples into buckets.  LEVEL 1: tuples come from the compact survivor list, bucket = parent.
// LEVEL 2: tuples come from a segment of one parent, bucket = sub_of() = the position inside the parent's range, scaled to B2.
// Inside the block the tile is first grouped by bucket in shared memory, so that every bucket receives one contiguous
// run per tile.  cursor[] holds the next free slot of every destination bucket.
template <int LEVEL>
__global__ void __launch_bounds__(PART_THREADS, 3) part_kernel(uint64_t n_in, uint32_t B1, uint32_t B2, const uint64_t *__restrict__ in_keys,
                                                            const uint32_t *__restrict__ in_vals, const Segment *__restrict__ segs,
                                                            const uint32_t *__restrict__ tile_seg, uint32_t n_tiles,
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
    __shared__ uint32_t s_total, s_wtot[PART_THREADS / 32];
    const uint32_t NBK = (LEVEL == 1) ? B1 : B2;
    auto bucket_of = [&](uint64_t h) -> uint32_t {
        if (LEVEL == 1) return parent_of(h, B1);
        return sub_of(h, B1, B2);
    };

    // The loads of a tile are issued one tile AHEAD, into the registers the previous tile has just vacated (its tuples sit
    // in shared memory by then), so that they are in flight during the previous tile's write-out instead of stalling the
    // histogram (42 % of the stall samples were these loads, plus the barrier behind them).
    uint64_t key[PART_ITEMS];
    uint32_t val[PART_ITEMS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // source range of a tile -> number of tuples in it, cursor base of its parent; issues the loads
    auto fetch = [&](uint32_t tile, uint32_t &cnt, uint32_t &cur_base) {
        uint64_t src_lo;
        if (LEVEL == 2) {
            const Segment sg = segs[tile_seg[tile]];                    // (host-built map: a search here was 13 % of the kernel's instructions)
            const uint32_t done = (tile - sg.tile_start) * (uint32_t)PART_TILE;
            src_lo = (uint64_t)sg.beg + done;
            cnt = min((uint32_t)PART_TILE, sg.len - done);
            cur_base = sg.parent * B2;
        } else {
            src_lo = (uint64_t)tile * PART_TILE;
            cnt = (uint32_t)min((uint64_t)PART_TILE, n_in - src_lo);
            cur_base = 0;
        }
        const uint64_t *kp = in_keys + src_lo + threadIdx.x;
        const uint32_t *vp = in_vals + src_lo + threadIdx.x;
#pragma unroll
        for (int r = 0; r < PART_ITEMS; ++r)
            if ((uint32_t)(r * PART_THREADS) + threadIdx.x < cnt) { key[r] = kp[r * PART_THREADS]; val[r] = vp[r * PART_THREADS]; }
    };

    uint32_t tile = blockIdx.x;
    if (tile >= n_tiles) return;                                        // (uniform)
    uint32_t cnt, cur_base;
    fetch(tile, cnt, cur_base);
    for (;;) {
        for (uint32_t b = threadIdx.x; b < NBK; b += PART_THREADS) { s_cnt[b] = 0; s_fill[b] = 0; }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < PART_ITEMS; ++r)
            if ((uint32_t)(r * PART_THREADS) + threadIdx.x < cnt) atomicAdd(&s_cnt[bucket_of(key[r])], 1u);
        __syncthreads();
        // exclusive scan of s_cnt (NBK <= 1024 = 2 per thread; every warp takes part) + one global reservation per bucket
        {
            const uint32_t b0 = 2 * threadIdx.x;
            const uint32_t v0 = b0 < NBK ? s_cnt[b0] : 0u, v1 = b0 + 1 < NBK ? s_cnt[b0 + 1] : 0u;
            uint32_t x = v0 + v1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
            if (lane == 31) s_wtot[wid] = x;
            __syncthreads();
            uint32_t base = 0;
#pragma unroll
            for (int k = 0; k < PART_THREADS / 32; ++k) { const uint32_t t = s_wtot[k]; if (k < wid) base += t; }
            const uint32_t pre = base + x - (v0 + v1);
            if (b0 < NBK) s_start[b0] = pre;
            if (b0 + 1 < NBK) s_start[b0 + 1] = pre + v0;
            if (threadIdx.x == PART_THREADS - 1) s_total = base + x;
        }
        __syncthreads();
        // one global reservation per non-empty bucket; the answers are only needed for the write-out, so they stay in
        // flight (in registers) while the tile is staged and the next one is requested
        static_assert((1 << MAX_BUCKET_BITS) <= 2 * PART_THREADS, "two buckets per thread");
        uint32_t gb0 = 0, gb1 = 0;
        {
            const uint32_t b0 = threadIdx.x, b1 = threadIdx.x + PART_THREADS;
            const uint32_t c0 = b0 < NBK ? s_cnt[b0] : 0u, c1 = b1 < NBK ? s_cnt[b1] : 0u;
            if (c0) gb0 = atomicAdd(&cursor[cur_base + b0], c0);
            if (c1) gb1 = atomicAdd(&cursor[cur_base + b1], c1);
        }
#pragma unroll
        for (int r = 0; r < PART_ITEMS; ++r) {
            if ((uint32_t)(r * PART_THREADS) + threadIdx.x < cnt) {
                const uint32_t b = bucket_of(key[r]);                   // (recomputed: cheaper than 8 registers held across the scan)
                const uint32_t pos = s_start[b] + atomicAdd(&s_fill[b], 1u);
                st_keys[pos] = key[r];
                st_vals[pos] = val[r];
            }
        }
        // the next tile's tuples start travelling now
        const uint32_t next = tile + gridDim.x;
        const bool more = next < n_tiles;                               // (uniform)
        uint32_t cnt_next = 0, base_next = 0;
        if (more) fetch(next, cnt_next, base_next);
        s_gbase[threadIdx.x] = gb0;
        s_gbase[threadIdx.x + PART_THREADS] = gb1;
        __syncthreads();
        const uint32_t total = s_total;
        for (uint32_t s = threadIdx.x; s < total; s += PART_THREADS) {
            uint64_t k = st_keys[s];
            uint32_t b = bucket_of(k);
            uint32_t dst = s_gbase[b] + (s - s_start[b]);
            out_keys[dst] = k;
            out_vals[dst] = st_vals[s];
        }
        __syncthreads();
        if (!more) break;
        tile = next; cnt = cnt_next; cur_base = base_next;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// pair accumulator
// ---------------------------------------------------------------------------------------------------------------
// Two layouts:
//   dense  (N(N-1)/2 <= 2^26): one uint32 counter per pair at index row*(row-1)/2 + col -- an increment is a single
//          fire-and-forget atomic, and an ordered scan of the array yields the pairs already sorted by (row, col);
//   hashed (larger N): open addressing, uint64 key (row << 32 | col) + uint32 count.  Sized by a guess, grown between
//          passes when it fills up; an insert that finds no slot sets *overflow and the whole call is redone larger.
struct PairSlot { uint64_t key; uint32_t val; uint32_t pad; };      // key and counter share a 32-byte sector: one DRAM access per increment
struct PairAcc {
    uint32_t *dense;        // non-null selects the dense layout
    PairSlot *slots;
    uint64_t cap_mask;
    unsigned long long *status;     // [0] overflow flag, [1] keys inserted so far
};
constexpr uint32_t MAX_PROBES = 4096;

// returns 1 when the key was new
__device__ __forceinline__ uint32_t table_add(const PairAcc &A, uint64_t key, uint32_t inc)
{
    uint64_t h = fmix64(key) & A.cap_mask;
    for (uint32_t probes = 0; probes < MAX_PROBES; ++probes) {
        uint64_t cur = A.slots[h].key;
        uint32_t fresh = 0;
        if (cur == SLOT_EMPTY) {
            cur = atomicCAS((unsigned long long *)&A.slots[h].key, (unsigned long long)SLOT_EMPTY, (unsigned long long)key);
            if (cur == SLOT_EMPTY) { cur = key; fresh = 1; }
        }
        if (cur == key) { atomicAdd(&A.slots[h].val, inc); return fresh; }
        h = (h + 1) & A.cap_mask;
    }
    A.status[0] = 1;
    return 0;
}

// `inc` increments for the pair (a, b), a != b
__device__ __forceinline__ uint32_t pair_add(const PairAcc &A, uint32_t a, uint32_t b, uint32_t inc = 1u)
{
    const uint32_t hi = a > b ? a : b, lo = a > b ? b : a;
    if (A.dense) { atomicAdd(&A.dense[(uint64_t)hi * (hi - 1) / 2 + lo], inc); return 0; }
    return table_add(A, ((uint64_t)hi << 32) | lo, inc);
}

__device__ __forceinline__ void flush_fresh(const PairAcc &A, uint32_t fresh)
{
    fresh = __reduce_add_sync(0xffffffffu, fresh);
    if ((threadIdx.x & 31) == 0 && fresh) atomicAdd(&A.status[1], (unsigned long long)fresh);
}

// move every entry of an old table into a new (larger) one
__global__ void __launch_bounds__(256) rehash_kernel(const PairSlot *__restrict__ old, uint64_t ocap, PairAcc A)
{
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < ocap; i += (uint64_t)gridDim.x * blockDim.x) {
        const PairSlot e = old[i];
        if (e.key != SLOT_EMPTY) table_add(A, e.key, e.val);
    }
}

// all slots empty: key = ~0, count 0
__global__ void __launch_bounds__(256) table_clear_kernel(PairSlot *__restrict__ slots, uint64_t cap)
{
    uint4 *p = (uint4 *)slots;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < cap; i += (uint64_t)gridDim.x * blockDim.x)
        p[i] = make_uint4(0xffffffffu, 0xffffffffu, 0u, 0u);
}

// shared-memory layout of bucket_kernel (dynamic, 52 KB -> 4 blocks per SM).  The hashes are only needed while the tuples
// are grouped; their space then holds the group-sorted member list and the scans.
constexpr uint32_t MEMBER_DUP = 0x8000u;       // member[]: this tuple repeats a genome that is already in its group
struct BucketSmem {
    union {
        uint64_t keys[BUCKET_CAP + 1];         // h of every tuple (grouping)
        struct {
            uint32_t poff[BUCKET_CAP + 1];     // pairs of all groups before tuple i's group (non-zero width only at representatives)
            uint16_t off[BUCKET_CAP];          // first member of the group whose representative is tuple i
            uint16_t member[BUCKET_CAP];       // tuple indices, group by group
        } g;
    } u;
    uint32_t gids[BUCKET_CAP];
    uint32_t table[BUCKET_SLOTS];              // slot -> representative (first inserted tuple) of the slot's k-mer
    uint32_t cnt[BUCKET_CAP];                  // members of the group whose representative is tuple i
    uint16_t rep[BUCKET_CAP];                  // representative of tuple i's group
    uint32_t warp_tot[2][8];
    uint32_t total_pairs;
};

// One block per final bucket.  Equal k-mers are grouped through a shared-memory hash table (the first tuple of a k-mer
// becomes the group's representative), a counting sort lays every group out contiguously, a tuple whose genome already
// occurs earlier in its group is a within-genome duplicate (counted in dup_cnt, skipped afterwards), and the pairs of ALL
// groups are then enumerated as one flat index space split evenly over the threads -- every lane does the same amount
// of work whatever the group sizes (walking per-tuple chains instead left most lanes of a warp idle: 5.3 G warp
// instructions at c3, 40 % of them in the pair loop).
__global__ void __launch_bounds__(256) bucket_flat_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals,
                                                          const uint32_t *__restrict__ off, uint32_t n_buckets,
                                                          uint32_t *__restrict__ dup_cnt, PairAcc A,
                                                          uint32_t *__restrict__ big_list, uint32_t *__restrict__ n_big)
{
    extern __shared__ unsigned char smem_raw[];
    BucketSmem &S = *(BucketSmem *)smem_raw;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    uint32_t fresh = 0;
    for (uint32_t bkt = blockIdx.x; bkt < n_buckets; bkt += gridDim.x) {
        const uint32_t beg = off[bkt], size = off[bkt + 1] - beg;
        if (size < 2) continue;                                              // uniform for the block
        if (size > BUCKET_CAP) {                                             // too big for shared memory: generic path
            if (tid == 0) big_list[atomicAdd(n_big, 1u)] = bkt;
            continue;
        }
        for (int s = tid; s < BUCKET_SLOTS; s += 256) S.table[s] = 0xffffffffu;
        for (uint32_t i = tid; i < size; i += 256) { S.u.keys[i] = keys[beg + i]; S.gids[i] = vals[beg + i]; S.cnt[i] = 0; }
        {   // the bucket this block takes next: pull its lines into L2 while this one is grouped
            const uint32_t nxt = bkt + gridDim.x;
            if (nxt < n_buckets) {
                const uint32_t nb = off[nxt], ns = min(off[nxt + 1] - nb, (uint32_t)BUCKET_CAP);
                for (uint32_t i = tid * 16; i < ns; i += 256 * 16) {
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(keys + nb + i));
                    if ((i & 31) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(vals + nb + i));
                }
            }
        }
        __syncthreads();
        // ---- 1. group: every tuple learns its representative, representatives count their members
        for (uint32_t i = tid; i < size; i += 256) {
            const uint64_t k = S.u.keys[i];
            uint32_t s = (uint32_t)k & (BUCKET_SLOTS - 1);                    // low bits: independent of the bucket bits
            uint32_t r;
            for (;;) {
                uint32_t cur = S.table[s];
                if (cur == 0xffffffffu) {
                    cur = atomicCAS(&S.table[s], 0xffffffffu, i);
                    if (cur == 0xffffffffu) { r = i; break; }                 // first tuple with this key
                }
                if (S.u.keys[cur] == k) { r = cur; break; }
                s = (s + 1) & (BUCKET_SLOTS - 1);
            }
            S.rep[i] = (uint16_t)r;
            atomicAdd(&S.cnt[r], 1u);
        }
        __syncthreads();
        // ---- 2. exclusive scans over the tuples (8 per thread): members -> off, pairs m (m - 1) / 2 -> poff
        {
            uint32_t c[8], sum_c = 0, sum_p = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t i = tid * 8 + j;
                c[j] = i < size ? S.cnt[i] : 0u;
                sum_c += c[j]; sum_p += c[j] * (c[j] - 1) / 2;               // (c == 0: 0 * 0xffffffff / 2 = 0)
            }
            uint32_t xc = sum_c, xp = sum_p;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t yc = __shfl_up_sync(0xffffffffu, xc, o), yp = __shfl_up_sync(0xffffffffu, xp, o);
                if (lane >= o) { xc += yc; xp += yp; }
            }
            if (lane == 31) { S.warp_tot[0][wid] = xc; S.warp_tot[1][wid] = xp; }
            __syncthreads();                                                  // (also: nobody reads the hashes any more)
            uint32_t bc = 0, bp = 0, tp = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) { const uint32_t t0 = S.warp_tot[0][k], t1 = S.warp_tot[1][k]; if (k < wid) { bc += t0; bp += t1; } tp += t1; }
            uint32_t rc = bc + xc - sum_c, rp = bp + xp - sum_p;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t i = tid * 8 + j;
                if (i < size) { S.u.g.off[i] = (uint16_t)rc; S.u.g.poff[i] = rp; }
                rc += c[j]; rp += c[j] * (c[j] - 1) / 2;
            }
            if (tid == 0) { S.u.g.poff[size] = tp; S.total_pairs = tp; }
        }
        __syncthreads();
        // ---- 3. counting sort: the members of every group, contiguous
        for (uint32_t i = tid; i < size; i += 256) {
            const uint32_t r = S.rep[i];
            S.u.g.member[S.u.g.off[r] + atomicSub(&S.cnt[r], 1u) - 1] = (uint16_t)i;
        }
        __syncthreads();
        const uint32_t P = S.total_pairs;
        if (P == 0) continue;                                                 // all singletons (uniform: every thread read the same word)
        // ---- 4. duplicates: the same genome earlier in the group's member list
        for (uint32_t x = tid; x < size; x += 256) {
            const uint32_t i = S.u.g.member[x] & (MEMBER_DUP - 1);
            const uint32_t start = S.u.g.off[S.rep[i]];
            if (start == x) continue;
            const uint32_t g = S.gids[i];
            for (uint32_t y = start; y < x; ++y)
                if (S.gids[S.u.g.member[y] & (MEMBER_DUP - 1)] == g) {
                    S.u.g.member[x] = (uint16_t)(i | MEMBER_DUP);
                    atomicAdd(&dup_cnt[g], 1u);
                    break;
                }
        }
        __syncthreads();
        // ---- 5. pair increments: the pairs (a, b), b < a, of all groups as one index space, an equal share per thread
        {
            const uint32_t chunk = (P + 255) / 256;
            uint32_t p0 = tid * chunk;
            const uint32_t p1 = min(P, p0 + chunk);
            if (p0 < p1) {
                uint32_t lo = 0, hi = size;                                   // the last tuple index with poff <= p0: a representative
                while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (S.u.g.poff[mid] <= p0) lo = mid + 1; else hi = mid; }
                uint32_t r = lo - 1;
                uint32_t q = p0 - S.u.g.poff[r];
                uint32_t pc = S.u.g.poff[r + 1] - S.u.g.poff[r];
                uint32_t start = S.u.g.off[r];
                uint32_t a = (uint32_t)((1.0f + sqrtf(1.0f + 8.0f * (float)q)) * 0.5f);
                while (a * (a - 1) / 2 > q) --a;
                while ((a + 1) * a / 2 <= q) ++a;
                uint32_t b = q - a * (a - 1) / 2;
                for (; p0 < p1; ++p0) {
                    const uint32_t xa = S.u.g.member[start + a], xb = S.u.g.member[start + b];
                    if (!((xa | xb) & MEMBER_DUP)) fresh += pair_add(A, S.gids[xa], S.gids[xb]);
                    if (++q == pc) {                                          // next group with pairs
                        if (p0 + 1 == p1) break;
                        do { ++r; pc = S.u.g.poff[r + 1] - S.u.g.poff[r]; } while (pc == 0);
                        start = S.u.g.off[r]; q = 0; a = 1; b = 0;
                    } else if (++b == a) { ++a; b = 0; }
                }
            }
        }
        __syncthreads();
    }
    if (!A.dense) flush_fresh(A, fresh);
}

// The same grouping with per-tuple CHAINS instead of group arrays: a tuple finds the slot of its key and swaps itself in
// as the slot's newest member, keeping the previous one as its predecessor; its chain is then exactly the set of tuples
// with the same k-mer inserted before it, so walking it enumerates every unordered pair of the group once.  Fewer
// phases (three block-wide barriers fewer) but the walks leave lanes idle when group sizes differ inside a warp: faster
// for small families (c2, c3: 8.4 vs 13.2 ms), no faster for families of hundreds; vb_prefilter_run picks the flat kernel
// for large groups on the hashed table (VB_PREFILTER_BUCKET=chain|flat forces one).
// (Tried and dropped in round 2: compacting the tuples that have a chain into a work list and walking them with lane refill
// -- a lane whose chain ends takes the next tuple.  Same 8.4 ms at c3 with warp-private list slices, 9.4 ms with a shared
// counter, 40.6 vs 36.9 ms at c3_s200: the kernel is not bound by idle lanes but by many small costs at c3 -- 5 G warp
// instructions spread over load, insert, duplicate check and pair loop, 55 % of the issue slots used -- and by the L2
// atomic rate, 1.1e11 increments/s, once families are large.  profiles/r02_ab_kernels.txt.)
constexpr uint32_t CHAIN_END = 0xffffu;
struct ChainSmem {
    uint64_t keys[BUCKET_CAP];                 // h of every tuple
    uint32_t gids[BUCKET_CAP];
    uint32_t table[BUCKET_SLOTS];              // slot -> the most recently inserted tuple with the slot's key
    uint16_t prev[BUCKET_CAP];                 // the previous tuple with the same k-mer
    uint32_t dup[BUCKET_CAP / 32];             // bit i: tuple i repeats a genome that is already in its chain
};

__global__ void __launch_bounds__(256) bucket_chain_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals,
                                                           const uint32_t *__restrict__ off, uint32_t n_buckets,
                                                           uint32_t *__restrict__ dup_cnt, PairAcc A,
                                                           uint32_t *__restrict__ big_list, uint32_t *__restrict__ n_big)
{
    extern __shared__ unsigned char smem_raw[];
    ChainSmem &S = *(ChainSmem *)smem_raw;
    const int tid = threadIdx.x;
    uint32_t fresh = 0;
    for (uint32_t bkt = blockIdx.x; bkt < n_buckets; bkt += gridDim.x) {
        const uint32_t beg = off[bkt], size = off[bkt + 1] - beg;
        if (size < 2) continue;                                              // uniform for the block
        if (size > BUCKET_CAP) {                                             // too big for shared memory: generic path
            if (tid == 0) big_list[atomicAdd(n_big, 1u)] = bkt;
            continue;
        }
        for (int s = tid; s < BUCKET_SLOTS; s += 256) S.table[s] = 0xffffffffu;
        if (tid < BUCKET_CAP / 32) S.dup[tid] = 0;
        for (uint32_t i = tid; i < size; i += 256) { S.keys[i] = keys[beg + i]; S.gids[i] = vals[beg + i]; }
        {   // the bucket this block takes next: pull its lines into L2 while this one is grouped
            const uint32_t nxt = bkt + gridDim.x;
            if (nxt < n_buckets) {
                const uint32_t nb = off[nxt], ns = min(off[nxt + 1] - nb, (uint32_t)BUCKET_CAP);
                for (uint32_t i = tid * 16; i < ns; i += 256 * 16) {
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(keys + nb + i));
                    if ((i & 31) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(vals + nb + i));
                }
            }
        }
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
        // ---- duplicates: same genome earlier in the chain (flagged in a bit mask, read by the next phase)
        for (uint32_t i = tid; i < size; i += 256) {
            const uint32_t g = S.gids[i];
            uint32_t j = S.prev[i];
            while (j != CHAIN_END) {
                if (S.gids[j] == g) {
                    atomicOr(&S.dup[i >> 5], 1u << (i & 31));
                    atomicAdd(&dup_cnt[g], 1u);
                    break;
                }
                j = S.prev[j];
            }
        }
        __syncthreads();
        // ---- pair increments: every non-duplicate tuple with every non-duplicate tuple before it in its chain
        for (uint32_t i = tid; i < size; i += 256) {
            uint32_t j = S.prev[i];
            if (j == CHAIN_END || ((S.dup[i >> 5] >> (i & 31)) & 1u)) continue;
            const uint32_t g = S.gids[i];
            while (j != CHAIN_END) {
                if (!((S.dup[j >> 5] >> (j & 31)) & 1u)) fresh += pair_add(A, g, S.gids[j]);
                j = S.prev[j];
            }
        }
        __syncthreads();
    }
    if (!A.dense) flush_fresh(A, fresh);
}

// Generic path for buckets that do not fit shared memory: the block sorts the bucket in place in global memory by
// (h, genome) with the all-ascending bitonic network, then scans the runs of equal k-mers.  A run of at least HUGE_MIN
// tuples (a k-mer shared by thousands of genomes -- what kmer-db calls a bubble, bubble_helper.h:79-151) is not expanded
// here: its position goes to huge_list and the bubble kernels below take it.
struct HugeRun { uint32_t beg, len; };

__global__ void __launch_bounds__(1024) big_bucket_kernel(uint64_t *__restrict__ keys, uint32_t *__restrict__ vals,
                                                          const uint32_t *__restrict__ off, const uint32_t *__restrict__ big_list,
                                                          const uint32_t *__restrict__ n_big, uint32_t *__restrict__ dup_cnt, PairAcc A,
                                                          HugeRun *__restrict__ huge_list, uint32_t huge_cap, unsigned long long *__restrict__ n_huge)
{
    const uint32_t nb = *n_big;
    uint32_t fresh = 0;
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
            const uint64_t key = K[i];
            const uint32_t g = V[i];
            // the run of equal keys around i, by binary search in the sorted bucket
            uint32_t first = 0, end = c;
            { uint32_t x = 0, y = i; while (x < y) { const uint32_t mid = x + (y - x) / 2; if (K[mid] < key) x = mid + 1; else y = mid; } first = x; }
            { uint32_t x = i, y = c; while (y - x > 1) { const uint32_t mid = x + (y - x) / 2; if (K[mid] == key) x = mid; else y = mid; } end = y; }
            const bool huge = end - first >= HUGE_MIN;
            if (huge && i == first) {
                const unsigned long long at = atomicAdd(n_huge, 1ULL);
                if (at < huge_cap) huge_list[at] = {beg + first, end - first};
            }
            if (huge) continue;
            if (i > 0 && K[i - 1] == key && V[i - 1] == g) { atomicAdd(&dup_cnt[g], 1u); continue; }
            uint32_t prev = g;
            for (uint32_t j = i; j-- > 0 && K[j] == key;) {
                uint32_t gj = V[j];
                if (gj != prev) { fresh += pair_add(A, g, gj); prev = gj; }
            }
        }
        __syncthreads();
    }
    if (!A.dense) flush_fresh(A, fresh);
}

// ---- bubbles: one k-mer shared by thousands of genomes ------------------------------------------------------------
// bubble_prepare: the run is sorted by genome, so the distinct genomes are the positions whose predecessor differs;
// they are compacted into members[] (at the run's own offset, so no allocation depends on the data), duplicates are
// counted, and a 128-bit signature of the member set is formed.  Identical member sets -- "core" k-mers of one clade --
// are then expanded ONCE with a weight (kmer-db's pattern collapse, prefix_kmer_db.cpp:181-240, for the case where it
// matters): the host groups the signatures, bubble_equal verifies the candidates element by element.
struct HugeInfo { uint32_t beg, m; uint64_t sig1, sig2; };

__global__ void __launch_bounds__(1024) bubble_prepare_kernel(const uint32_t *__restrict__ vals, const HugeRun *__restrict__ runs,
                                                              uint32_t n_runs, uint32_t *__restrict__ members, HugeInfo *__restrict__ info,
                                                              uint32_t *__restrict__ dup_cnt)
{
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t s_total, s_run;
    __shared__ unsigned long long s_sig[2];
    for (uint32_t r = blockIdx.x; r < n_runs; r += gridDim.x) {
        const HugeRun run = runs[r];
        const uint32_t *V = vals + run.beg;
        if (threadIdx.x == 0) { s_run = 0; s_sig[0] = 0; s_sig[1] = 0; }
        __syncthreads();
        unsigned long long sig1 = 0, sig2 = 0;
        for (uint32_t base = 0; base < run.len; base += 1024) {
            const uint32_t i = base + threadIdx.x;
            uint32_t g = 0, keep = 0;
            if (i < run.len) {
                g = V[i];
                keep = (i == 0 || V[i - 1] != g) ? 1u : 0u;
                if (!keep) atomicAdd(&dup_cnt[g], 1u);
            }
            const uint32_t pre = block_exscan_1024(keep, warp_tot, &s_total);
            if (keep) {
                members[run.beg + s_run + pre] = g;
                sig1 += fmix64(0x9E3779B97F4A7C15ULL + g);
                sig2 += fmix64(0xC2B2AE3D27D4EB4FULL ^ ((uint64_t)g << 17 | g));
            }
            __syncthreads();
            if (threadIdx.x == 0) s_run += s_total;
            __syncthreads();
        }
        atomicAdd(&s_sig[0], sig1); atomicAdd(&s_sig[1], sig2);
        __syncthreads();
        if (threadIdx.x == 0) info[r] = {run.beg, s_run, (uint64_t)s_sig[0], (uint64_t)s_sig[1]};
        __syncthreads();
    }
}

// same[i] = 1 when member list i equals member list rep[i] (both of length m)
__global__ void __launch_bounds__(256) bubble_equal_kernel(const uint32_t *__restrict__ members, const HugeInfo *__restrict__ info,
                                                           const uint32_t *__restrict__ rep, uint32_t n, uint32_t *__restrict__ differ)
{
    for (uint32_t r = blockIdx.y; r < n; r += gridDim.y) {
        const uint32_t q = rep[r];
        if (q == r) continue;
        const uint32_t *a = members + info[r].beg, *b = members + info[q].beg;
        const uint32_t m = info[r].m;
        bool bad = false;
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) bad |= a[i] != b[i];
        if (bad) differ[r] = 1;
    }
}

// all pairs of one member set, 128 x 128 tiles of the lower triangle, `weight` increments each
struct BubbleJob { uint32_t beg, m, weight, tile0; };      // tile0: number of tiles of the jobs before this one
__global__ void __launch_bounds__(256) bubble_pairs_kernel(const uint32_t *__restrict__ members, const BubbleJob *__restrict__ jobs,
                                                           uint32_t n_jobs, uint32_t n_tiles, PairAcc A)
{
    __shared__ uint32_t s_row[128], s_col[128];
    __shared__ uint32_t s_job;
    uint32_t fresh = 0;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        if (threadIdx.x == 0) {
            uint32_t lo = 0, hi = n_jobs;
            while (hi - lo > 1) { uint32_t mid = (lo + hi) / 2; if (jobs[mid].tile0 <= tile) lo = mid; else hi = mid; }
            s_job = lo;
        }
        __syncthreads();
        const BubbleJob jb = jobs[s_job];
        // tile t of the job -> (tr, tc), tc <= tr, row-major over the lower triangle of T x T tiles
        const uint32_t t = tile - jb.tile0;
        uint32_t tr = (uint32_t)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
        while ((uint64_t)tr * (tr + 1) / 2 > t) --tr;
        while ((uint64_t)(tr + 1) * (tr + 2) / 2 <= t) ++tr;
        const uint32_t tc = t - (uint32_t)((uint64_t)tr * (tr + 1) / 2);
        const uint32_t *M = members + jb.beg;
        if (threadIdx.x < 128) { const uint32_t i = tr * 128 + threadIdx.x; s_row[threadIdx.x] = i < jb.m ? M[i] : 0xffffffffu; }
        else { const uint32_t i = tc * 128 + threadIdx.x - 128; s_col[threadIdx.x - 128] = i < jb.m ? M[i] : 0xffffffffu; }
        __syncthreads();
        for (uint32_t e = threadIdx.x; e < 128 * 128; e += 256) {
            const uint32_t r = e >> 7, c = e & 127;
            if (tr == tc && c >= r) continue;                     // diagonal tile: strictly lower part only
            const uint32_t a = s_row[r], b = s_col[c];
            if (a == 0xffffffffu || b == 0xffffffffu) continue;
            fresh += pair_add(A, a, b, jb.weight);
        }
        __syncthreads();
    }
    if (!A.dense) flush_fresh(A, fresh);
}

// ---- deferred bubbles (hashed accumulator) ---------------------------------------------------------------------------
// Expanding a k-mer shared by 20 000 genomes costs 2 x 10^8 table entries -- nearly all of them pairs that share nothing
// else and can never reach --min-kmers.  kmer-db applies its bubbles per row when the matrix is compacted
// (array.h:392-446); the analogue here: the member sets are kept (BubSet), every rank learns all of them, and
//   * a pair of two "heavy" genomes (total bubble weight W >= min_kmers each: such a pair could pass on bubbles alone) gets
//     its bubble counts by expansion, restricted to the heavy members of each set;
//   * every other pair can only pass if it also shares an ordinary k-mer, i.e. if it is in some rank's table anyway: the
//     finalize kernel adds  sum of w over the sets that contain both genomes  (intersection of two short presence lists)
//     to the summed partial counts before it applies the thresholds.
struct BubSet { uint32_t off, m, w; };          // members [off, off + m) of the bubble store, multiplicity w

__global__ void __launch_bounds__(256) bub_count_kernel(const uint32_t *__restrict__ members, const BubSet *__restrict__ sets, uint32_t n_sets,
                                                        uint32_t *__restrict__ cnt, uint32_t *__restrict__ W)
{
    for (uint32_t b = blockIdx.x; b < n_sets; b += gridDim.x) {
        const BubSet st = sets[b];
        for (uint32_t i = threadIdx.x; i < st.m; i += blockDim.x) {
            const uint32_t g = members[st.off + i];
            atomicAdd(&cnt[g], 1u);
            atomicAdd(&W[g], st.w);
        }
    }
}
__global__ void __launch_bounds__(256) bub_fill_kernel(const uint32_t *__restrict__ members, const BubSet *__restrict__ sets, uint32_t n_sets,
                                                       const uint32_t *__restrict__ off, uint32_t *__restrict__ fill, uint32_t *__restrict__ pres)
{
    for (uint32_t b = blockIdx.x; b < n_sets; b += gridDim.x) {
        const BubSet st = sets[b];
        for (uint32_t i = threadIdx.x; i < st.m; i += blockDim.x) {
            const uint32_t g = members[st.off + i];
            pres[off[g] + atomicAdd(&fill[g], 1u)] = b;
        }
    }
}
// the heavy members (W >= thr) of sets [first, first + n_own), compacted in order into out at the set's own offset
__global__ void __launch_bounds__(1024) bub_filter_kernel(const uint32_t *__restrict__ members, const BubSet *__restrict__ sets, uint32_t first,
                                                          uint32_t n_own, const uint32_t *__restrict__ W, uint32_t thr,
                                                          uint32_t *__restrict__ out, uint32_t *__restrict__ m_out)
{
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t s_total, s_run;
    for (uint32_t q = blockIdx.x; q < n_own; q += gridDim.x) {
        const BubSet st = sets[first + q];
        if (threadIdx.x == 0) s_run = 0;
        __syncthreads();
        for (uint32_t base = 0; base < st.m; base += 1024) {
            const uint32_t i = base + threadIdx.x;
            uint32_t g = 0, keep = 0;
            if (i < st.m) { g = members[st.off + i]; keep = W[g] >= thr ? 1u : 0u; }
            const uint32_t pre = block_exscan_1024(keep, warp_tot, &s_total);
            if (keep) out[st.off + s_run + pre] = g;
            __syncthreads();
            if (threadIdx.x == 0) s_run += s_total;
            __syncthreads();
        }
        if (threadIdx.x == 0) m_out[q] = s_run;
        __syncthreads();
    }
}

__global__ void totals_kernel(const uint32_t *__restrict__ valid_cnt, const uint32_t *__restrict__ dup_cnt, uint32_t n,
                              uint32_t *__restrict__ totals)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) totals[i] = valid_cnt[i] - dup_cnt[i];          // (mod 2^32: the ranks' partial values sum to the total)
}

// ---------------------------------------------------------------------------------------------------------------
// finish: thresholds, ordered output
// ---------------------------------------------------------------------------------------------------------------
struct EmitParams {
    uint32_t min_kmers;
    double min_ident;           // < 0: no ani test (partial counts)
    int k;
    int gbits;
    uint32_t world, rank;       // multi-GPU: owner(g) = g % world
    // deferred bubbles (null: none): presence lists per genome (CSR over set ids), the sets, total bubble weight per genome
    const uint32_t *pres_off, *pres, *bub_W;
    const BubSet *bub_sets;
    uint32_t heavy_thr;
};

// what the bubbles add to the pair (r, c): nothing for two heavy genomes (their pairs were expanded), else the weights of
// the sets that contain both
__device__ __forceinline__ uint32_t bubble_common(const EmitParams &ep, uint32_t r, uint32_t c)
{
    if (!ep.pres_off) return 0;
    if (ep.bub_W[r] >= ep.heavy_thr && ep.bub_W[c] >= ep.heavy_thr) return 0;
    const uint32_t a0 = ep.pres_off[r], a1 = ep.pres_off[r + 1], b0 = ep.pres_off[c], b1 = ep.pres_off[c + 1];
    uint32_t sum = 0;
    for (uint32_t x = a0; x < a1; ++x) {
        const uint32_t id = ep.pres[x];
        for (uint32_t y = b0; y < b1; ++y)
            if (ep.pres[y] == id) { sum += ep.bub_sets[id].w; break; }
    }
    return sum;
}
constexpr double ANI_MARGIN = 1e-9;         // |device ani - host ani| is ~1e-16; anything closer to the threshold is re-checked on the host
constexpr uint32_t BORDERLINE = 0x80000000u;

// ani-shorter (params.cpp:28-32) with the device's log(): returns 2 = passes for sure, 1 = borderline, 0 = fails
__device__ __forceinline__ int ani_test(uint32_t common, uint32_t tr, uint32_t tc, const EmitParams &ep, float &ani_out)
{
    const double j = (double)common / (double)(tr < tc ? tr : tc);
    const double d = (j == 0) ? 1.0 : (-1.0 / ep.k) * log((2 * j) / (j + 1));
    const double a = 1.0 - d;
    ani_out = (float)a;
    if (ep.min_ident < 0) return 2;
    if (a >= ep.min_ident + ANI_MARGIN) return 2;
    if (a >= ep.min_ident - ANI_MARGIN) return 1;
    return 0;
}

__device__ __forceinline__ void tri_decode(uint64_t t, uint32_t &row, uint32_t &col)
{
    uint32_t r = (uint32_t)((1.0 + sqrt(1.0 + 8.0 * (double)t)) * 0.5);
    while ((uint64_t)r * (r - 1) / 2 > t) --r;
    while ((uint64_t)(r + 1) * r / 2 <= t) ++r;
    row = r; col = (uint32_t)(t - (uint64_t)r * (r - 1) / 2);
}

// One rank, dense layout: ordered compaction of the triangular counter array.  pass 0: passing entries per block;
// pass 1 (after a scan of the block counts): write them in index order = sorted by (row, col).
__global__ void __launch_bounds__(256) dense_emit_kernel(const uint32_t *__restrict__ dense, uint64_t n_entries, uint64_t per_block,
                                                         const uint32_t *__restrict__ totals, EmitParams ep, int write,
                                                         uint32_t *__restrict__ block_cnt, uint64_t *__restrict__ out_keys,
                                                         uint32_t *__restrict__ out_vals, float *__restrict__ out_ani,
                                                         unsigned long long *__restrict__ n_border)
{
    __shared__ uint32_t warp_cnt[8];
    __shared__ uint32_t s_run;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint64_t lo = blockIdx.x * per_block, hi = min(n_entries, lo + per_block);
    if (threadIdx.x == 0) s_run = write ? block_cnt[blockIdx.x] : 0;        // pass 1: block_cnt holds the exclusive offsets
    __syncthreads();
    for (uint64_t base = lo; base < hi; base += 256) {
        const uint64_t t = base + threadIdx.x;
        int ok = 0;
        uint32_t v = 0, row = 0, col = 0;
        float ani = 0;
        if (t < hi) {
            v = dense[t];
            if (v >= ep.min_kmers && v > 0) {
                tri_decode(t, row, col);
                ok = ani_test(v, totals[row], totals[col], ep, ani);
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, ok != 0);
        if (lane == 0) warp_cnt[wid] = __popc(m);
        __syncthreads();
        uint32_t before = 0, all = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) { uint32_t c = warp_cnt[j]; if (j < wid) before += c; all += c; }
        if (ok && write) {
            const uint32_t o = s_run + before + __popc(m & ((1u << lane) - 1));
            out_keys[o] = ((uint64_t)row << 32) | col;
            out_vals[o] = v | (ok == 1 ? BORDERLINE : 0u);
            out_ani[o] = ani;
            if (ok == 1) atomicAdd(n_border, 1ULL);
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

// Accumulator -> unordered list of compact entries (key = row << gbits | col, value = count).
//   world == 1: entries that can still pass (count >= min_kmers, ani test not failed) -> destination 0
//   world  > 1: every non-zero partial count, once for each distinct owner of its two genomes
// pass 0 counts per destination into dest_cnt[world]; pass 1 appends at dest_cur[d] (pre-set to the region starts).
__global__ void __launch_bounds__(256) acc_emit_kernel(PairAcc A, uint64_t n_entries, const uint32_t *__restrict__ totals, EmitParams ep,
                                                       int write, unsigned long long *__restrict__ dest_cur,
                                                       uint64_t *__restrict__ out_keys, uint32_t *__restrict__ out_vals, uint4 *__restrict__ out_rec)
{
    const int lane = threadIdx.x & 31;
    // n_entries rounded up to whole warps by the loop condition: every lane of a warp takes part in the ballots
    for (uint64_t i0 = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) & ~31ULL; i0 < n_entries; i0 += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = i0 + lane;
        uint32_t v = 0, row = 0, col = 0;
        if (i < n_entries) {
            if (A.dense) { v = A.dense[i]; if (v) tri_decode(i, row, col); }
            else { const PairSlot e = A.slots[i]; if (e.key != SLOT_EMPTY) { v = e.val; row = (uint32_t)(e.key >> 32); col = (uint32_t)e.key; } }
        }
        uint32_t d1 = 0xffffffffu, d2 = 0xffffffffu;
        if (v) {
            if (ep.world == 1) {
                float ani;
                if (ep.pres_off || (v >= ep.min_kmers && ani_test(v, totals[row], totals[col], ep, ani))) d1 = 0;
            } else {
                d1 = row % ep.world; d2 = col % ep.world;
                if (d2 == d1) d2 = 0xffffffffu;
            }
        }
        for (uint32_t d = 0; d < ep.world; ++d) {
            const bool mine = d1 == d || d2 == d;
            const unsigned m = __ballot_sync(0xffffffffu, mine);
            if (!m) continue;
            const int leader = __ffs(m) - 1;
            unsigned long long base = 0;
            if (lane == leader) base = atomicAdd(&dest_cur[d], (unsigned long long)__popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (mine && write) {
                const unsigned long long o = base + __popc(m & ((1u << lane) - 1));
                const uint64_t ck = ((uint64_t)row << ep.gbits) | col;
                if (out_rec) out_rec[o] = make_uint4((uint32_t)ck, (uint32_t)(ck >> 32), v, 0u);      // one 16-byte record: one all-to-all
                else { out_keys[o] = ck; out_vals[o] = v; }
            }
        }
    }
}

__global__ void fill_u64_kernel(uint64_t *p, uint64_t n, uint64_t v)
{
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = v;
}

// Sorted compact entries (equal keys adjacent: the partial counts of one pair from several ranks / passes) -> final
// list.  The head of every run sums the run, applies the thresholds and, if it passes, is written in order.
// pass 0: passing heads per block; pass 1: write (block_cnt holds exclusive offsets).  mine_only: keep only pairs whose
// row this rank owns (the list reported to rank 0).
__global__ void __launch_bounds__(256) finalize_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals, uint64_t n,
                                                       uint64_t per_block, const uint32_t *__restrict__ totals, EmitParams ep, int write,
                                                       int mine_only, uint32_t *__restrict__ block_cnt, uint64_t *__restrict__ out_keys,
                                                       uint32_t *__restrict__ out_vals, float *__restrict__ out_ani,
                                                       unsigned long long *__restrict__ n_border)
{
    __shared__ uint32_t warp_cnt[8];
    __shared__ uint32_t s_run;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint64_t lo = blockIdx.x * per_block, hi = min(n, lo + per_block);
    const uint64_t cmask = (1ULL << ep.gbits) - 1;
    if (threadIdx.x == 0) s_run = write ? block_cnt[blockIdx.x] : 0;
    __syncthreads();
    for (uint64_t base = lo; base < hi; base += 256) {
        const uint64_t t = base + threadIdx.x;
        int ok = 0;
        uint32_t sum = 0, row = 0, col = 0;
        float ani = 0;
        if (t < hi) {
            const uint64_t k = keys[t];
            if (k != KEY_SENTINEL && (t == 0 || keys[t - 1] != k)) {
                uint64_t s = 0;
                for (uint64_t j = t; j < n && keys[j] == k; ++j) s += vals[j];
                row = (uint32_t)(k >> ep.gbits); col = (uint32_t)(k & cmask);
                sum = (uint32_t)s + bubble_common(ep, row, col);
                if (sum >= ep.min_kmers && sum > 0 && (!mine_only || row % ep.world == ep.rank))
                    ok = ani_test(sum, totals[row], totals[col], ep, ani);
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, ok != 0);
        if (lane == 0) warp_cnt[wid] = __popc(m);
        __syncthreads();
        uint32_t before = 0, all = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) { uint32_t c = warp_cnt[j]; if (j < wid) before += c; all += c; }
        if (ok && write) {
            const uint32_t o = s_run + before + __popc(m & ((1u << lane) - 1));
            out_keys[o] = ((uint64_t)row << 32) | col;
            out_vals[o] = sum | (ok == 1 ? BORDERLINE : 0u);
            out_ani[o] = ani;
            if (ok == 1) atomicAdd(n_border, 1ULL);
        }
        __syncthreads();
        if (threadIdx.x == 0) s_run += all;
        __syncthreads();
    }
    if (!write && threadIdx.x == 0) block_cnt[blockIdx.x] = s_run;
}

// 16-byte exchange records {key, value} <-> the sort's key / value arrays (small exchanges travel as ONE all-to-all)
__global__ void unpack_rec_kernel(const uint4 *__restrict__ rec, uint64_t n, uint64_t *__restrict__ keys, uint32_t *__restrict__ vals)
{
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 r = rec[i];
        keys[i] = ((uint64_t)r.y << 32) | r.x; vals[i] = r.z;
    }
}
// final list (row << 32 | col, value) -> records with the compact sort key (row << gbits | col)
__global__ void pack_final_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals, uint64_t n, int gbits, uint4 *__restrict__ rec)
{
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t ck = ((keys[i] >> 32) << gbits) | (uint32_t)keys[i];
        rec[i] = make_uint4((uint32_t)ck, (uint32_t)(ck >> 32), vals[i], 0u);
    }
}

// final key (row << 32 | col) -> compact sort key (row << gbits | col)
__global__ void compact_keys_kernel(const uint64_t *__restrict__ in, uint64_t n, int gbits, uint64_t *__restrict__ out)
{
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        out[i] = ((in[i] >> 32) << gbits) | (uint32_t)in[i];
}

int grid_for(uint64_t n, int threads = 256, int max_blocks = 148 * 16)
{
    return (int)std::max<uint64_t>(1, std::min<uint64_t>((n + threads - 1) / threads, (uint64_t)max_blocks));
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
struct PoolScope {                       // buffers allocated inside live in the stream-ordered pool, not in the call's arena
    bool saved;
    PoolScope() : saved(vb_tls_pool_alloc) { vb_tls_pool_alloc = true; }
    ~PoolScope() { vb_tls_pool_alloc = saved; }
};

void comm_check(int rc, const char *what)
{
    if (rc != 0) throw vb_error(VB_ERR_INTERNAL, std::string("collective failed: ") + what);
}

// exclusive scan on the device, n <= 2^20 (see the kernels above)
void dev_exscan(vb_ctx *ctx, cudaStream_t st, const uint32_t *in, uint32_t n, uint32_t *out, uint32_t *out2, uint32_t *scratch /* 1025 */)
{
    const uint32_t chunks = (n + 1023) / 1024;
    if (chunks > 1024) throw vb_error(VB_ERR_INTERNAL, "dev_exscan: too many counters");
    exscan_chunks_kernel<<<chunks, 1024, 0, st>>>(in, n, out, scratch);
    VB_LAUNCH_CHECK(ctx);
    exscan_totals_kernel<<<1, 1024, 0, st>>>(scratch, chunks, out + n);
    VB_LAUNCH_CHECK(ctx);
    exscan_add_kernel<<<chunks, 1024, 0, st>>>(out, n, scratch, out2);
    VB_LAUNCH_CHECK(ctx);
}

struct Accumulator {
    PairAcc acc = {nullptr, nullptr, 0, nullptr};
    DevBuf<uint32_t> dense;
    DevBuf<PairSlot> slots;
    uint64_t cap = 0;                    // dense: N(N-1)/2 entries; hashed: slots
    bool is_dense() const { return acc.dense != nullptr; }
};

void acc_alloc_hashed(vb_ctx *ctx, cudaStream_t st, Accumulator &a, uint64_t cap, unsigned long long *status)
{
    if (cap > (1ULL << 33)) throw vb_error(VB_ERR_MEM, "pair table would exceed 2^33 slots");
    PoolScope pool;
    a.slots.alloc(cap);
    table_clear_kernel<<<(int)std::min<uint64_t>((cap + 255) / 256, 148 * 16), 256, 0, st>>>(a.slots.p, cap);
    VB_LAUNCH_CHECK(ctx);
    a.cap = cap;
    a.acc = {nullptr, a.slots.p, cap - 1, status};
}

// grow the hashed table to new_cap slots, keeping its contents
void acc_grow(vb_ctx *ctx, cudaStream_t st, Accumulator &a, uint64_t new_cap)
{
    Accumulator b;
    acc_alloc_hashed(ctx, st, b, new_cap, a.acc.status);
    rehash_kernel<<<grid_for(a.cap), 256, 0, st>>>(a.slots.p, a.cap, b.acc);
    VB_LAUNCH_CHECK(ctx);
    a.slots = std::move(b.slots);
    a.cap = b.cap; a.acc = b.acc;
}

struct FinalList {                       // the device-side result of a prefilter call
    uint64_t n = 0;
    DevBuf<uint64_t> keys;               // row << 32 | col, sorted
    DevBuf<uint32_t> vals;               // common | BORDERLINE
    DevBuf<float> ani;
};

// sorted compact entries -> FinalList (ordered compaction in two passes)
void finalize_sorted(vb_ctx *ctx, cudaStream_t st, const uint64_t *skeys, const uint32_t *svals, uint64_t n, const uint32_t *totals,
                     const EmitParams &em, bool mine_only, unsigned long long *d_scalar, unsigned long long *d_border, FinalList &out, bool pool)
{
    out.n = 0;
    if (!n) return;
    const uint32_t n_blocks = (uint32_t)std::min<uint64_t>(4096, (n + 255) / 256);
    const uint64_t per_block = ((n + n_blocks - 1) / n_blocks + 255) / 256 * 256;
    DevBuf<uint32_t> block_cnt(4096);
    finalize_kernel<<<n_blocks, 256, 0, st>>>(skeys, svals, n, per_block, totals, em, 0, mine_only ? 1 : 0, block_cnt.p, nullptr, nullptr, nullptr, d_border);
    VB_LAUNCH_CHECK(ctx);
    scan_blocks_kernel<<<1, 1024, 0, st>>>(block_cnt.p, n_blocks, d_scalar);
    VB_LAUNCH_CHECK(ctx);
    unsigned long long n_out = 0;
    VB_CUDA(cudaMemcpyAsync(&n_out, d_scalar, sizeof(n_out), cudaMemcpyDeviceToHost, st));
    VB_CUDA(cudaStreamSynchronize(st));
    out.n = n_out;
    if (!n_out) return;
    {
        const bool saved = vb_tls_pool_alloc;
        vb_tls_pool_alloc = pool;
        try { out.keys.alloc(n_out); out.vals.alloc(n_out); out.ani.alloc(n_out); } catch (...) { vb_tls_pool_alloc = saved; throw; }
        vb_tls_pool_alloc = saved;
    }
    finalize_kernel<<<n_blocks, 256, 0, st>>>(skeys, svals, n, per_block, totals, em, 1, mine_only ? 1 : 0, block_cnt.p, out.keys.p, out.vals.p, out.ani.p, d_border);
    VB_LAUNCH_CHECK(ctx);
}

// unordered compact entries (n real ones in buffers of n_pad) -> sorted; returns pointers into a/b
struct SortBufs { DevBuf<uint64_t> ka, kb; DevBuf<uint32_t> va, vb; };
void sort_entries(vb_ctx *ctx, cudaStream_t st, SortBufs &sb, uint64_t n, uint64_t n_pad, int key_bits, rsort::Workspace &ws,
                  const uint64_t *&skeys, const uint32_t *&svals)
{
    if (n_pad > n) {
        fill_u64_kernel<<<grid_for(n_pad - n), 256, 0, st>>>(sb.ka.p + n, n_pad - n, KEY_SENTINEL);
        VB_LAUNCH_CHECK(ctx);
    }
    const bool in_b = rsort::sort_kv<8>(ctx, sb.ka.p, sb.va.p, sb.kb.p, sb.vb.p, n_pad, key_bits, ws);
    skeys = in_b ? sb.kb.p : sb.ka.p;
    svals = in_b ? sb.vb.p : sb.va.p;
}

}  // namespace

vb_pairs *vb_pairs_alloc(uint64_t n, uint32_t n_genomes);

void vb_drop_dev_pairs(vb_ctx *ctx)
{
    if (!ctx->dev_pairs) return;
    cudaStream_t saved = vb_tls_stream;
    vb_tls_stream = (cudaStream_t)ctx->stream;
    delete ctx->dev_pairs;
    ctx->dev_pairs = nullptr;
    vb_tls_stream = saved;
}

void vb_prefilter_run(vb_ctx *ctx, const vb_prefilter_job &job, const vb_prefilter_params *p, vb_pairs **out_pairs)
{
    const vb_genomes *g = job.g;
    const vb_comm *comm = job.comm;
    const uint32_t world = comm ? (uint32_t)comm->world : 1u, rank = comm ? (uint32_t)comm->rank : 0u;
    const bool partial = job.shard_count > 1;                   // vb_prefilter_partial: raw counts of one k-mer shard
    if (job.shard_count == 0 || job.shard_index >= job.shard_count) throw vb_error(VB_ERR_ARG, "bad k-mer shard");
    if (p->k < 10 || p->k > 31) throw vb_error(VB_ERR_ARG, "k must be in [10, 31]");
    if (!(p->kmers_fraction > 0)) throw vb_error(VB_ERR_ARG, "kmers_fraction must be > 0");
    if (world > 1 && (partial || p->max_seqs > 0)) throw vb_error(VB_ERR_ARG, "k-mer shards / --max-seqs are not available in the multi-GPU pipeline");
    if (world > 1024) throw vb_error(VB_ERR_ARG, "more than 1024 ranks");
    cudaStream_t st = (cudaStream_t)ctx->stream;
    VB_CUDA(cudaSetDevice(ctx->device));
    ctx->clear_timings("prefilter.");
    vb_drop_dev_pairs(ctx);
    const uint32_t n_local = g->count();
    const uint32_t n = job.n_total ? job.n_total : n_local;    // genomes of the whole set
    if ((uint64_t)job.gid_base + n_local > n) throw vb_error(VB_ERR_ARG, "genome block outside the set");
    EventTimer t_all(st), t_up(st);
    double ms_ext = 0, ms_part = 0, ms_group = 0, ms_exch = 0;
    double peer_bytes = 0;                   // bytes this rank copied into other ranks' receive buffers (NVLink)
    // stage timers are read at the end of the call (reading one waits for its stop event)
    std::vector<std::pair<std::unique_ptr<EventTimer>, double *>> laps;
    auto lap_start = [&](double &acc_ms) -> EventTimer * {
        laps.emplace_back(std::make_unique<EventTimer>(st), &acc_ms);
        laps.back().first->start();
        return laps.back().first.get();
    };

    t_all.start();
    DevBuf<uint32_t> counters(3 * (size_t)std::max<uint32_t>(n, 1) + 1);       // valid | dup | totals (+ 1: overflow flag of all ranks)
    uint32_t *valid_cnt = counters.p, *dup_cnt = counters.p + n, *totals = counters.p + 2 * (size_t)n;
    VB_CUDA(cudaMemsetAsync(counters.p, 0, counters.bytes(), st));
    ExtractParams ep;
    ep.k = p->k;
    ep.shift = 0; ep.tail_mask = 0;
    if (2 * p->k - 32 < 8) { ep.shift = 8 - (2 * p->k - 32); ep.tail_mask = (1ULL << ep.shift) - 1; }
    ep.use_filter = p->kmers_fraction < 1.0;
    ep.max_thr = (uint64_t)((double)UINT64_MAX * (0.0 + p->kmers_fraction));
    ep.c = (uint64_t)std::ceil((double)p->k / 4);
    ep.gid_base = job.gid_base;
    // status words: [0] table overflow, [1] keys in the hashed table, [2] bubbles found, [3] scratch counter,
    // [4] screen: tuples that found their slot marked, [5] screen: slots marked twice, [6] collect cursor, [7] scratch
    DevBuf<unsigned long long> scalars(16);
    VB_CUDA(cudaMemsetAsync(scalars.p, 0, scalars.bytes(), st));
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, ctx->device);

    // ---- plan: passes and level-1 buckets, from sizes every rank knows
    const double frac = std::min(1.0, p->kmers_fraction);
    const double est_local = (double)vb_store_slots(g, VB_STORE_PAD) * frac / job.shard_count;
    const double est_all = std::max(est_local, job.est_kmers_all * frac / job.shard_count);
    const char *pe = getenv("VB_PREFILTER_PASSES");
    // one pass groups at most 2^20 buckets of ~1 280 tuples (over all ranks) and must fit the device: 36 B per tuple
    const double per_pass_cap = std::min(1.0e9, 0.5 * (double)ctx->mem_total / 36.0 * world);
    uint32_t passes = pe ? (uint32_t)std::max(1, atoi(pe)) : (uint32_t)std::ceil(est_all / per_pass_cap);
    passes = std::max(1u, std::min(passes, 4096u));
    const double est_pass = est_all / passes;

    // singleton screen (one rank, small inputs): >= 4 slots per expected k-mer, at most 2^28 slots (64 MB, L2-resident)
    const char *seen_env = getenv("VB_PREFILTER_SEEN");                    // "0": off, "N": force 2^N slots
    int seen_bits = 20;
    while ((double)(1ULL << seen_bits) < 4.0 * est_pass && seen_bits < 28) ++seen_bits;
    bool use_seen = world == 1 && est_pass <= (double)(1ULL << 27);
    if (seen_env && world == 1) { int v = atoi(seen_env); use_seen = v > 0; if (v >= 10 && v <= 34) seen_bits = v; }

    uint32_t Bper, B1;
    {
        const double est_keep = use_seen ? 0.5 * est_pass : est_pass;       // the screen drops about half of the tuples
        int bits = 0;
        while (est_keep / (double)(1ULL << bits) > BUCKET_TARGET && bits < 2 * MAX_BUCKET_BITS) ++bits;
        // parents over all ranks: the two levels get (about) the same fan-out -- a tile of 4 096 tuples then leaves runs of
        // the same length in both; an odd bit goes to level 2, whose fan-out need not be a power of two
        const char *l1_env = getenv("VB_PREFILTER_L1BITS");                  // test hook
        uint32_t want = 1u << std::min(MAX_BUCKET_BITS, l1_env ? std::max(0, atoi(l1_env)) : bits / 2);
        Bper = 1;
        while (Bper * 2 * world <= std::max(want, world) && Bper * 2 * world <= (1u << MAX_BUCKET_BITS)) Bper *= 2;
        B1 = Bper * world;
    }
    const uint32_t n_fine = B1 << FINE_BITS;

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
    ep.shard_count = job.shard_count * passes;
    ep.shard_index = job.shard_index * passes;                  // pass 0
    auto screen_range = [&](const DevGenomes &d, uint64_t lo_all, uint64_t hi_all) {
        for (uint64_t lo = lo_all; lo < hi_all; lo += chunk) {
            const uint64_t hi = std::min(hi_all, lo + chunk);
            screen_kernel<<<grid_of(lo, hi, 8), 256, 0, st>>>(d.seq2.p, d.inv_kdb.p, d.tile_gid.p, lo, hi, ep, seen, valid_cnt,
                                                               scalars.p + 4);
            VB_LAUNCH_CHECK(ctx);
        }
    };
    // when the genomes have to be uploaded, the screen pass (of pass 0) over each chunk of genomes is enqueued while the
    // next chunk is still on the PCIe bus
    bool screened = false;
    // Larger inputs have no screen; there the whole extraction of pass 0 rides on the upload instead: the tuples of every
    // chunk of genomes are collected (into buffers sized for every slot, so no counting pass) while the next chunk is on
    // the bus, and only the last chunk's extraction is left when the copy ends.
    const uint64_t slots_pre = vb_store_slots(g, VB_STORE_PAD);
    auto exact_alloc_for = [&](uint64_t slots) {
        return passes > 1 || job.shard_count > 1 || slots >= (1ULL << 31) || 12.0 * (double)slots > 0.125 * (double)ctx->mem_total ||
               getenv("VB_PREFILTER_EXACT") != nullptr;
    };
    const bool early_collect = !use_seen && !exact_alloc_for(slots_pre) && slots_pre <= chunk && !vb_has_dev_genomes(ctx, g, VB_STORE_PAD) &&
                               getenv("VB_PREFILTER_NO_EARLY") == nullptr;                  // (test hook: extraction after the upload)
    DevBuf<uint32_t> fine_early, vals_early;
    DevBuf<uint64_t> keys_early;
    bool collected_early = false;
    if (early_collect) {
        fine_early.alloc(n_fine);
        VB_CUDA(cudaMemsetAsync(fine_early.p, 0, fine_early.bytes(), st));
        keys_early.alloc(slots_pre + 64);
        vals_early.alloc(slots_pre + 64);
    }
    const vb_chunk_fn hook = [&](const DevGenomes &d, uint64_t lo, uint64_t hi) {
        if (use_seen) { screen_range(d, lo, hi); screened = true; }
        else if (early_collect && hi > lo && d.total_slots == slots_pre) {
            ep.shard_count = job.shard_count * passes;
            ep.shard_index = job.shard_index * passes;
            collect_kernel<true><<<grid_of(lo, hi, 8), 256, 0, st>>>(d.seq2.p, d.inv_kdb.p, d.tile_gid.p, lo, hi, ep, seen, valid_cnt,
                                                                      scalars.p + 6, 1, keys_early.p, vals_early.p, fine_early.p, B1);
            VB_LAUNCH_CHECK(ctx);
            collected_early = true;
        }
    };
    t_up.start();
    const DevGenomes &dg = vb_get_dev_genomes(ctx, g, VB_STORE_PAD, nullptr, &hook);
    t_up.stop();
    const uint64_t n_slots = dg.total_slots;

    // ---- the accumulator of all passes
    const unsigned long long max_pairs = (unsigned long long)n * (n > 0 ? n - 1 : 0) / 2;
    const bool force_hash = getenv("VB_PREFILTER_HASH") != nullptr;   // test hook: exercise the large-N layout
    const bool want_dense = !force_hash && max_pairs <= (1ULL << 26);
    const uint64_t cap_env = getenv("VB_PREFILTER_TABLE") ? strtoull(getenv("VB_PREFILTER_TABLE"), nullptr, 10) : 0;   // test hook: initial slots
    uint64_t cap_hint = 1024;
    if (cap_env) { while (cap_hint < cap_env) cap_hint <<= 1; }
    else if (ctx->pair_hint_uid == g->uid && ctx->pair_hint_n == n && ctx->pair_hint_k == p->k && ctx->pair_hint_entries > 0) {
        // the same genome set again: size the table for the number of distinct pairs the last call saw (load <= 0.4), so
        // that it is as cache-friendly as it can be; a wrong guess costs a grow or, at worst, a redo
        while ((double)cap_hint < 2.5 * (double)ctx->pair_hint_entries) cap_hint <<= 1;
    } else {
        const double guess = std::min((double)max_pairs * 2.0, std::max(1048576.0, est_all / world / 4.0));
        while ((double)cap_hint < guess) cap_hint <<= 1;
    }
    rsort::Workspace ws;
    static std::mutex attr_mutex;
    static bool attr_done_dev[64] = {false};                 // the opt-in is per device (one process may hold several contexts)
    const size_t part_smem = PART_TILE * 12 + 4 * (1 << MAX_BUCKET_BITS) * sizeof(uint32_t);
    {
        std::lock_guard<std::mutex> lock(attr_mutex);
        bool &attr_done = attr_done_dev[ctx->device & 63];
        if (!attr_done) {
            VB_CUDA(cudaFuncSetAttribute(part_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)part_smem));
            VB_CUDA(cudaFuncSetAttribute(part_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)part_smem));
            VB_CUDA(cudaFuncSetAttribute(bucket_flat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BucketSmem)));
            VB_CUDA(cudaFuncSetAttribute(bucket_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ChainSmem)));
            attr_done = true;
        }
    }

    // grouping kernel: chains for small families, flat pair enumeration for large ones (see bucket_chain_kernel)
    // The flat kernel pays when groups are large AND every increment is a dependent look-up in the hashed table (a thread
    // that walks a chain of 200 serialises 200 global round trips); known from the previous call on the same set: more
    // than 32 distinct pairs per genome.  c3 (families of 20, dense counters): chains 8.4 ms, flat 13.2 ms.
    const char *bk_env = getenv("VB_PREFILTER_BUCKET");
    const bool hint_ok = ctx->pair_hint_uid == g->uid && ctx->pair_hint_n == n && ctx->pair_hint_k == p->k;
    const bool flat_buckets = bk_env ? strcmp(bk_env, "flat") == 0 : (hint_ok && ctx->pair_hint_entries > 32ULL * n);
    unsigned long long n_survivors_all = 0, n_tuples_grouped = 0, n_bubbles = 0;
    Accumulator A;
    // deferred bubbles (hashed accumulator): member sets of all passes, kept until the thresholds are applied
    DevBuf<uint32_t> bub_members;            // the store: concatenated member lists
    uint64_t bub_used = 0;
    std::vector<BubSet> bub_sets;            // this rank's sets (offsets into bub_members)
    DevBuf<uint32_t> bub_all_members, bub_pres_off, bub_pres, bub_W;   // after the passes: all ranks' sets, presence lists
    DevBuf<BubSet> bub_all_sets;
    bool bub_active = false;
    for (int attempt = 0;; ++attempt) {
    // (a retry re-runs all passes into a larger table: an insert that found no slot is lost, so nothing can be salvaged)
    if (attempt > 0) {
        VB_CUDA(cudaMemsetAsync(counters.p, 0, counters.bytes(), st));
        VB_CUDA(cudaMemsetAsync(scalars.p, 0, scalars.bytes(), st));
        if (use_seen) VB_CUDA(cudaMemsetAsync(seen_words.p, 0, seen_words.bytes(), st));
        screened = false;
        n_survivors_all = 0; n_tuples_grouped = 0; n_bubbles = 0;
        bub_used = 0; bub_sets.clear(); bub_active = false;
    }
    if (want_dense) {
        if (!A.dense.p) { PoolScope pool; A.dense.alloc(std::max<unsigned long long>(max_pairs, 1)); }
        VB_CUDA(cudaMemsetAsync(A.dense.p, 0, A.dense.bytes(), st));
        A.acc = {A.dense.p, nullptr, 0, scalars.p};
        A.cap = max_pairs;
    } else {
        A.slots.release();
        acc_alloc_hashed(ctx, st, A, cap_hint, scalars.p);
    }

    for (uint32_t pass = 0; pass < passes; ++pass) {
        ep.shard_index = job.shard_index * passes + pass;
        // ---- extract: hash once, (optionally) screen out singletons, compact list + fine histogram
        EventTimer *t_ext = lap_start(ms_ext);
        const bool have_early = collected_early && attempt == 0 && pass == 0;    // this pass' tuples were collected during the upload
        DevBuf<uint32_t> fine_hist;
        if (have_early) fine_hist = std::move(fine_early);
        else {
            fine_hist.alloc(n_fine);
            VB_CUDA(cudaMemsetAsync(fine_hist.p, 0, fine_hist.bytes(), st));
        }
        unsigned long long *d_cursor = scalars.p + 6;
        if (use_seen && pass > 0) VB_CUDA(cudaMemsetAsync(seen_words.p, 0, seen_words.bytes(), st));
        // small inputs: room for every slot's tuple, no counting pass.  Large inputs / several passes: count first.
        const bool exact_alloc = !have_early && exact_alloc_for(n_slots);
        auto for_chunks = [&](auto &&launch) {
            for (uint64_t lo = 0; lo < n_slots; lo += chunk) launch(lo, std::min(n_slots, lo + chunk));
        };
        bool counted_valid = pass > 0;                           // valid k-mers per genome are counted by the first full scan of every pass' shard
        // (every pass sees only its own k-mers, so each pass counts the valid k-mers of its shard; they add up)
        counted_valid = false;
        if (use_seen) {
            if (!(screened && pass == 0)) screen_range(dg, 0, n_slots);   // (pass 0: already enqueued chunk by chunk during the upload)
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
                                                                              valid_cnt, d_cursor, 0, nullptr, nullptr, nullptr, B1);
                    VB_LAUNCH_CHECK(ctx);
                });
                counted_valid = true;
                VB_CUDA(cudaMemcpyAsync(cnt, d_cursor, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
                VB_CUDA(cudaMemsetAsync(d_cursor, 0, sizeof(unsigned long long), st));
                VB_CUDA(cudaStreamSynchronize(st));
                list_cap = cnt[0] + 64;
            }
        }
        if (list_cap >= (1ULL << 32)) throw vb_error(VB_ERR_ARG, "more than 2^32 k-mer tuples in one prefilter pass: raise VB_PREFILTER_PASSES or use more GPUs");
        DevBuf<uint64_t> keys0;                              // compact list of surviving (hash, genome) tuples
        DevBuf<uint32_t> vals0;
        if (have_early) { keys0 = std::move(keys_early); vals0 = std::move(vals_early); }
        else { keys0.alloc(list_cap); vals0.alloc(list_cap); }
        if (!have_early) for_chunks([&](uint64_t lo, uint64_t hi) {
            if (counted_valid)
                collect_kernel<false><<<grid_of(lo, hi, 8), 256, 0, st>>>(dg.seq2.p, dg.inv_kdb.p, dg.tile_gid.p, lo, hi, ep, seen, valid_cnt,
                                                                           d_cursor, 1, keys0.p, vals0.p, fine_hist.p, B1);
            else
                collect_kernel<true><<<grid_of(lo, hi, 8), 256, 0, st>>>(dg.seq2.p, dg.inv_kdb.p, dg.tile_gid.p, lo, hi, ep, seen, valid_cnt,
                                                                          d_cursor, 1, keys0.p, vals0.p, fine_hist.p, B1);
            VB_LAUNCH_CHECK(ctx);
        });
        // level-1 histogram (tuples per parent, all ranks' parents) -> offsets of the level-1 / send buffer
        DevBuf<uint32_t> hist1(B1), off1(B1 + 1), cursor1(B1), scan_tmp(1025);
        coarsen_kernel<<<(B1 * 32 + 255) / 256, 256, 0, st>>>(fine_hist.p, 1, 0, 1, B1, hist1.p);
        VB_LAUNCH_CHECK(ctx);
        dev_exscan(ctx, st, hist1.p, B1, off1.p, cursor1.p, scan_tmp.p);
        // every rank's offsets, gathered (one rank: just its own), and the survivor count: one synchronisation
        std::vector<uint32_t> h_off1_all((size_t)world * (B1 + 1));
        DevBuf<uint32_t> off1_all(world > 1 ? (size_t)world * (B1 + 1) : 1);
        if (world > 1) {
            comm_check(comm->all_gather(comm->user, off1.p, off1_all.p, sizeof(uint32_t) * (B1 + 1)), "all_gather(level-1 offsets)");
            VB_CUDA(cudaMemcpyAsync(h_off1_all.data(), off1_all.p, sizeof(uint32_t) * h_off1_all.size(), cudaMemcpyDeviceToHost, st));
        } else
            VB_CUDA(cudaMemcpyAsync(h_off1_all.data(), off1.p, sizeof(uint32_t) * (B1 + 1), cudaMemcpyDeviceToHost, st));
        VB_CUDA(cudaStreamSynchronize(st));
        const uint32_t *my_off1 = h_off1_all.data() + (size_t)rank * (B1 + 1);
        const uint64_t n_keep = my_off1[B1];
        n_survivors_all += n_keep;
        t_ext->stop();

        // ---- level 1: tuples -> parents; ordered by parent, the level-1 buffer is the send buffer of all-to-all #1 (several
        // ranks): every parent's tuples travel to the rank that owns the parent.
        EventTimer *t_part = lap_start(ms_part);
        const uint32_t tiles1 = (uint32_t)((n_keep + PART_TILE - 1) / PART_TILE);
        const uint64_t fine_slice = (uint64_t)Bper << FINE_BITS;                 // fine-histogram entries per destination rank
        std::vector<uint64_t> sc(world, 0), rc(world, 0), recv_of(world, 0);     // tuples I send to / receive from d; all tuples d receives
        for (uint32_t d = 0; d < world; ++d) {
            sc[d] = my_off1[(d + 1) * Bper] - my_off1[d * Bper];
            for (uint32_t s2 = 0; s2 < world; ++s2) {
                const uint32_t *o = h_off1_all.data() + (size_t)s2 * (B1 + 1);
                const uint64_t c = o[(d + 1) * Bper] - o[d * Bper];
                recv_of[d] += c;
                if (d == rank) rc[s2] = c;
            }
        }
        uint64_t n_recv = world > 1 ? recv_of[rank] : n_keep;
        if (n_recv >= (1ULL << 32)) throw vb_error(VB_ERR_ARG, "more than 2^32 k-mer tuples arrive at one rank in one pass: raise VB_PREFILTER_PASSES");
        DevBuf<uint64_t> keys1(n_keep + 64);
        DevBuf<uint32_t> vals1(n_keep + 64);
        if (tiles1) {
            part_kernel<1><<<std::min<uint32_t>(tiles1, n_sm * 8), PART_THREADS, part_smem, st>>>(
                n_keep, B1, 1, keys0.p, vals0.p, nullptr, nullptr, tiles1, cursor1.p, keys1.p, vals1.p);
            VB_LAUNCH_CHECK(ctx);
        }
        t_part->stop();
        // Two ways to move the tuples: NCCL all-to-alls, or -- when the ranks' receive buffers are mapped into each other's
        // address space (vb_peer_xbuf, CUDA IPC) -- one peer copy per destination straight out of the level-1 buffer (a
        // destination's parents are contiguous there) over NVLink, with a one-word all-reduce as the barrier behind them.
        // Layout of a receive buffer: [keys of all segments][vals of all segments] ... [fine histograms: world slices] at the end.
        bool use_peer = world > 1 && job.xbuf != nullptr;
        uint64_t fine_at = 0;
        if (use_peer) {
            fine_at = (job.xbuf->cap - 4 * fine_slice * world) / 16 * 16;
            for (uint32_t d = 0; d < world; ++d)
                if (4 * fine_slice * world + 64 > job.xbuf->cap || 12 * recv_of[d] + 64 > fine_at) use_peer = false;   // (every rank decides alike)
        }
        auto vals_base = [&](uint64_t n_all) { return (8 * n_all + 15) / 16 * 16; };       // byte offset of the vals region in a receive buffer

        const uint64_t *rkeys = keys1.p;
        const uint32_t *rvals = vals1.p;
        DevBuf<uint64_t> keysR;
        DevBuf<uint32_t> valsR, fineR, d_flag(1);
        const uint32_t *fine_src = fine_hist.p + ((size_t)rank * Bper << FINE_BITS);
        uint32_t fine_n_src = 1;
        std::vector<Segment> segs;
        if (world > 1) {
            EventTimer *t_x = lap_start(ms_exch);
            if (use_peer) {
                // destinations in a rotated order, so that the ranks do not all write to the same peer at the same time
                for (uint32_t t = 0; t < world; ++t) {
                    const uint32_t d = (rank + 1 + t) % world;
                    uint64_t before = 0;                                        // tuples of the ranks before me in d's buffer
                    for (uint32_t s2 = 0; s2 < rank; ++s2) {
                        const uint32_t *o = h_off1_all.data() + (size_t)s2 * (B1 + 1);
                        before += o[(d + 1) * Bper] - o[d * Bper];
                    }
                    char *base = (char *)job.xbuf->peer[d];
                    if (d != rank) peer_bytes += 12.0 * (double)sc[d] + 4.0 * (double)fine_slice;
                    if (sc[d]) {
                        VB_CUDA(cudaMemcpyAsync(base + 8 * before, keys1.p + my_off1[d * Bper], 8 * sc[d], cudaMemcpyDeviceToDevice, st));
                        VB_CUDA(cudaMemcpyAsync(base + vals_base(recv_of[d]) + 4 * before, vals1.p + my_off1[d * Bper], 4 * sc[d], cudaMemcpyDeviceToDevice, st));
                    }
                    VB_CUDA(cudaMemcpyAsync(base + fine_at + 4 * fine_slice * rank, fine_hist.p + fine_slice * d, 4 * fine_slice, cudaMemcpyDeviceToDevice, st));
                }
                // the all-reduce returns once every rank's copies are behind it
                VB_CUDA(cudaMemsetAsync(d_flag.p, 0, sizeof(uint32_t), st));
                comm_check(comm->all_reduce_sum_u32(comm->user, d_flag.p, 1), "all_reduce(exchange barrier)");
                rkeys = (const uint64_t *)job.xbuf->local;
                rvals = (const uint32_t *)((const char *)job.xbuf->local + vals_base(n_recv));
                fine_src = (const uint32_t *)((const char *)job.xbuf->local + fine_at);
            } else {
                keysR.alloc(n_recv + 64); valsR.alloc(n_recv + 64);
                comm_check(comm->all_to_all(comm->user, keys1.p, sc.data(), keysR.p, rc.data(), 8), "all_to_all(tuple hashes)");
                comm_check(comm->all_to_all(comm->user, vals1.p, sc.data(), valsR.p, rc.data(), 4), "all_to_all(tuple genomes)");
                // the fine histograms of my parents, from every rank
                fineR.alloc((size_t)world * fine_slice);
                std::vector<uint64_t> fc(world, fine_slice);
                comm_check(comm->all_to_all(comm->user, fine_hist.p, fc.data(), fineR.p, fc.data(), 4), "all_to_all(bucket histograms)");
                fine_src = fineR.p;
                rkeys = keysR.p; rvals = valsR.p;
            }
            fine_n_src = world;
            uint64_t at = 0;
            for (uint32_t s2 = 0; s2 < world; ++s2) {                     // segments: (source rank, parent)
                const uint32_t *o = h_off1_all.data() + (size_t)s2 * (B1 + 1);
                for (uint32_t j = 0; j < Bper; ++j) {
                    const uint32_t len = o[rank * Bper + j + 1] - o[rank * Bper + j];
                    if (len) segs.push_back({(uint32_t)at, len, j, 0});
                    at += len;
                }
            }
            t_x->stop();
        } else {
            for (uint32_t j = 0; j < B1; ++j) {
                const uint32_t len = my_off1[j + 1] - my_off1[j];
                if (len) segs.push_back({my_off1[j], len, j, 0});
            }
        }
        n_tuples_grouped += n_recv;
        uint32_t tiles2 = 0;
        for (auto &sg : segs) { sg.tile_start = tiles2; tiles2 += (sg.len + PART_TILE - 1) / PART_TILE; }
        std::vector<uint32_t> tile_seg(tiles2);                               // tile -> its segment
        for (size_t i = 0; i < segs.size(); ++i)
            std::fill(tile_seg.begin() + segs[i].tile_start, tile_seg.begin() + (i + 1 < segs.size() ? segs[i + 1].tile_start : tiles2), (uint32_t)i);

        // ---- level 2: parents -> final buckets
        t_part = lap_start(ms_part);
        const uint32_t B2 = (uint32_t)std::min<uint64_t>(1u << MAX_BUCKET_BITS, std::max<uint64_t>(1, (n_recv / Bper + BUCKET_TARGET - 1) / BUCKET_TARGET));
        const uint32_t NB = Bper * B2;
        DevBuf<uint32_t> hist(NB), off(NB + 1), cursor2(NB), big_list(NB + 1), n_big(1);
        DevBuf<Segment> d_segs(std::max<size_t>(segs.size(), 1));
        VB_CUDA(cudaMemsetAsync(n_big.p, 0, sizeof(uint32_t), st));
        DevBuf<uint32_t> d_tile_seg(std::max<size_t>(tile_seg.size(), 1));
        if (!segs.empty()) VB_CUDA(cudaMemcpyAsync(d_segs.p, segs.data(), sizeof(Segment) * segs.size(), cudaMemcpyHostToDevice, st));
        if (tiles2) VB_CUDA(cudaMemcpyAsync(d_tile_seg.p, tile_seg.data(), sizeof(uint32_t) * tiles2, cudaMemcpyHostToDevice, st));
        coarsen_kernel<<<(int)(((uint64_t)NB * 32 + 255) / 256), 256, 0, st>>>(fine_src, fine_n_src, (uint64_t)Bper << FINE_BITS, B2, NB, hist.p);
        VB_LAUNCH_CHECK(ctx);
        dev_exscan(ctx, st, hist.p, NB, off.p, cursor2.p, scan_tmp.p);
        DevBuf<uint64_t> keys2(n_recv + 64);
        DevBuf<uint32_t> vals2(n_recv + 64);
        if (tiles2) {
            part_kernel<2><<<std::min<uint32_t>(tiles2, n_sm * 8), PART_THREADS, part_smem, st>>>(
                n_recv, B1, B2, rkeys, rvals, d_segs.p, d_tile_seg.p, tiles2, cursor2.p, keys2.p, vals2.p);
            VB_LAUNCH_CHECK(ctx);
        }
        t_part->stop();

        // ---- group: shared-memory chains per bucket, oversized buckets, bubbles
        EventTimer *t_grp = lap_start(ms_group);
        const uint32_t huge_cap = 1u << 16;
        DevBuf<HugeRun> huge_list(huge_cap);
        VB_CUDA(cudaMemsetAsync(scalars.p + 2, 0, sizeof(unsigned long long), st));
        const int bgrid = (int)std::min<uint32_t>(NB, n_sm * 12);
        if (flat_buckets)
            bucket_flat_kernel<<<bgrid, 256, sizeof(BucketSmem), st>>>(keys2.p, vals2.p, off.p, NB, dup_cnt, A.acc, big_list.p, n_big.p);
        else
            bucket_chain_kernel<<<bgrid, 256, sizeof(ChainSmem), st>>>(keys2.p, vals2.p, off.p, NB, dup_cnt, A.acc, big_list.p, n_big.p);
        VB_LAUNCH_CHECK(ctx);
        big_bucket_kernel<<<64, 1024, 0, st>>>(keys2.p, vals2.p, off.p, big_list.p, n_big.p, dup_cnt, A.acc, huge_list.p, huge_cap, scalars.p + 2);
        VB_LAUNCH_CHECK(ctx);
        unsigned long long h_status[3] = {0, 0, 0};
        VB_CUDA(cudaMemcpyAsync(h_status, scalars.p, sizeof(h_status), cudaMemcpyDeviceToHost, st));
        VB_CUDA(cudaStreamSynchronize(st));
        if (h_status[2] > huge_cap) throw vb_error(VB_ERR_INTERNAL, "more than 65536 bubble k-mers in one pass");
        if (h_status[2] > 0) {
            // bubbles: compact the member sets, collapse identical ones, expand each set once with its multiplicity
            const uint32_t nh = (uint32_t)h_status[2];
            n_bubbles += nh;
            DevBuf<uint32_t> members(n_recv + 64);
            DevBuf<HugeInfo> d_info(nh);
            bubble_prepare_kernel<<<std::min<uint32_t>(nh, n_sm * 2), 1024, 0, st>>>(vals2.p, huge_list.p, nh, members.p, d_info.p, dup_cnt);
            VB_LAUNCH_CHECK(ctx);
            std::vector<HugeInfo> info(nh);
            VB_CUDA(cudaMemcpyAsync(info.data(), d_info.p, sizeof(HugeInfo) * nh, cudaMemcpyDeviceToHost, st));
            VB_CUDA(cudaStreamSynchronize(st));
            std::vector<uint32_t> order(nh), rep(nh);
            for (uint32_t i = 0; i < nh; ++i) order[i] = i;
            std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
                if (info[a].m != info[b].m) return info[a].m < info[b].m;
                if (info[a].sig1 != info[b].sig1) return info[a].sig1 < info[b].sig1;
                if (info[a].sig2 != info[b].sig2) return info[a].sig2 < info[b].sig2;
                return a < b;
            });
            for (uint32_t i = 0; i < nh; ++i) {
                const uint32_t a = order[i];
                rep[a] = a;
                if (i > 0) {
                    const uint32_t b = order[i - 1];
                    if (info[a].m == info[b].m && info[a].sig1 == info[b].sig1 && info[a].sig2 == info[b].sig2) rep[a] = rep[b];
                }
            }
            const bool no_collapse = getenv("VB_PREFILTER_NO_COLLAPSE") != nullptr;              // test hook
            if (no_collapse) for (uint32_t i = 0; i < nh; ++i) rep[i] = i;
            DevBuf<uint32_t> d_rep(nh), d_differ(nh);
            std::vector<uint32_t> differ(nh, 0);
            VB_CUDA(cudaMemcpyAsync(d_rep.p, rep.data(), sizeof(uint32_t) * nh, cudaMemcpyHostToDevice, st));
            VB_CUDA(cudaMemsetAsync(d_differ.p, 0, sizeof(uint32_t) * nh, st));
            bubble_equal_kernel<<<dim3(8, std::min<uint32_t>(nh, 32768)), 256, 0, st>>>(members.p, d_info.p, d_rep.p, nh, d_differ.p);
            VB_LAUNCH_CHECK(ctx);
            VB_CUDA(cudaMemcpyAsync(differ.data(), d_differ.p, sizeof(uint32_t) * nh, cudaMemcpyDeviceToHost, st));
            VB_CUDA(cudaStreamSynchronize(st));
            std::vector<uint32_t> weight(nh, 0);
            for (uint32_t i = 0; i < nh; ++i) { if (differ[i]) rep[i] = i; weight[rep[i]]++; }   // (a signature collision: expand on its own)
            std::vector<BubbleJob> jobs;
            uint64_t tiles = 0;
            for (uint32_t i = 0; i < nh; ++i) {
                if (!weight[i] || info[i].m < 2) continue;
                const uint64_t T = (info[i].m + 127) / 128;
                if (A.is_dense() && tiles + T * (T + 1) / 2 >= (1ULL << 32)) throw vb_error(VB_ERR_ARG, "bubble k-mers too large for one pass");
                jobs.push_back({info[i].beg, info[i].m, weight[i], (uint32_t)tiles});
                tiles += T * (T + 1) / 2;
            }
            if (!A.is_dense()) {
                // hashed accumulator: the sets are kept, their pairs are resolved when the thresholds are applied (see BubSet)
                uint64_t need = bub_used;
                for (auto &jb : jobs) need += jb.m;
                if (need > bub_members.n) {
                    PoolScope pool;
                    DevBuf<uint32_t> bigger(std::max<uint64_t>(need, 2 * bub_members.n) + 1024);
                    if (bub_used) VB_CUDA(cudaMemcpyAsync(bigger.p, bub_members.p, sizeof(uint32_t) * bub_used, cudaMemcpyDeviceToDevice, st));
                    VB_CUDA(cudaStreamSynchronize(st));
                    bub_members = std::move(bigger);
                }
                for (auto &jb : jobs) {
                    VB_CUDA(cudaMemcpyAsync(bub_members.p + bub_used, members.p + jb.beg, sizeof(uint32_t) * jb.m, cudaMemcpyDeviceToDevice, st));
                    bub_sets.push_back({(uint32_t)bub_used, jb.m, jb.weight});
                    bub_used += jb.m;
                }
                if (bub_used >= (1ULL << 32)) throw vb_error(VB_ERR_ARG, "more than 2^32 bubble members on one rank");
                VB_CUDA(cudaStreamSynchronize(st));           // members is released below
            } else if (!jobs.empty()) {
                DevBuf<BubbleJob> d_jobs(jobs.size());
                VB_CUDA(cudaMemcpyAsync(d_jobs.p, jobs.data(), sizeof(BubbleJob) * jobs.size(), cudaMemcpyHostToDevice, st));
                bubble_pairs_kernel<<<(unsigned)std::min<uint64_t>(tiles, (uint64_t)n_sm * 16), 256, 0, st>>>(members.p, d_jobs.p, (uint32_t)jobs.size(),
                                                                                                              (uint32_t)tiles, A.acc);
                VB_LAUNCH_CHECK(ctx);
                VB_CUDA(cudaStreamSynchronize(st));           // d_jobs / members are released below
            }
        }
        // the table is grown between passes when it is more than half full (never inside a pass)
        if (!A.is_dense() && pass + 1 < passes && h_status[1] > A.cap / 2) acc_grow(ctx, st, A, A.cap * 4);
        VB_CUDA(cudaMemsetAsync(scalars.p + 4, 0, 3 * sizeof(unsigned long long), st));     // screen / collect counters of the next pass
        t_grp->stop();
    }

    // ---- deferred bubbles: every rank learns all sets; presence lists; pairs of heavy genomes expanded (see BubSet)
    if (!A.is_dense()) {
        unsigned long long mine[2] = {bub_sets.size(), bub_used};
        std::vector<unsigned long long> cnts(2 * (size_t)world, 0);
        cnts[0] = mine[0]; cnts[1] = mine[1];
        if (world > 1) {
            DevBuf<unsigned long long> d_mine(2), d_cnts(2 * (size_t)world);
            VB_CUDA(cudaMemcpyAsync(d_mine.p, mine, sizeof(mine), cudaMemcpyHostToDevice, st));
            comm_check(comm->all_gather(comm->user, d_mine.p, d_cnts.p, sizeof(mine)), "all_gather(bubble counts)");
            VB_CUDA(cudaMemcpyAsync(cnts.data(), d_cnts.p, sizeof(unsigned long long) * cnts.size(), cudaMemcpyDeviceToHost, st));
            VB_CUDA(cudaStreamSynchronize(st));
        }
        uint64_t max_sets = 0, max_mem = 0, all_sets = 0;
        for (uint32_t r = 0; r < world; ++r) { max_sets = std::max<uint64_t>(max_sets, cnts[2 * r]); max_mem = std::max<uint64_t>(max_mem, cnts[2 * r + 1]); all_sets += cnts[2 * r]; }
        if (all_sets) {
            bub_active = true;
            if ((uint64_t)world * max_mem >= (1ULL << 32)) throw vb_error(VB_ERR_ARG, "more than 2^32 bubble members");
            PoolScope pool;
            // members of all ranks, rank r's at r * max_mem; set descriptors rebuilt on the host with global offsets
            std::vector<BubSet> sets_all;
            uint32_t my_first = 0;
            bub_all_members.alloc((uint64_t)world * max_mem + 1);
            if (world > 1) {
                DevBuf<uint32_t> pad_mem(max_mem + 1), desc_send(3 * max_sets + 3), desc_all((uint64_t)world * (3 * max_sets + 3));
                VB_CUDA(cudaMemsetAsync(pad_mem.p, 0, pad_mem.bytes(), st));
                if (bub_used) VB_CUDA(cudaMemcpyAsync(pad_mem.p, bub_members.p, sizeof(uint32_t) * bub_used, cudaMemcpyDeviceToDevice, st));
                std::vector<uint32_t> h_desc(3 * max_sets + 3, 0);
                for (size_t i = 0; i < bub_sets.size(); ++i) { h_desc[3 * i] = bub_sets[i].off; h_desc[3 * i + 1] = bub_sets[i].m; h_desc[3 * i + 2] = bub_sets[i].w; }
                VB_CUDA(cudaMemcpyAsync(desc_send.p, h_desc.data(), sizeof(uint32_t) * h_desc.size(), cudaMemcpyHostToDevice, st));
                comm_check(comm->all_gather(comm->user, pad_mem.p, bub_all_members.p, sizeof(uint32_t) * max_mem), "all_gather(bubble members)");
                comm_check(comm->all_gather(comm->user, desc_send.p, desc_all.p, sizeof(uint32_t) * (3 * max_sets + 3)), "all_gather(bubble sets)");
                std::vector<uint32_t> h_all((size_t)world * (3 * max_sets + 3));
                VB_CUDA(cudaMemcpyAsync(h_all.data(), desc_all.p, sizeof(uint32_t) * h_all.size(), cudaMemcpyDeviceToHost, st));
                VB_CUDA(cudaStreamSynchronize(st));
                for (uint32_t r = 0; r < world; ++r) {
                    if (r == rank) my_first = (uint32_t)sets_all.size();
                    const uint32_t *d = h_all.data() + (size_t)r * (3 * max_sets + 3);
                    for (uint64_t i = 0; i < cnts[2 * r]; ++i) sets_all.push_back({d[3 * i] + (uint32_t)(r * max_mem), d[3 * i + 1], d[3 * i + 2]});
                }
            } else {
                VB_CUDA(cudaMemcpyAsync(bub_all_members.p, bub_members.p, sizeof(uint32_t) * bub_used, cudaMemcpyDeviceToDevice, st));
                sets_all = bub_sets;
            }
            const uint32_t ns = (uint32_t)sets_all.size();
            bub_all_sets.alloc(ns);
            VB_CUDA(cudaMemcpyAsync(bub_all_sets.p, sets_all.data(), sizeof(BubSet) * ns, cudaMemcpyHostToDevice, st));
            // presence lists: sets per genome (CSR) and the genomes' total bubble weights
            if ((uint64_t)n + 1 > (1u << 20)) throw vb_error(VB_ERR_ARG, "bubble k-mers with more than 2^20 genomes are not supported");
            bub_pres_off.alloc((size_t)n + 2);
            bub_W.alloc(std::max<uint32_t>(n, 1));
            DevBuf<uint32_t> cnt(std::max<uint32_t>(n, 1)), fill(std::max<uint32_t>(n, 1)), scan_tmp2(1025);
            VB_CUDA(cudaMemsetAsync(cnt.p, 0, cnt.bytes(), st));
            VB_CUDA(cudaMemsetAsync(fill.p, 0, fill.bytes(), st));
            VB_CUDA(cudaMemsetAsync(bub_W.p, 0, bub_W.bytes(), st));
            const int sgrid = (int)std::min<uint32_t>(ns, (uint32_t)n_sm * 8);
            bub_count_kernel<<<sgrid, 256, 0, st>>>(bub_all_members.p, bub_all_sets.p, ns, cnt.p, bub_W.p);
            VB_LAUNCH_CHECK(ctx);
            dev_exscan(ctx, st, cnt.p, n, bub_pres_off.p, nullptr, scan_tmp2.p);
            uint32_t total_pres = 0;
            VB_CUDA(cudaMemcpyAsync(&total_pres, bub_pres_off.p + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            VB_CUDA(cudaStreamSynchronize(st));
            bub_pres.alloc((size_t)total_pres + 1);
            bub_fill_kernel<<<sgrid, 256, 0, st>>>(bub_all_members.p, bub_all_sets.p, ns, bub_pres_off.p, fill.p, bub_pres.p);
            VB_LAUNCH_CHECK(ctx);
            // pairs of two heavy genomes: expansion of this rank's own sets, restricted to their heavy members
            const uint32_t heavy_thr = partial ? 1u : (uint32_t)std::max(p->min_kmers, 1);
            const uint32_t n_own = (uint32_t)bub_sets.size();
            if (n_own) {
                DevBuf<uint32_t> heavy_members((uint64_t)world * max_mem + 1), m_out(n_own);
                bub_filter_kernel<<<std::min<uint32_t>(n_own, (uint32_t)n_sm * 2), 1024, 0, st>>>(bub_all_members.p, bub_all_sets.p, my_first, n_own, bub_W.p,
                                                                                               heavy_thr, heavy_members.p, m_out.p);
                VB_LAUNCH_CHECK(ctx);
                std::vector<uint32_t> h_m(n_own);
                unsigned long long entries_now = 0;
                VB_CUDA(cudaMemcpyAsync(h_m.data(), m_out.p, sizeof(uint32_t) * n_own, cudaMemcpyDeviceToHost, st));
                VB_CUDA(cudaMemcpyAsync(&entries_now, scalars.p + 1, sizeof(entries_now), cudaMemcpyDeviceToHost, st));
                VB_CUDA(cudaStreamSynchronize(st));
                std::vector<BubbleJob> jobs;
                uint64_t tiles = 0, new_pairs = 0;
                for (uint32_t q = 0; q < n_own; ++q) {
                    if (h_m[q] < 2) continue;
                    const uint64_t T = (h_m[q] + 127) / 128;
                    if (tiles + T * (T + 1) / 2 >= (1ULL << 32)) throw vb_error(VB_ERR_ARG, "bubble k-mers too large to expand");
                    jobs.push_back({sets_all[my_first + q].off, h_m[q], sets_all[my_first + q].w, (uint32_t)tiles});
                    tiles += T * (T + 1) / 2;
                    new_pairs += (uint64_t)h_m[q] * (h_m[q] - 1) / 2;
                }
                if (!jobs.empty()) {
                    uint64_t need = 1024;                        // make room: every pair may be new
                    while (need < 2 * (entries_now + new_pairs)) need <<= 1;
                    if (need > A.cap) acc_grow(ctx, st, A, need);
                    DevBuf<BubbleJob> d_jobs(jobs.size());
                    VB_CUDA(cudaMemcpyAsync(d_jobs.p, jobs.data(), sizeof(BubbleJob) * jobs.size(), cudaMemcpyHostToDevice, st));
                    bubble_pairs_kernel<<<(unsigned)std::min<uint64_t>(tiles, (uint64_t)n_sm * 16), 256, 0, st>>>(heavy_members.p, d_jobs.p, (uint32_t)jobs.size(),
                                                                                                                  (uint32_t)tiles, A.acc);
                    VB_LAUNCH_CHECK(ctx);
                    VB_CUDA(cudaStreamSynchronize(st));
                }
            }
        }
    }

    // ---- totals (all ranks: all-reduce) and the overflow flag of all ranks
    totals_kernel<<<(n + 255) / 256 + 1, 256, 0, st>>>(valid_cnt, dup_cnt, n, totals);
    VB_LAUNCH_CHECK(ctx);
    unsigned long long h_over[2] = {0, 0};
    if (world > 1) {
        VB_CUDA(cudaMemcpyAsync(h_over + 1, scalars.p + 1, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        VB_CUDA(cudaMemcpyAsync(totals + n, scalars.p, sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));     // low word of the flag
        comm_check(comm->all_reduce_sum_u32(comm->user, totals, (uint64_t)n + 1), "all_reduce(total k-mers)");
        uint32_t any = 0;
        VB_CUDA(cudaMemcpyAsync(&any, totals + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        VB_CUDA(cudaStreamSynchronize(st));
        h_over[0] = any;
    } else {
        VB_CUDA(cudaMemcpyAsync(h_over, scalars.p, sizeof(h_over), cudaMemcpyDeviceToHost, st));
        VB_CUDA(cudaStreamSynchronize(st));
    }
    if (!h_over[0]) {
        ctx->pair_hint_uid = g->uid; ctx->pair_hint_n = n; ctx->pair_hint_k = p->k;
        ctx->pair_hint_entries = A.is_dense() ? 0 : h_over[1];
        break;
    }
    if (attempt >= 6) throw vb_error(VB_ERR_MEM, "pair table overflow");
    cap_hint = A.cap * 8;
    }   // attempts

    // ---- finish
    EventTimer t_emit(st);
    t_emit.start();
    EmitParams em;
    em.min_kmers = partial ? 1u : (uint32_t)std::max(p->min_kmers, 0);
    em.min_ident = partial ? -1.0 : std::max(p->min_ident, 0.0);
    em.k = p->k;
    em.gbits = 1;
    while ((1ULL << em.gbits) < n) em.gbits++;
    em.world = world; em.rank = rank;
    em.pres_off = bub_active ? bub_pres_off.p : nullptr;
    em.pres = bub_pres.p; em.bub_W = bub_W.p; em.bub_sets = bub_all_sets.p;
    em.heavy_thr = partial ? 1u : (uint32_t)std::max(p->min_kmers, 1);
    FinalList fin;                       // pairs that involve this rank's genomes (one rank: all pairs)
    const bool keep_dev = job.keep_dev && !partial && p->max_seqs <= 0;
    if (A.is_dense() && world == 1) {
        // dense layout: ordered compaction, the output is born sorted by (row, col)
        const uint32_t n_blocks = (uint32_t)std::min<uint64_t>(4096, (max_pairs + 255) / 256 + 1);
        const uint64_t per_block = ((max_pairs + n_blocks - 1) / n_blocks + 255) / 256 * 256;
        DevBuf<uint32_t> block_cnt(4096);
        dense_emit_kernel<<<n_blocks, 256, 0, st>>>(A.dense.p, max_pairs, per_block, totals, em, 0, block_cnt.p, nullptr, nullptr, nullptr, scalars.p + 8);
        VB_LAUNCH_CHECK(ctx);
        scan_blocks_kernel<<<1, 1024, 0, st>>>(block_cnt.p, n_blocks, scalars.p + 7);
        VB_LAUNCH_CHECK(ctx);
        unsigned long long n_emit = 0;
        VB_CUDA(cudaMemcpyAsync(&n_emit, scalars.p + 7, sizeof(n_emit), cudaMemcpyDeviceToHost, st));
        VB_CUDA(cudaStreamSynchronize(st));
        fin.n = n_emit;
        if (n_emit) {
            { const bool saved = vb_tls_pool_alloc; vb_tls_pool_alloc = keep_dev;
              try { fin.keys.alloc(n_emit); fin.vals.alloc(n_emit); fin.ani.alloc(n_emit); } catch (...) { vb_tls_pool_alloc = saved; throw; }
              vb_tls_pool_alloc = saved; }
            dense_emit_kernel<<<n_blocks, 256, 0, st>>>(A.dense.p, max_pairs, per_block, totals, em, 1, block_cnt.p, fin.keys.p, fin.vals.p, fin.ani.p, scalars.p + 8);
            VB_LAUNCH_CHECK(ctx);
        }
    } else {
        // accumulator -> per-destination lists of compact entries (count, then write)
        DevBuf<unsigned long long> dest_cur(world);
        VB_CUDA(cudaMemsetAsync(dest_cur.p, 0, sizeof(unsigned long long) * world, st));
        const uint64_t n_entries = A.cap;
        acc_emit_kernel<<<grid_for(n_entries), 256, 0, st>>>(A.acc, n_entries, totals, em, 0, dest_cur.p, nullptr, nullptr, nullptr);
        VB_LAUNCH_CHECK(ctx);
        std::vector<unsigned long long> sc(world), all_counts((size_t)world * world, 0);
        if (world > 1) {
            DevBuf<unsigned long long> d_all((size_t)world * world);
            comm_check(comm->all_gather(comm->user, dest_cur.p, d_all.p, sizeof(unsigned long long) * world), "all_gather(pair counts)");
            VB_CUDA(cudaMemcpyAsync(all_counts.data(), d_all.p, sizeof(unsigned long long) * all_counts.size(), cudaMemcpyDeviceToHost, st));
            VB_CUDA(cudaStreamSynchronize(st));
        } else {
            VB_CUDA(cudaMemcpyAsync(all_counts.data(), dest_cur.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
            VB_CUDA(cudaStreamSynchronize(st));
        }
        uint64_t n_send = 0, n_recv = 0;
        std::vector<unsigned long long> starts(world);
        std::vector<uint64_t> scv(world), rcv(world);
        for (uint32_t d = 0; d < world; ++d) {
            scv[d] = all_counts[(size_t)rank * world + d]; starts[d] = n_send; n_send += scv[d];
            rcv[d] = all_counts[(size_t)d * world + rank]; n_recv += rcv[d];
        }
        if (n_recv >= (1ULL << 32) - rsort::TILE || n_send >= (1ULL << 32)) throw vb_error(VB_ERR_ARG, "more than 2^32 partial pair counts at one rank");
        const uint64_t n_pad = (n_recv + rsort::TILE - 1) / rsort::TILE * rsort::TILE;
        SortBufs sb;
        DevBuf<uint4> send_rec(world > 1 ? n_send + 1 : 1), recv_rec(world > 1 ? n_recv + 1 : 1);
        sb.ka.alloc(n_pad + 1); sb.kb.alloc(n_pad + 1); sb.va.alloc(n_pad + 1); sb.vb.alloc(n_pad + 1);
        VB_CUDA(cudaMemcpyAsync(dest_cur.p, starts.data(), sizeof(unsigned long long) * world, cudaMemcpyHostToDevice, st));
        acc_emit_kernel<<<grid_for(n_entries), 256, 0, st>>>(A.acc, n_entries, totals, em, 1, dest_cur.p, sb.ka.p, sb.va.p,
                                                             world > 1 ? send_rec.p : nullptr);
        VB_LAUNCH_CHECK(ctx);
        if (world > 1) {
            EventTimer *t_x = lap_start(ms_exch);
            comm_check(comm->all_to_all(comm->user, send_rec.p, scv.data(), recv_rec.p, rcv.data(), 16), "all_to_all(partial pair counts)");
            if (n_recv) { unpack_rec_kernel<<<grid_for(n_recv), 256, 0, st>>>(recv_rec.p, n_recv, sb.ka.p, sb.va.p); VB_LAUNCH_CHECK(ctx); }
            t_x->stop();
        }
        const uint64_t *skeys = nullptr;
        const uint32_t *svals = nullptr;
        if (n_pad) sort_entries(ctx, st, sb, n_recv, n_pad, 2 * em.gbits, ws, skeys, svals);
        finalize_sorted(ctx, st, skeys, svals, n_pad, totals, em, false, scalars.p + 7, scalars.p + 8, fin, keep_dev);
    }
    A.dense.release(); A.slots.release();

    // ---- what goes to the host: one rank: everything; several ranks: the pairs whose row this rank owns, gathered on rank 0
    // (read-backs go through the context's page-locked staging buffer: pageable targets would make every copy synchronous)
    uint64_t n_host = 0;                 // pairs that go to the host
    char *pin = nullptr;
    uint64_t *h_keys = nullptr;
    uint32_t *h_vals = nullptr, *h_tot = nullptr;
    auto stage = [&](uint64_t n_out) {
        n_host = n_out;
        pin = (char *)vb_pinned(ctx, 12 * n_out + 4 * (size_t)n + 64);
        h_keys = (uint64_t *)pin; h_vals = (uint32_t *)(pin + 8 * n_out); h_tot = (uint32_t *)(pin + 12 * n_out);
        if (n) VB_CUDA(cudaMemcpyAsync(h_tot, totals, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, st));
    };
    if (world == 1) {
        stage(fin.n);
        if (fin.n) {
            VB_CUDA(cudaMemcpyAsync(h_keys, fin.keys.p, sizeof(uint64_t) * fin.n, cudaMemcpyDeviceToHost, st));
            VB_CUDA(cudaMemcpyAsync(h_vals, fin.vals.p, sizeof(uint32_t) * fin.n, cudaMemcpyDeviceToHost, st));
        }
    } else {
        // row-owned subset (still sorted) -> rank 0, which merges the ranks' runs with one more sort
        FinalList mine;
        {
            const uint64_t n_in = fin.n, n_pad = (n_in + rsort::TILE - 1) / rsort::TILE * rsort::TILE;
            SortBufs sb;
            sb.ka.alloc(n_pad + 1); sb.va.alloc(n_pad + 1);
            if (n_in) {
                compact_keys_kernel<<<grid_for(n_in), 256, 0, st>>>(fin.keys.p, n_in, em.gbits, sb.ka.p);
                VB_LAUNCH_CHECK(ctx);
                VB_CUDA(cudaMemcpyAsync(sb.va.p, fin.vals.p, sizeof(uint32_t) * n_in, cudaMemcpyDeviceToDevice, st));
            }
            EmitParams raw = em;
            raw.min_kmers = 0; raw.min_ident = -1.0;             // already thresholded: only select the rows this rank owns
            raw.pres_off = nullptr;
            finalize_sorted(ctx, st, sb.ka.p, sb.va.p, n_in, totals, raw, true, scalars.p + 7, scalars.p + 9, mine, false);
        }
        DevBuf<unsigned long long> d_cnt(1), d_all(world);
        const unsigned long long my_n = mine.n;
        VB_CUDA(cudaMemcpyAsync(d_cnt.p, &my_n, sizeof(my_n), cudaMemcpyHostToDevice, st));
        comm_check(comm->all_gather(comm->user, d_cnt.p, d_all.p, sizeof(unsigned long long)), "all_gather(result sizes)");
        std::vector<unsigned long long> cnts(world);
        VB_CUDA(cudaMemcpyAsync(cnts.data(), d_all.p, sizeof(unsigned long long) * world, cudaMemcpyDeviceToHost, st));
        VB_CUDA(cudaStreamSynchronize(st));
        uint64_t total = 0;
        std::vector<uint64_t> scv(world, 0), rcv(world, 0);
        scv[0] = mine.n;
        if (rank == 0) for (uint32_t s = 0; s < world; ++s) { rcv[s] = cnts[s]; total += cnts[s]; }
        if (total >= (1ULL << 32) - rsort::TILE) throw vb_error(VB_ERR_ARG, "more than 2^32 candidate pairs");
        const uint64_t n_pad = (total + rsort::TILE - 1) / rsort::TILE * rsort::TILE;
        SortBufs sb;
        sb.ka.alloc(n_pad + 1); sb.kb.alloc(n_pad + 1); sb.va.alloc(n_pad + 1); sb.vb.alloc(n_pad + 1);
        DevBuf<uint4> send_rec(mine.n + 1), recv_rec(total + 1);
        if (mine.n) { pack_final_kernel<<<grid_for(mine.n), 256, 0, st>>>(mine.keys.p, mine.vals.p, mine.n, em.gbits, send_rec.p); VB_LAUNCH_CHECK(ctx); }
        comm_check(comm->all_to_all(comm->user, send_rec.p, scv.data(), recv_rec.p, rcv.data(), 16), "all_to_all(result pairs)");
        if (total) { unpack_rec_kernel<<<grid_for(total), 256, 0, st>>>(recv_rec.p, total, sb.ka.p, sb.va.p); VB_LAUNCH_CHECK(ctx); }
        stage(rank == 0 ? total : 0);
        if (rank == 0 && total) {
            const uint64_t *skeys = nullptr;
            const uint32_t *svals = nullptr;
            sort_entries(ctx, st, sb, total, n_pad, 2 * em.gbits, ws, skeys, svals);
            VB_CUDA(cudaMemcpyAsync(h_keys, skeys, sizeof(uint64_t) * total, cudaMemcpyDeviceToHost, st));
            VB_CUDA(cudaMemcpyAsync(h_vals, svals, sizeof(uint32_t) * total, cudaMemcpyDeviceToHost, st));
            VB_CUDA(cudaStreamSynchronize(st));
            const uint64_t cmask = (1ULL << em.gbits) - 1;      // compact -> (row << 32 | col)
            const int gb = em.gbits;
            uint64_t *hk = h_keys;
            vb_parallel_for(total, 1 << 18, 16, [&](uint64_t lo, uint64_t hi) {
                for (uint64_t i = lo; i < hi; ++i) hk[i] = ((hk[i] >> gb) << 32) | (hk[i] & cmask);
            });
        }
    }
    unsigned long long *n_border_pin = (unsigned long long *)(pin + 12 * n_host + 4 * (size_t)n + 8 - (12 * n_host + 4 * (size_t)n) % 8);
    VB_CUDA(cudaMemcpyAsync(n_border_pin, scalars.p + 8, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    t_emit.stop();
    t_all.stop();
    VB_CUDA(cudaStreamSynchronize(st));
    const unsigned long long n_border_dev = *n_border_pin;
    if (world > 1 && n_border_dev && fin.n) {
        // a pair within 1e-9 of the ani threshold (practically never): decide it exactly on the host and rebuild this rank's list
        std::vector<uint64_t> fk(fin.n);
        std::vector<uint32_t> fv(fin.n);
        std::vector<float> fa(fin.n);
        VB_CUDA(cudaMemcpyAsync(fk.data(), fin.keys.p, sizeof(uint64_t) * fin.n, cudaMemcpyDeviceToHost, st));
        VB_CUDA(cudaMemcpyAsync(fv.data(), fin.vals.p, sizeof(uint32_t) * fin.n, cudaMemcpyDeviceToHost, st));
        VB_CUDA(cudaMemcpyAsync(fa.data(), fin.ani.p, sizeof(float) * fin.n, cudaMemcpyDeviceToHost, st));
        VB_CUDA(cudaStreamSynchronize(st));
        uint64_t w = 0;
        for (uint64_t i = 0; i < fin.n; ++i) {
            if (fv[i] & BORDERLINE) {
                const uint32_t r = (uint32_t)(fk[i] >> 32), c = (uint32_t)fk[i];
                if (!(vb_ani_shorter(fv[i] & ~BORDERLINE, h_tot[r], h_tot[c], p->k) >= p->min_ident)) continue;
            }
            fk[w] = fk[i]; fa[w] = fa[i]; ++w;
        }
        VB_CUDA(cudaMemcpyAsync(fin.keys.p, fk.data(), sizeof(uint64_t) * w, cudaMemcpyHostToDevice, st));
        VB_CUDA(cudaMemcpyAsync(fin.ani.p, fa.data(), sizeof(float) * w, cudaMemcpyHostToDevice, st));
        VB_CUDA(cudaStreamSynchronize(st));
        fin.n = w;
    }

    // ---- host: the exact IEEE-double metric (params.cpp:28-32) for the output, the borderline cases re-decided exactly
    const auto hp0 = std::chrono::steady_clock::now();
    const uint64_t n_emit = n_host;
    vb_pairs *res = vb_pairs_alloc(n_emit, n);
    std::vector<uint8_t> drop;
    bool any_border = false;
    for (uint64_t i = 0; i < n_emit && !any_border; ++i) any_border = (h_vals[i] & BORDERLINE) != 0;
    vb_parallel_for(n_emit, 4096, 32, [&](uint64_t lo, uint64_t hi) {
        for (uint64_t i = lo; i < hi; ++i) {
            const uint32_t r = (uint32_t)(h_keys[i] >> 32), c = (uint32_t)h_keys[i], v = h_vals[i] & ~BORDERLINE;
            res->row[i] = r; res->col[i] = c; res->common[i] = v;
            res->ani[i] = partial ? 0.0 : vb_ani_shorter(v, h_tot[r], h_tot[c], p->k);
        }
    });
    bool list_changed = false;
    if (any_border && !partial) {
        uint64_t w = 0;
        for (uint64_t i = 0; i < n_emit; ++i) {
            if ((h_vals[i] & BORDERLINE) && !(res->ani[i] >= p->min_ident)) { list_changed = true; continue; }
            res->row[w] = res->row[i]; res->col[w] = res->col[i]; res->common[w] = res->common[i]; res->ani[w] = res->ani[i]; ++w;
        }
        res->n_pairs = w;
    }
    // --max-seqs: the per-row sampler needs complete counts, so a k-mer shard leaves it to vb_pairs_merge
    if (!partial && p->max_seqs > 0) {
        std::vector<uint32_t> o_row(res->row, res->row + res->n_pairs), o_col(res->col, res->col + res->n_pairs),
            o_common(res->common, res->common + res->n_pairs);
        std::vector<double> o_ani(res->ani, res->ani + res->n_pairs);
        vb_sample_rows(n, (uint32_t)p->max_seqs, o_row, o_col, o_common, o_ani);
        vb_pairs *r2 = vb_pairs_alloc(o_row.size(), n);
        for (uint64_t o = 0; o < o_row.size(); ++o) { r2->row[o] = o_row[o]; r2->col[o] = o_col[o]; r2->common[o] = o_common[o]; r2->ani[o] = o_ani[o]; }
        vb_pairs_free(res);
        res = r2;
    }
    for (uint32_t i = 0; i < n; ++i) res->total_kmers[i] = h_tot[i];
    res->k = p->k;
    res->kmers_fraction = p->kmers_fraction;
    *out_pairs = res;
    // the device copy for the align stage (several ranks: this rank's share, whatever rank 0 reports)
    if (keep_dev && !(list_changed && world == 1)) {
        auto *dp = new DevPairs();
        dp->uid = vb_pairs_uid(res);
        dp->n = fin.n;
        dp->keys = std::move(fin.keys);
        dp->ani = std::move(fin.ani);
        ctx->dev_pairs = dp;
    }
    ctx->set_timing("prefilter.host_post_ms",
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - hp0).count());

    for (auto &l : laps) *l.second += l.first->ms();
    ctx->set_timing("prefilter.total_ms", t_all.ms());
    ctx->set_timing("prefilter.upload_pack_ms", t_up.ms());
    ctx->set_timing("prefilter.extract_ms", ms_ext);
    ctx->set_timing("prefilter.sort_ms", ms_part);
    ctx->set_timing("prefilter.segment_ms", ms_group);
    ctx->set_timing("prefilter.exchange_ms", ms_exch);
    ctx->set_timing("prefilter.peer_bytes", peer_bytes);
    ctx->set_timing("prefilter.emit_ms", t_emit.ms());
    ctx->set_timing("prefilter.passes", (double)passes);
    ctx->set_timing("prefilter.tuples", (double)dg.total_slots);
    ctx->set_timing("prefilter.survivors", (double)n_survivors_all);
    ctx->set_timing("prefilter.grouped", (double)n_tuples_grouped);
    ctx->set_timing("prefilter.bubbles", (double)n_bubbles);
    ctx->set_timing("prefilter.table_slots", (double)A.cap);
    ctx->set_timing("prefilter.candidates", (double)res->n_pairs);
}
