// vb_align: the LZ-ANI pairwise parse on one B200 -- one warp per directed (reference, query) pair.
//
// Reference computation (paths under /root/reference/3rd_party/lz-ani/src/):
//   reference text + indexes  CParser::prepare_reference  parser.cpp:16-34, :53-189
//   query text                CParser::prepare_data       parser.cpp:37-50
//   the parse                 CParser::parse              parser.cpp:482-716 (+ :192-449 helpers)
//   per-pair statistics       CParser::calc_stats         parser.cpp:734-783
//
// This is a re-formulation, not a transcription (DESIGN.md "align" has the derivation):
//   * texts are three bit planes (code bit 0, code bit 1, "not ACGT") in 16-byte records of 32 symbols; an N never
//     matches anything (the reference gets this from code 4 in the reference text vs code 5 in the query), so every
//     comparison is  (lo^lo') | (hi^hi') | N_q | N_r  on 32 bases -- two 128-bit loads per text, no bit shuffling.
//   * the reference's 4 MB open-addressing table of 11-mer positions becomes a per-reference table of
//     (fingerprint, position) slots; lookups take the max over ALL entries with the same k-mer, ties to the
//     smallest position, which is what the reference's ascending insertion + first-wins scan computes.
//   * the short-seed CSR (4^7 buckets) is not built at all: the reference only ever searches a window of
//     [pred - lit, pred + mrd) positions, which 32 lanes scan directly in the packed text.
//   * no factor list: calc_stats only needs per-component sums, kept as running state.
//   * the position-by-position search for the next seed is done 32 query positions at a time (one per lane);
//     the approximate extension consumes 1024 bases per warp step from mismatch bit masks.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <functional>
#include <mutex>

#include "dev_util.cuh"
#include "radix_sort.cuh"

namespace {

constexpr uint32_t HT_EMPTY = 0xffffffffu;

struct LzParams { int mal, msl, mrd, mqd, reg, aw, am, ar; };

struct RefDesc {
    uint64_t rec_off;    // record offset of the text in ref_rec (one uint4 per 32 symbols)
    uint64_t ht_off;     // slot offset of the anchor table
    uint32_t ht_cap;     // home slots (range reduction by multiply-shift); the table has ht_tail more slots behind them that
    uint32_t ht_tail;    // take the overflow of the last home slots (a chain wraps to slot 0 only beyond them)
    uint32_t pos_bits;   // a slot is fingerprint << pos_bits | position, 2^pos_bits > n
    uint32_t n;          // text length: 2*len + 3*mrd
    uint32_t len;        // genome length
    uint32_t gid;        // genome id in the packed store
};

// A text is an array of 16-byte records, one per 32 symbols: .x = bit 0 of the 32 codes (A0 C1 G2 T3), .y = bit 1,
// .z = "not ACGT" flags (bit b <-> symbol 32 c + b); .w is not used here.  Every text is followed by padding records
// that are all N, so reading up to 96 symbols past n needs no bounds check.
struct Text {
    const uint4 *rec;
    int n;
};

struct W3 { uint32_t lo, hi, nv; };

// the 32 symbols starting at position p >= 0: two 128-bit loads, three funnel shifts
__device__ __forceinline__ W3 fetch3(const uint4 *__restrict__ rec, uint64_t p)
{
    const uint4 a = __ldg(rec + (p >> 5)), b = __ldg(rec + (p >> 5) + 1);
    const unsigned sh = (unsigned)(p & 31);
    W3 w;
    w.lo = __funnelshift_r(a.x, b.x, sh); w.hi = __funnelshift_r(a.y, b.y, sh); w.nv = __funnelshift_r(a.z, b.z, sh);
    return w;
}

// k-mer of `len` <= 31 symbols as an integer: low plane in bits [0, 32), high plane in bits [32, 64)
__device__ __forceinline__ uint64_t kmer_code(const W3 &w, int len)
{
    const uint32_t m = (1u << len) - 1;
    return (uint64_t)(w.lo & m) | ((uint64_t)(w.hi & m) << 32);
}

__device__ __forceinline__ uint64_t fmix64(uint64_t k)
{
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}

// home slot of a hash in a table of `cap` slots (low 32 bits; the fingerprint comes from the high 32)
__device__ __forceinline__ uint32_t ht_slot(uint64_t h, uint32_t cap) { return __umulhi((uint32_t)h, cap); }

// Hash of a canonical anchor k-mer: low word -> home slot (its HIGH bits count: multiply-shift), high word -> fingerprint
// (its high bits count).  k-mers of up to 16 symbols fold into 32 bits and take two multiplicative hashes -- the index
// build hashes every reference position once per table partition, so the hash is on its critical path.
__device__ __forceinline__ uint64_t anchor_hash(uint64_t can, int len)
{
    if (len <= 16) {
        const uint32_t x = (uint32_t)can | ((uint32_t)(can >> 32) << 16);
        const uint32_t h1 = x * 0x9E3779B1u;
        const uint32_t h2 = (x ^ (x >> 15)) * 0x85EBCA77u;
        return ((uint64_t)h2 << 32) | h1;
    }
    return fmix64(can);
}

// min(K, reverse complement of K) under the integer order of kmer_code -- any fixed order works, the table is built and
// probed with the same rule.  Reverse complement in plane form: complement both planes, reverse the symbol order.
__device__ __forceinline__ uint64_t canonical_kmer(uint64_t code, int len)
{
    const uint32_t l = (uint32_t)code, h = (uint32_t)(code >> 32);
    const uint64_t rc = (uint64_t)(__brev(~l) >> (32 - len)) | ((uint64_t)(__brev(~h) >> (32 - len)) << 32);
    return code < rc ? code : rc;
}

// ---------------------------------------------------------------------------------------------------------------
// k5a: reference text  R = fwd | N^mrd | N^mrd | revcomp(fwd) | N^mrd   (parser.cpp:16-34), packed
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) build_ref_text_kernel(const uint4 *__restrict__ grec, const uint64_t *__restrict__ gofs,
                                                             const RefDesc *__restrict__ refs, uint32_t n_refs, int mrd,
                                                             uint4 *__restrict__ ref_rec, const uint8_t *__restrict__ is_ref)
{
    for (uint32_t r = blockIdx.y; r < n_refs; r += gridDim.y) {
        if (is_ref && !is_ref[r]) continue;
        const RefDesc d = refs[r];
        const uint4 *src = grec + (gofs[d.gid] >> 5);             // genomes start on 128-slot boundaries
        const uint32_t L = d.len;
        const uint32_t rc0 = L + 2 * (uint32_t)mrd;              // first position of the reverse-complement part
        const uint32_t n_chunks = (d.n + 31) / 32 + 4;           // + padding chunks (all N)
        for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < n_chunks; c += gridDim.x * blockDim.x) {
            const uint32_t x0 = c * 32;
            uint32_t lo, hi, bad;
            if (x0 + 32 <= L) {                                   // inside the forward copy
                const W3 w = fetch3(src, x0);
                lo = w.lo; hi = w.hi; bad = w.nv;
            } else if (x0 >= rc0 && x0 + 32 <= rc0 + L) {         // inside the reverse complement
                const W3 w = fetch3(src, L - 32 - (x0 - rc0));
                lo = __brev(~w.lo); hi = __brev(~w.hi); bad = __brev(w.nv);
            } else {                                              // boundary chunk: base by base
                lo = 0; hi = 0; bad = 0;
                for (uint32_t j = 0; j < 32; ++j) {
                    const uint32_t x = x0 + j;
                    uint32_t code = 0, isn = 1;
                    if (x < L) {
                        const uint4 q = __ldg(src + (x >> 5));
                        code = ((q.x >> (x & 31)) & 1u) | (((q.y >> (x & 31)) & 1u) << 1); isn = (q.z >> (x & 31)) & 1u;
                    } else if (x >= rc0 && x < rc0 + L) {
                        const uint32_t sp = L - 1 - (x - rc0);
                        const uint4 q = __ldg(src + (sp >> 5));
                        code = 3u - (((q.x >> (sp & 31)) & 1u) | (((q.y >> (sp & 31)) & 1u) << 1)); isn = (q.z >> (sp & 31)) & 1u;
                    }
                    if (isn) { code = 0; bad |= 1u << j; }
                    lo |= (code & 1u) << j; hi |= (code >> 1) << j;
                }
            }
            ref_rec[d.rec_off + c] = make_uint4(lo, hi, bad, 0u);  // N positions keep whatever code bits they had: the
        }                                                          // N plane decides every comparison
    }
}

// ---------------------------------------------------------------------------------------------------------------
// k5b: anchor table (parser.cpp:146-189).  The reference text holds every mal-mer of the genome twice -- at forward
// position p and, reverse-complemented, at rc0 + (len - mal - p) -- so only FORWARD positions are stored, under the
// canonical k-mer min(K, revcomp K); a lookup canonicalises the query k-mer and tries both text positions of every
// hit.  Half the inserts and half the table for the same candidate set.
// slot (32 bit) = fingerprint << pos_bits | forward position; fingerprint = top 32 - pos_bits bits of the hash
// ---------------------------------------------------------------------------------------------------------------
// One block per reference (blocks fetch references from a shared counter), no global atomics and no table clear: the
// table is built partition by partition (IDX_PS slots = 32 KB) in SHARED memory and every partition leaves with one
// coalesced store; five blocks are resident per SM, so the phases of different references overlap.
//   1. every position is hashed (a thread takes IDX_CHUNK consecutive positions from one pair of 128-bit loads) and a
//      shared-memory histogram counts the entries per partition;
//   2. after a scan of the histogram the positions are hashed AGAIN and (home slot, entry) goes straight to its place
//      in a per-block scratch list ordered by partition (the text is 10 KB and cached; parking the entries in a first
//      list instead and sorting that cost 2 x 320 KB of DRAM traffic per 40 kb reference and exposed its latency);
//   3. partition by partition: clear, insert the carried and the own entries (shared-memory CAS, all lanes busy), store.
// An entry that reaches the end of its partition is carried into the next one (rare unless the genome is a long repeat);
// the table ends with ht_tail spare slots for the carries of the last home partition, and what is still carried at the
// very end wraps around: the first partitions are then read back, topped up and stored again.  Any insertion order
// gives a table the lookup reads the same way (it takes the maximum over the whole probe chain).
constexpr uint32_t IDX_PS = 8192;
constexpr int IDX_THREADS = 256;
constexpr int IDX_CHUNK = 8;
constexpr uint32_t IDX_MAXP = 1024;                  // histogram bins; longer tables put several partitions into one bin

struct Window { uint64_t lo, hi, nv; };             // >= 40 symbols starting at a position that is a multiple of 8
__device__ __forceinline__ Window load_window(const uint4 *__restrict__ rec, uint32_t p0)
{
    const uint4 a = __ldg(rec + (p0 >> 5)), b = __ldg(rec + (p0 >> 5) + 1);
    const unsigned sh = p0 & 31;
    Window w;
    w.lo = (((uint64_t)b.x << 32) | a.x) >> sh; w.hi = (((uint64_t)b.y << 32) | a.y) >> sh; w.nv = (((uint64_t)b.z << 32) | a.z) >> sh;
    return w;
}
// hash of the mal-mer at offset j of the window; false when it holds an N
__device__ __forceinline__ bool window_hash(const Window &w, int j, int mal, uint32_t nmask, uint64_t &h)
{
    if ((uint32_t)(w.nv >> j) & nmask) return false;
    const uint64_t code = (uint64_t)((uint32_t)(w.lo >> j) & nmask) | ((uint64_t)((uint32_t)(w.hi >> j) & nmask) << 32);
    h = anchor_hash(canonical_kmer(code, mal), mal);
    return true;
}

__global__ void __launch_bounds__(IDX_THREADS, 5) build_ref_index_kernel(const RefDesc *__restrict__ refs, uint32_t n_refs, int mal,
                                                                        const uint4 *__restrict__ ref_rec, uint32_t *__restrict__ ht,
                                                                        const uint8_t *__restrict__ is_ref, uint32_t *__restrict__ carry_all,
                                                                        uint32_t carry_stride, uint2 *__restrict__ ent_all, uint32_t ent_stride,
                                                                        unsigned int *__restrict__ next_ref)
{
    __shared__ uint32_t tab[IDX_PS];
    __shared__ uint32_t s_off[IDX_MAXP + 1], s_cur[IDX_MAXP];
    __shared__ uint32_t s_carry_n[2], s_ref, s_wtot[IDX_THREADS / 32];
    const uint32_t nmask = (1u << mal) - 1;
    uint32_t *const carry0 = carry_all + (size_t)blockIdx.x * 2 * carry_stride;
    auto carry = [&](int which) { return carry0 + (which ? carry_stride : 0u); };
    uint2 *ent = ent_all + (size_t)blockIdx.x * ent_stride;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_ref = atomicAdd(next_ref, 1u);
        __syncthreads();
        const uint32_t r = s_ref;
        if (r >= n_refs) return;
        if (is_ref && !is_ref[r]) continue;
        const RefDesc d = refs[r];
        const uint4 *rec = ref_rec + d.rec_off;
        uint32_t *out = ht + d.ht_off;
        const uint32_t n_pos = d.len >= (uint32_t)mal ? d.len - mal + 1 : 0;
        const uint32_t n_chunks = (n_pos + IDX_CHUNK - 1) / IDX_CHUNK;
        const uint32_t total = d.ht_cap + d.ht_tail;
        const uint32_t n_home_parts = (d.ht_cap + IDX_PS - 1) / IDX_PS;
        const uint32_t G = (n_home_parts + IDX_MAXP - 1) / IDX_MAXP;        // partitions per histogram bin (1 unless the genome is > 2 Mb)
        const uint32_t n_bins = (n_home_parts + G - 1) / G;
        if (threadIdx.x < 2) s_carry_n[threadIdx.x] = 0;
        for (uint32_t i = threadIdx.x; i <= n_bins; i += IDX_THREADS) s_off[i] = 0;
        __syncthreads();
        // ---- 1. hash every position; count the entries of every bin (s_off[bin + 1])
        for (uint32_t c = threadIdx.x; c < n_chunks; c += IDX_THREADS) {
            const uint32_t p0 = c * IDX_CHUNK;
            const Window w = load_window(rec, p0);
#pragma unroll
            for (int j = 0; j < IDX_CHUNK; ++j) {
                uint64_t h;
                if (p0 + j < n_pos && window_hash(w, j, mal, nmask, h)) atomicAdd(&s_off[ht_slot(h, d.ht_cap) / (IDX_PS * G) + 1], 1u);
            }
        }
        __syncthreads();
        // ---- 2. exclusive scan of the bin counts (n_bins <= 1024 = 4 per thread), then the counting sort
        {
            const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
            uint32_t v[4], sum = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) { const uint32_t i = threadIdx.x * 4 + j; v[j] = i < n_bins ? s_off[i + 1] : 0u; sum += v[j]; }
            uint32_t x = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
            if (lane == 31) s_wtot[wid] = x;
            __syncthreads();
            uint32_t base = 0;
            for (int k = 0; k < wid; ++k) base += s_wtot[k];
            uint32_t run = base + x - sum;
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t i = threadIdx.x * 4 + j;
                if (i < n_bins) { s_cur[i] = run; s_off[i] = run; }
                run += v[j];
                if (i + 1 == n_bins) s_off[n_bins] = run;
            }
        }
        __syncthreads();
        // (home slot, entry) of every position, grouped by bin: hashed a second time rather than parked in a first scratch
        // list -- the text is 10 KB and cached, the list was 320 KB written to and read back from DRAM
        for (uint32_t c = threadIdx.x; c < n_chunks; c += IDX_THREADS) {
            const uint32_t p0 = c * IDX_CHUNK;
            const Window w = load_window(rec, p0);
#pragma unroll
            for (int j = 0; j < IDX_CHUNK; ++j) {
                uint64_t h;
                if (p0 + j < n_pos && window_hash(w, j, mal, nmask, h)) {
                    const uint32_t home = ht_slot(h, d.ht_cap);
                    ent[atomicAdd(&s_cur[home / (IDX_PS * G)], 1u)] =
                        make_uint2(home, ((uint32_t)(h >> 32) >> d.pos_bits << d.pos_bits) | (p0 + j));
                }
            }
        }
        __syncthreads();
        // ---- 3. the partitions
        int cur = 0;                                    // carry(cur): entries carried INTO this partition
        for (uint32_t pbase = 0; pbase < total; pbase += IDX_PS, cur ^= 1) {
            const uint32_t plen = min(IDX_PS, total - pbase);
            const uint32_t n_in = s_carry_n[cur];
            for (uint32_t i = threadIdx.x; i < plen; i += IDX_THREADS) tab[i] = HT_EMPTY;
            __syncthreads();
            if (threadIdx.x == 0) s_carry_n[cur ^ 1] = 0;
            auto insert = [&](uint32_t s, uint32_t val) {
                for (; s < plen; ++s)
                    if (atomicCAS(&tab[s], HT_EMPTY, val) == HT_EMPTY) return;
                carry(cur ^ 1)[atomicAdd(&s_carry_n[cur ^ 1], 1u)] = val;          // (at most n_pos entries exist in all)
            };
            __syncthreads();
            for (uint32_t i = threadIdx.x; i < n_in; i += IDX_THREADS) insert(0, carry(cur)[i]);
            if (pbase < d.ht_cap) {
                const uint32_t bin = (pbase / IDX_PS) / G;
                const uint32_t lo = s_off[bin], hi = s_off[bin + 1];
                for (uint32_t i = lo + threadIdx.x; i < hi; i += 2 * IDX_THREADS) {          // two loads in flight per thread
                    const uint32_t i1 = i + IDX_THREADS;
                    const uint2 e0 = ent[i];
                    uint2 e1 = make_uint2(0u, 0u);
                    if (i1 < hi) e1 = ent[i1];
                    if (G == 1 || e0.x - pbase < plen) insert(e0.x - pbase, e0.y);   // (a shared bin: only this partition's entries)
                    if (i1 < hi && (G == 1 || e1.x - pbase < plen)) insert(e1.x - pbase, e1.y);
                }
            }
            __syncthreads();
            uint4 *o4 = (uint4 *)(out + pbase);         // tables start 16-byte aligned, partition sizes are multiples of 1024
            const uint4 *t4 = (const uint4 *)tab;
            for (uint32_t i = threadIdx.x; i < plen / 4; i += IDX_THREADS) o4[i] = t4[i];
            __syncthreads();
        }
        // entries still carried at the end of the table (a genome that is mostly one repeat) wrap around to its start:
        // the finished partitions are read back, topped up and stored again until nothing is left to carry
        for (uint32_t pbase = 0; s_carry_n[cur] != 0; pbase = pbase + IDX_PS < total ? pbase + IDX_PS : 0u, cur ^= 1) {
            const uint32_t plen = min(IDX_PS, total - pbase);
            const uint32_t n_in = s_carry_n[cur];
            uint4 *o4 = (uint4 *)(out + pbase);
            uint4 *t4 = (uint4 *)tab;
            for (uint32_t i = threadIdx.x; i < plen / 4; i += IDX_THREADS) t4[i] = o4[i];
            __syncthreads();
            if (threadIdx.x == 0) s_carry_n[cur ^ 1] = 0;
            __syncthreads();
            for (uint32_t i = threadIdx.x; i < n_in; i += IDX_THREADS) {
                const uint32_t val = carry(cur)[i];
                uint32_t s = 0;
                for (; s < plen; ++s)
                    if (atomicCAS(&tab[s], HT_EMPTY, val) == HT_EMPTY) break;
                if (s == plen) carry(cur ^ 1)[atomicAdd(&s_carry_n[cur ^ 1], 1u)] = val;
            }
            __syncthreads();
            for (uint32_t i = threadIdx.x; i < plen / 4; i += IDX_THREADS) o4[i] = t4[i];
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// comparison primitives
// ---------------------------------------------------------------------------------------------------------------
// mismatch flags of Q[qp + t] vs R[rp + t], t = 0..31; anything outside either text counts as a mismatch
__device__ __forceinline__ uint32_t mm32(const Text &Q, int qp, const Text &R, int rp)
{
    int rem = min(Q.n - qp, R.n - rp);
    if (rem <= 0 || qp < 0 || rp < 0) return 0xffffffffu;
    const W3 q = fetch3(Q.rec, (uint64_t)qp), r = fetch3(R.rec, (uint64_t)rp);
    uint32_t m = (q.lo ^ r.lo) | (q.hi ^ r.hi) | q.nv | r.nv;
    if (rem < 32) m |= 0xffffffffu << rem;
    return m;
}

// same, for the 32 positions BEFORE (qp, rp), most recent first: bit t <-> Q[qp-1-t] vs R[rp-1-t]; t >= lim mismatch
__device__ __forceinline__ uint32_t mm32_back(const Text &Q, int qp, const Text &R, int rp, int lim)
{
    if (lim <= 0) return 0xffffffffu;
    int qs = qp - 32, rs = rp - 32;
    int sh = 0;
    if (qs < 0 || rs < 0) { sh = max(-qs, -rs); qs += sh; rs += sh; }      // sh < 32 because lim > 0
    const W3 q = fetch3(Q.rec, (uint64_t)qs), r = fetch3(R.rec, (uint64_t)rs);
    uint32_t m = (q.lo ^ r.lo) | (q.hi ^ r.hi) | q.nv | r.nv;
    m <<= sh;                       // ascending positions now end at bit 31 = position qp-1
    m = __brev(m);                  // bit t = position qp-1-t
    if (lim < 32) m |= 0xffffffffu << lim;
    return m;
}

// parser.cpp:192-207 (one lane): length of the exact match of Q[qp..] and R[rp..], known to be >= start
__device__ __forceinline__ int equal_len(const Text &Q, int qp, const Text &R, int rp, int start)
{
    int r = start;
    for (;;) {
        uint32_t m = mm32(Q, qp + r, R, rp + r);
        if (m) return r + __ffs(m) - 1;
        r += 32;
    }
}

// number of matching positions among the first len of Q[qp..] vs R[rp..] (one lane)
__device__ __forceinline__ int count_matches(const Text &Q, int qp, const Text &R, int rp, int len)
{
    int c = 0;
    for (int o = 0; o < len; o += 32) {
        uint32_t m = ~mm32(Q, qp + o, R, rp + o);
        int rem = len - o;
        if (rem < 32) m &= (1u << rem) - 1;
        c += __popc(m);
    }
    return c;
}

// Sliding-window rule of try_extend_forward/backward (parser.cpp:377-441) on one lane's 32 flags.
// x = (this lane's mismatch flags << 32) | the 32 flags before them.  Returns the flags of positions where the
// number of mismatches among the last aw positions exceeds am (only mismatch positions can be such positions).
__device__ __noinline__ uint32_t window_violations(uint64_t x, int aw, int am)
{
    uint32_t mm = (uint32_t)(x >> 32);
    uint32_t viol = 0;
    if (__popcll(x >> (33 - aw)) <= am) return 0;                 // not enough mismatches in reach of this lane
    const uint64_t wmask = (aw >= 64) ? ~0ULL : ((1ULL << aw) - 1);
    uint32_t rest = mm;
    while (rest) {
        int b = __ffs(rest) - 1;
        rest &= rest - 1;
        if (__popcll((x >> (33 + b - aw)) & wmask) > am) viol |= 1u << b;
    }
    return viol;
}

// The same rule for the default parameters (aw = 15, am = 7), branch-free: the 15 shifted copies of the flags are
// summed per position with a carry-save adder tree (13 adders); "more than 7" is bit 3 of the 4-bit count.
__device__ __forceinline__ uint32_t window_violations_15_7(uint64_t x)
{
    const uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
#define VB_Y(s) __funnelshift_l(lo, hi, s)                         /* bit b = flag of position b - s */
#define VB_FA(a, b, c, sum, car) { uint32_t _a = (a), _b = (b), _c = (c); sum = _a ^ _b ^ _c; car = (_a & _b) | (_c & (_a ^ _b)); }
    uint32_t s0, s1, s2, s3, s4, c0, c1, c2, c3, c4;
    VB_FA(hi, VB_Y(1), VB_Y(2), s0, c0)
    VB_FA(VB_Y(3), VB_Y(4), VB_Y(5), s1, c1)
    VB_FA(VB_Y(6), VB_Y(7), VB_Y(8), s2, c2)
    VB_FA(VB_Y(9), VB_Y(10), VB_Y(11), s3, c3)
    VB_FA(VB_Y(12), VB_Y(13), VB_Y(14), s4, c4)
    uint32_t t0, d0, t1, d1, d2;
    VB_FA(s0, s1, s2, t0, d0)
    t1 = s3 ^ s4; d1 = s3 & s4;
    d2 = t0 & t1;                                                  // (t0 ^ t1 = bit 0 of the count, unused)
    uint32_t e0, f0, e1, f1, e2, f2, f3;
    VB_FA(c0, c1, c2, e0, f0)
    VB_FA(c3, c4, d0, e1, f1)
    VB_FA(d1, d2, e0, e2, f2)
    f3 = e1 & e2;                                                  // (e1 ^ e2 = bit 1 of the count, unused)
    uint32_t g0, h0;
    VB_FA(f0, f1, f2, g0, h0)
    uint32_t h1 = g0 & f3;                                         // (g0 ^ f3 = bit 2)
    return h0 | h1;                                                // bit 3: count >= 8
#undef VB_Y
#undef VB_FA
}

// positions that END a run of >= ar consecutive matches (the run may start in the previous 32 flags)
__device__ __forceinline__ uint32_t run_ends(uint64_t x, int ar)
{
    uint64_t z = ~x, acc = z;
    if (ar == 3) acc = z & (z << 1) & (z << 2);
    else {
#pragma unroll 1
        for (int s = 1; s < ar; ++s) acc &= z << s;
    }
    return (uint32_t)(acc >> 32);
}

struct ExtResult { int len; int matches; };

__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// parser.cpp:377-409 (forward) and :412-441 (backward) for the whole warp: lane l looks at offsets [base + 32 l,
// base + 32 l + 32) counted from (qp, rp) in the direction of the scan.  lim_all = number of offsets the reference's loop
// may examine (forward: until either text ends; backward: e < max_len, qp - e > 0, rp - e > 0); offsets beyond it read as
// mismatches, which can neither end a run nor be reached before the stop.  Returns the extension length (end of the last
// run of >= ar matches before the window rule fires) and the number of matching positions inside it.
// One body for both directions keeps the kernel's instruction footprint small (the parse is instruction-fetch bound).
__device__ __noinline__ ExtResult extend(const Text Q, int qp, const Text R, int rp, int lim_all, bool backward, int aw, int am,
                                         int ar, int lane)
{
    ExtResult res = {0, 0};
    uint32_t prev_hi = 0;           // flags of the 32 positions before this super-chunk (virtual matches at start)
    int cum = 0;                    // matches in all earlier super-chunks
    auto flags_at = [&](int off) -> uint32_t {      // mismatch flags of this lane's 32 offsets starting at off
        if (backward) return mm32_back(Q, qp - off, R, rp - off, lim_all - off);
        const int lim = lim_all - off;
        if (lim <= 0) return 0xffffffffu;
        const W3 q = fetch3(Q.rec, (uint64_t)(qp + off)), r = fetch3(R.rec, (uint64_t)(rp + off));
        uint32_t f = (q.lo ^ r.lo) | (q.hi ^ r.hi) | q.nv | r.nv;
        if (lim < 32) f |= 0xffffffffu << lim;
        return f;
    };
    for (int base = 0;; base += 1024) {
        const int off = base + 32 * lane;
        if (lim_all > off + 1024) {
            // the next super-chunk is pulled into L1 while this one is evaluated (no register target, so no stall)
            const int d = backward ? -min(off + 1024 + 32, min(qp, rp)) : off + 1024;
            prefetch_l1(Q.rec + ((qp + d) >> 5)); prefetch_l1(R.rec + ((rp + d) >> 5));
        }
        const uint32_t m = flags_at(off);
        uint32_t pm = __shfl_up_sync(0xffffffffu, m, 1);
        if (lane == 0) pm = prev_hi;
        uint64_t x = ((uint64_t)m << 32) | pm;
        uint32_t viol = (aw == 15 && am == 7) ? window_violations_15_7(x) : window_violations(x, aw, am);
        uint32_t ends = run_ends(x, ar);
        const int stop_all = lim_all - base;                      // first offset of this super-chunk not examined
        unsigned vb = __ballot_sync(0xffffffffu, viol != 0);
        int limit = 1024;
        if (vb) {
            int vl = __ffs(vb) - 1;
            uint32_t v = __shfl_sync(0xffffffffu, viol, vl);
            limit = 32 * vl + __ffs(v) - 1;
        }
        bool done = vb != 0;
        if (stop_all <= limit) { limit = max(stop_all, 0); done = true; }
        int my_lim = limit - 32 * lane;                           // positions of this lane that are before the stop
        uint32_t keep = my_lim >= 32 ? 0xffffffffu : (my_lim <= 0 ? 0u : ((1u << my_lim) - 1));
        ends &= keep;
        unsigned eb = __ballot_sync(0xffffffffu, ends != 0);
        if (eb) {
            int el = 31 - __clz(eb);
            uint32_t e = __shfl_sync(0xffffffffu, ends, el);
            int last_local = 32 * el + (31 - __clz(e)) + 1;
            int upto = last_local - 32 * lane;
            uint32_t cm = upto >= 32 ? 0xffffffffu : (upto <= 0 ? 0u : ((1u << upto) - 1));
            res.len = base + last_local;
            res.matches = cum + __reduce_add_sync(0xffffffffu, __popc(~m & cm));
        }
        if (done) return res;
        cum += __reduce_add_sync(0xffffffffu, __popc(~m));
        prev_hi = __shfl_sync(0xffffffffu, m, 31);
    }
}

// parser.h:134-188
__device__ __forceinline__ double prob_len(int len) { return ldexp(1.0, -2 * len); }
__device__ __forceinline__ double ipow(double base, uint32_t e)
{
    double r = 1.0;
    while (e) {
        if (e & 1) r = __dmul_rn(r, base);
        base = __dmul_rn(base, base);
        e >>= 1;
    }
    return r;
}

// ---------------------------------------------------------------------------------------------------------------
// seed searches
// ---------------------------------------------------------------------------------------------------------------
// k-mer of `len` symbols at Q[p]; false when it holds an N or leaves the text
__device__ __forceinline__ bool kmer_at(const Text &T, int p, int len, uint64_t &code)
{
    if (p < 0 || p + len > T.n) return false;
    const W3 w = fetch3(T.rec, (uint64_t)p);
    if (w.nv & ((1u << len) - 1)) return false;
    code = kmer_code(w, len);
    return true;
}

// whole warp, parser.cpp:514-531 / :585-602: longest exact match among all reference positions of Q's mal-mer at i
// (>= mal), ties to the smallest position.  Lanes read 32 consecutive slots of the probe chain at a time.
__device__ void anchor_search(const uint32_t *__restrict__ tab, uint32_t cap, uint32_t ttotal, uint32_t pos_bits, const Text &Q, int i,
                              const Text &R, const LzParams &P, int lane, int &best_len, int &best_pos)
{
    best_len = 0; best_pos = 0;
    uint64_t code;
    if (!kmer_at(Q, i, P.mal, code)) return;
    uint64_t h = anchor_hash(canonical_kmer(code, P.mal), P.mal);
    uint32_t fp = (uint32_t)(h >> 32) >> pos_bits;
    uint32_t slot0 = ht_slot(h, cap);
    const uint32_t pmask = (1u << pos_bits) - 1;
    const int len = (R.n - 3 * P.mrd) / 2;                  // genome length; the reverse complement starts at rc0
    const int rc0 = len + 2 * P.mrd;
    int my_len = 0, my_pos = 0x7fffffff;
    for (uint32_t step = 0;; step += 32) {
        uint32_t at = slot0 + step + lane;                  // chains wrap only at the very end of the table (spare slots behind
        if (at >= ttotal) at -= ttotal;                     // the home range take the ordinary overflows)
        if (at >= ttotal) at -= ttotal;
        uint32_t s = __ldg(tab + at);
        unsigned empties = __ballot_sync(0xffffffffu, s == HT_EMPTY);
        bool in_chain = empties == 0 || lane < (__ffs(empties) - 1);
        if (in_chain && (s >> pos_bits) == fp) {
            const int pf = (int)(s & pmask);
#pragma unroll 1
            for (int strand = 0; strand < 2; ++strand) {    // forward occurrence, reverse-complement occurrence
                const int pos = strand ? rc0 + (len - P.mal - pf) : pf;
                int ml = equal_len(Q, i, R, pos, 0);
                if (ml >= P.mal && (ml > my_len || (ml == my_len && pos < my_pos))) { my_len = ml; my_pos = pos; }
            }
        }
        if (empties || step + 32 >= ttotal) break;          // an empty slot ends the chain; else the whole table was seen
    }
    int mx = __reduce_max_sync(0xffffffffu, my_len);
    if (mx == 0) return;
    int cand = (my_len == mx) ? my_pos : 0x7fffffff;
    best_len = mx;
    best_pos = __reduce_min_sync(0xffffffffu, cand);
}

// whole warp, parser.cpp:548-580: short seeds in the window [pred - lit, pred + mrd): longest continuation,
// ties to the position closest to pred, then to the smaller position
__device__ void close_search(const Text &Q, int i, const Text &R, int pred, int lit, const LzParams &P, int lane,
                             int &best_len, int &best_pos)
{
    best_len = 0; best_pos = 0;
    uint64_t qk;
    if (!kmer_at(Q, i, P.msl, qk)) return;
    const int lo = max(pred - lit, 0), hi = pred + P.mrd;
    int my_len = 0, my_dist = 0x7fffffff, my_pos = 0x7fffffff;
    for (int pos = lo + lane; pos < hi; pos += 32) {
        uint64_t rk;
        if (!kmer_at(R, pos, P.msl, rk) || rk != qk) continue;
        int ml = equal_len(Q, i, R, pos, P.msl);
        int dist = abs(pos - pred);
        if (ml > my_len || (ml == my_len && dist < my_dist)) { my_len = ml; my_dist = dist; my_pos = pos; }
    }
    int mx = __reduce_max_sync(0xffffffffu, my_len);
    if (mx == 0) return;
    int d = (my_len == mx) ? my_dist : 0x7fffffff;
    int md = __reduce_min_sync(0xffffffffu, d);
    int c = (my_len == mx && my_dist == md) ? my_pos : 0x7fffffff;
    best_len = mx;
    best_pos = __reduce_min_sync(0xffffffffu, c);
}

// whole warp, parser.cpp:251-313: best number of matches when the first b gap symbols are aligned to the left
// context and the last to_scan - b to the right context
__device__ int gap_best_matches(const Text &Q, int d, const Text &R, int r_left, int r_end_right, int len, int lane)
{
    if (len <= 0) return 0;
    int to_scan = (r_end_right < r_left) ? len : min(r_end_right - r_left, len);
    int lim = min(to_scan, r_end_right);
    int best = 0;
    for (int b = lane; b <= to_scan; b += 32) {
        int t = to_scan - b;
        int left = count_matches(Q, d, R, r_left, b);
        int right = (t <= lim) ? count_matches(Q, d + len - t, R, r_end_right - t, t) : 0;
        best = max(best, left + right);
    }
    return __reduce_max_sync(0xffffffffu, best);
}

// ---- region output (lz-ani --out-alignment; parser.cpp:786-837 calc_regions) --------------------------------------
// A region is a component seen through its match factors: query span [first match, end of last match), reference span
// [min offset of a match factor, ref_end), where ref_end follows region_t::extend_region / update_ref_end
// (defs.h:114-142): for every match factor  ref_end = max(ref_end + literals since the previous match, offset + len).
// For a run of factors on ONE diagonal this collapses to  ref_end = max(ref_end + G, end of the last match), G = the
// literals in front of that last match, so the kernel only needs, per collinear part, the first and last matching
// position and the match count -- still no factor list.
struct RegionSink {
    int32_t *rec;                      // 7 ints per region: pair, q_start, q_end, r_start, r_end, matches, mismatches
    unsigned long long *count;         // regions produced (may exceed cap: the host then re-runs with a larger buffer)
    unsigned long long cap;
};

// whole warp: among positions [0, len) of Q[qp..] vs R[rp..]: first and last matching offset (-1: none) and the count
__device__ void match_extent(const Text &Q, int qp, const Text &R, int rp, int len, int lane, int &first, int &last, int &cnt)
{
    int f = 0x7fffffff, l = -1, c = 0;
    for (int o = 32 * lane; o < len; o += 1024) {
        uint32_t m = ~mm32(Q, qp + o, R, rp + o);
        int rem = len - o;
        if (rem < 32) m &= (1u << rem) - 1;
        if (m) { f = min(f, o + __ffs(m) - 1); l = o + 31 - __clz(m); c += __popc(m); }
    }
    f = __reduce_min_sync(0xffffffffu, f);
    first = f == 0x7fffffff ? -1 : f;
    last = __reduce_max_sync(0xffffffffu, l);
    cnt = __reduce_add_sync(0xffffffffu, c);
}

// compare_ranges_both_ways (parser.cpp:251-374) as seen by a region: the chosen split (ties -> the largest, :308) and the
// match extents of the part aligned to the left context and of the part aligned to the right context
struct GapParts { int best, split, to_scan, lf, ll, ml, rv0, rf, rl, mr; };
__device__ GapParts gap_parts(const Text &Q, int d, const Text &R, int r_left, int r_end_right, int len, int lane)
{
    GapParts g = {0, 0, 0, -1, -1, 0, 0, -1, -1, 0};
    if (len <= 0) return g;
    const int to_scan = (r_end_right < r_left) ? len : min(r_end_right - r_left, len);
    const int lim = min(to_scan, r_end_right);
    int best = -1, bsplit = 0;
    for (int b = lane; b <= to_scan; b += 32) {
        int t = to_scan - b;
        int left = count_matches(Q, d, R, r_left, b);
        int right = (t <= lim) ? count_matches(Q, d + len - t, R, r_end_right - t, t) : 0;
        if (left + right >= best) { best = left + right; bsplit = b; }
    }
    const int mx = __reduce_max_sync(0xffffffffu, best);
    const int sp = __reduce_max_sync(0xffffffffu, best == mx ? bsplit : -1);
    g.best = mx; g.split = sp; g.to_scan = to_scan;
    match_extent(Q, d, R, r_left, sp, lane, g.lf, g.ll, g.ml);
    const int from_right = to_scan - sp;                 // right part: v = 0 .. from_right - 1, i = from_right - v <= lim
    g.rv0 = max(0, from_right - lim);
    match_extent(Q, d + len - from_right + g.rv0, R, r_end_right - from_right + g.rv0, from_right - g.rv0, lane, g.rf, g.rl, g.mr);
    if (g.rf >= 0) { g.rf += g.rv0; g.rl += g.rv0; }
    return g;
}

__device__ __forceinline__ void region_emit(const RegionSink &S, uint32_t pair, int qs, int qe, int rs, int re, int m, int mm, int lane)
{
    if (lane != 0) return;
    unsigned long long at = atomicAdd(S.count, 1ULL);
    if (at >= S.cap) return;
    int32_t *o = S.rec + 7 * at;
    o[0] = (int32_t)pair; o[1] = qs; o[2] = qe; o[3] = rs; o[4] = re; o[5] = m; o[6] = mm;
}

// ---------------------------------------------------------------------------------------------------------------
// k6: the parse.  All state is warp-uniform; `lane` only selects the data a lane looks at.
// ---------------------------------------------------------------------------------------------------------------
constexpr int SEED_WORDS = 6;         // seed windows of up to 192 reference positions use the Shift-And path

template <bool REGIONS>
__device__ void parse_pair(const Text &Q, const Text &R, const uint32_t *__restrict__ tab, uint32_t tcap, uint32_t ttotal, uint32_t pos_bits,
                           const LzParams &P, int lane, uint32_t (*seed_masks)[4][SEED_WORDS + 1], int &out_match, int &out_lit,
                           int &out_comp, const RegionSink &sink, uint32_t pair_idx)
{
    int reg_rs = 0, reg_re = 0;                          // REGIONS: reference span of the live component
    const int nQ = Q.n;
    int i = 0, lit = 0, pred = 0;
    bool lost = true;
    int sum_match = 0, sum_lit = 0, n_comp = 0;          // components already final (calc_stats criterion applied)
    bool comp_active = false;
    int comp_match = 0, comp_lit = 0, comp_start = 0;    // the live component; comp_start == prev_region_start
    int prev_end = 0;                                    // prev_region_end
    int last_match_end = 0, saved_lme = 0;               // end of the last match factor now / before the live component

    while (i + P.msl < nQ) {
        // ---- 1. find the next query position with a seed, 32 positions per step ------------------------------
        int steps = nQ - P.msl - i;                      // positions the main loop would still visit
        if (!lost) steps = min(steps, P.mqd - lit + 1);  // after that many literal steps the parser is lost
        steps = min(steps, 32);
        bool flag = false;
        // both k-mers of query position i + lane from one fetch; the anchor-table probe is ISSUED here and resolved after
        // the seed-window work below, so its L2/DRAM round trip overlaps ~250 instructions instead of stalling the warp
        uint64_t qk = 0;
        bool qv = false, pr_live = false;
        uint32_t pr_s = HT_EMPTY, pr_slot = 0, pr_fp = 0;
        if (lane < steps) {
            const int qp = i + lane;
            const W3 qw = fetch3(Q.rec, (uint64_t)qp);
            qv = qp + P.msl <= Q.n && (qw.nv & ((1u << P.msl) - 1)) == 0;
            qk = kmer_code(qw, P.msl);
            if (qp + P.mal <= Q.n && (qw.nv & ((1u << P.mal) - 1)) == 0) {
                const uint64_t h = anchor_hash(canonical_kmer(kmer_code(qw, P.mal), P.mal), P.mal);
                pr_fp = (uint32_t)(h >> 32) >> pos_bits;
                pr_slot = ht_slot(h, tcap);
                pr_s = __ldg(tab + pr_slot);
                pr_live = true;
            }
        }
        if (!lost) {
            // short seeds: lane t (query position i + t) may use reference positions [lo, pred + t + mrd).  The window is
            // shared by all lanes, so it is turned once into four position bit masks E_c ("symbol at window position w is
            // c", an N sets no bit); lane t then ANDs E_{q_j} >> j over the msl symbols of its k-mer (Shift-And).
            const int lo = max(pred - lit, 0);
            const int my_w = pred + lane + P.mrd - lo;
            const int max_w = pred + (steps - 1) + P.mrd - lo;
            const int nw = (max_w + P.msl - 1 + 31) >> 5;
            if (nw <= SEED_WORDS) {
                uint32_t(*E)[SEED_WORDS + 1] = seed_masks[threadIdx.x >> 5];
                if (lane < nw) {                          // lane k turns window word k into the four symbol masks
                    const int pos = lo + 32 * lane;
                    const int rem = R.n - pos;            // symbols of this word inside the text
                    uint32_t ok = 0;
                    W3 w = {0, 0, 0};
                    if (rem > 0) {
                        w = fetch3(R.rec, (uint64_t)pos);
                        ok = ~w.nv;
                        if (rem < 32) ok &= (1u << rem) - 1;
                    }
                    E[0][lane] = ~w.lo & ~w.hi & ok; E[1][lane] = w.lo & ~w.hi & ok;
                    E[2][lane] = ~w.lo & w.hi & ok;  E[3][lane] = w.lo & w.hi & ok;
                }
                if (lane < 4) E[lane][nw] = 0;
                __syncwarp();
                if (qv) {
                    uint32_t acc[SEED_WORDS];
#pragma unroll
                    for (int k = 0; k < SEED_WORDS; ++k) acc[k] = 0xffffffffu;
                    for (int j = 0; j < P.msl; ++j) {
                        const uint32_t *e = E[((qk >> j) & 1) | ((qk >> (31 + j)) & 2)];
                        uint32_t cur = e[0];
#pragma unroll
                        for (int k = 0; k < SEED_WORDS; ++k) {
                            if (k < nw) {
                                uint32_t nxt = e[k + 1];
                                acc[k] &= __funnelshift_r(cur, nxt, j);
                                cur = nxt;
                            }
                        }
                    }
#pragma unroll
                    for (int k = 0; k < SEED_WORDS; ++k) {
                        if (k < nw) {
                            int rem = my_w - 32 * k;
                            uint32_t keep = rem >= 32 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << rem) - 1));
                            if (acc[k] & keep) flag = true;
                        }
                    }
                }
                __syncwarp();
            } else {
                // very wide windows (non-default mqd/mrd): k-mers computed once per round and broadcast with shuffles
                for (int j0 = 0; j0 < max_w; j0 += 32) {
                    uint64_t rk;
                    if (!kmer_at(R, lo + j0 + lane, P.msl, rk)) rk = ~0ULL;
#pragma unroll 1
                    for (int s = 0; s < 32; ++s) {
                        uint64_t c = __shfl_sync(0xffffffffu, rk, s);
                        if (qv && c == qk && j0 + s < my_w) flag = true;
                    }
                }
            }
        }
        while (pr_live) {                                 // resolve the probe: any entry with this fingerprint in the chain?
            if (pr_s == HT_EMPTY) break;
            if ((pr_s >> pos_bits) == pr_fp) { flag = true; break; }
            pr_slot = (pr_slot + 1 == ttotal) ? 0u : pr_slot + 1;
            pr_s = __ldg(tab + pr_slot);
        }
        unsigned fb = __ballot_sync(0xffffffffu, flag);
        int adv = fb ? (__ffs(fb) - 1) : steps;
        i += adv; lit += adv; pred += adv;
        if (!fb) {
            if (!lost && lit > P.mqd) lost = true;
            continue;
        }
        // ---- 2. exact evaluation at i (parser.cpp:503-624) ----------------------------------------------------
        int best_len = 0, best_pos = 0, a_len, a_pos;
        if (!lost) close_search(Q, i, R, pred, lit, P, lane, best_len, best_pos);
        anchor_search(tab, tcap, ttotal, pos_bits, Q, i, R, P, lane, a_len, a_pos);
        if (lost) { best_len = a_len; best_pos = a_pos; }
        else if (a_pos) {                                 // positions double as booleans in the reference (:604-606)
            if (!best_pos) { best_pos = a_pos; best_len = a_len; }
            else {
                double anchor_prob = ipow(1.0 - prob_len(a_len), (uint32_t)(2 * (R.n + 1 - a_len)));
                double close_prob = ipow(1.0 - prob_len(best_len), (uint32_t)(lit + P.mrd + 1 - best_len));
                if (anchor_prob > close_prob) { best_pos = a_pos; best_len = a_len; }
            }
        }
        if (best_len < P.msl) {                           // fingerprint collision or the position-0 quirk
            ++i; ++lit; ++pred;
            if (!lost && lit > P.mqd) lost = true;
            continue;
        }
        // ---- 3. account for the match (parser.cpp:626-698) ----------------------------------------------------
        int reg_buf = 0;                                  // REGIONS: literals pending in front of the anchor factor
        bool reg_fresh = false;
        if (!lost && abs(best_pos - pred) <= P.mrd) {
            int g;
            if (REGIONS) {
                const int r_left = pred - lit, r_end_right = best_pos + best_len;
                const GapParts gp = gap_parts(Q, i - lit, R, r_left, r_end_right, lit, lane);
                g = gp.best;
                if (gp.ml > 0) {
                    reg_rs = min(reg_rs, r_left + gp.lf);
                    reg_re = max(reg_re + (gp.ll + 1 - gp.ml), r_left + gp.ll + 1);
                    reg_buf = gp.split - 1 - gp.ll;
                } else reg_buf = gp.split;
                reg_buf += lit - gp.to_scan;
                const int from_right = gp.to_scan - gp.split;
                if (gp.mr > 0) {
                    const int rbase = r_end_right - from_right;
                    reg_rs = min(reg_rs, rbase + gp.rf);
                    reg_re = max(reg_re + reg_buf + (gp.rl + 1 - gp.mr), rbase + gp.rl + 1);
                    reg_buf = from_right - 1 - gp.rl;
                } else reg_buf += from_right;
            } else
                g = gap_best_matches(Q, i - lit, R, pred - lit, best_pos + best_len, lit, lane);
            comp_match += g + best_len;
            comp_lit += lit - g;
        } else {
            if (comp_active) {
                if (prev_end - comp_start < P.reg) last_match_end = saved_lme;      // region deleted (:643-657)
                else if (comp_match + comp_lit >= P.reg) {
                    sum_match += comp_match; sum_lit += comp_lit; ++n_comp;
                    if (REGIONS) region_emit(sink, pair_idx, comp_start, comp_start + comp_match + comp_lit, reg_rs, reg_re, comp_match, comp_lit, lane);
                }
            }
            reg_fresh = true;
            saved_lme = last_match_end;
            int tail = i - last_match_end;                // length of the literal run in front of the anchor
            ExtResult back = {0, 0};
            if (tail > 0) back = extend(Q, i, R, best_pos, min(tail, min(i, best_pos)), true, P.aw, P.am, P.ar, lane);
            comp_active = true;
            comp_start = i - back.len;
            comp_match = back.matches + best_len;
            comp_lit = back.len - back.matches;
        }
        i += best_len;
        pred = best_pos + best_len;
        lit = 0;
        lost = false;
        ExtResult fw = extend(Q, i, R, pred, min(Q.n - i, R.n - pred), false, P.aw, P.am, P.ar, lane);
        comp_match += fw.matches;
        comp_lit += fw.len - fw.matches;
        if (REGIONS) {
            if (reg_fresh) { reg_rs = best_pos - (i - best_len - comp_start); reg_re = pred + fw.len; }    // i - best_len - comp_start = backward extension
            else { reg_rs = min(reg_rs, best_pos); reg_re = max(reg_re + reg_buf + (fw.len - fw.matches), pred + fw.len); }
        }
        i += fw.len;
        pred += fw.len;
        prev_end = i;
        last_match_end = i;
    }
    // ---- tail (parser.cpp:710-713) ------------------------------------------------------------------------------
    if (!lost) {
        const int T = lit + (nQ - i);
        const int qs = i - lit, rs = pred - lit - P.msl;
        int mt = 0, last = -1, first = 0x7fffffff;
        for (int o = 32 * lane; o < T; o += 1024) {
            uint32_t m = ~mm32(Q, qs + o, R, rs + o);
            int rem = T - o;
            if (rem < 32) m &= (1u << rem) - 1;
            mt += __popc(m);
            if (m) { last = o + 31 - __clz(m); if (REGIONS) first = min(first, o + __ffs(m) - 1); }
        }
        mt = __reduce_add_sync(0xffffffffu, mt);
        last = __reduce_max_sync(0xffffffffu, last);
        if (mt > 0) {
            comp_match += mt; comp_lit += last + 1 - mt;
            if (REGIONS) {
                first = __reduce_min_sync(0xffffffffu, first);
                reg_rs = min(reg_rs, rs + first);
                reg_re = max(reg_re + (last + 1 - mt), rs + last + 1);
            }
        }
    }
    if (comp_active && comp_match + comp_lit >= P.reg) {
        sum_match += comp_match; sum_lit += comp_lit; ++n_comp;
        if (REGIONS) region_emit(sink, pair_idx, comp_start, comp_start + comp_match + comp_lit, reg_rs, reg_re, comp_match, comp_lit, lane);
    }
    out_match = sum_match; out_lit = sum_lit; out_comp = n_comp;
}

template <int MINB, bool REGIONS>
__global__ void __launch_bounds__(128, MINB) parse_kernel(const uint4 *__restrict__ grec,
                                                    const uint64_t *__restrict__ gofs, const uint32_t *__restrict__ glen,
                                                    const RefDesc *__restrict__ refs, const uint4 *__restrict__ ref_rec,
                                                    const uint32_t *__restrict__ ht,
                                                    const uint32_t *__restrict__ pair_ref, const uint32_t *__restrict__ pair_qry,
                                                    uint32_t n_pairs, LzParams P, unsigned int *__restrict__ cursor,
                                                    int32_t *__restrict__ stats, RegionSink sink, uint32_t pair_base)
{
    const int lane = threadIdx.x & 31;
    __shared__ uint32_t seed_masks[4][4][SEED_WORDS + 1];       // per warp: 4 symbol masks of the seed window
    for (;;) {
        uint32_t idx = 0;
        if (lane == 0) idx = atomicAdd(cursor, 1u);
        idx = __shfl_sync(0xffffffffu, idx, 0);
        if (idx >= n_pairs) return;
        const RefDesc d = refs[pair_ref[idx]];
        const uint32_t q = pair_qry[idx];
        Text R = {ref_rec + d.rec_off, (int)d.n};
        uint64_t qo = gofs[q];
        Text Q = {grec + (qo >> 5), (int)glen[q] + P.mrd};
        int m, l, c;
        parse_pair<REGIONS>(Q, R, ht + d.ht_off, d.ht_cap, d.ht_cap + d.ht_tail, d.pos_bits, P, lane, seed_masks, m, l, c, sink, pair_base + idx);
        if (lane == 0) { stats[3 * (uint64_t)idx] = m; stats[3 * (uint64_t)idx + 1] = l; stats[3 * (uint64_t)idx + 2] = c; }
    }
}


// ---------------------------------------------------------------------------------------------------------------
// Device-side pair list (the host glue of lz-ani around its matching loop -- filter symmetrisation filter.cpp:80-81,
// 253-289, re-numbering seq_reservoir.cpp:215-251, per-reference work list lz_matcher.cpp:190-270 -- done on the GPU
// from the prefilter's device-resident candidate list: no D2H / host sort / H2D between the two stages)
// ---------------------------------------------------------------------------------------------------------------
struct SchedView {
    const uint64_t *keys;          // sorted directed pairs: ref LZ id << gbits | query LZ id
    int gbits;
    const uint32_t *order;         // LZ id -> input id
    const int32_t *slot_of_gid;    // input id -> reference descriptor (-1: not a reference of this rank)
    const uint32_t *heavy;         // the expensive pairs, to be parsed first
    const uint8_t *heavy_flag;
    const unsigned long long *counts;   // [0] directed pairs, [1] heavy pairs
};

// relative cost of a parse (scheduling only): the query is scanned once (~0.25 instructions per base) and every seed
// event costs ~1000 instructions; events happen where the window rule (more than 7 mismatches in 15) fires, at a rate
// of about C(15,8) p^8 (1-p)^7 per base at divergence p
__device__ __forceinline__ float parse_cost(float ani, uint32_t qlen)
{
    const float pp = fminf(fmaxf(1.0f - ani, 0.0f), 0.5f), q = 1.0f - pp;
    const float p2 = pp * pp, p4 = p2 * p2, q2 = q * q, q4 = q2 * q2;
    return (0.25f + 1000.0f * 6435.0f * (p4 * p4) * (q4 * q2 * q)) * (float)qlen;
}

// candidate pairs (row << 32 | col, input ids) -> the directed pairs this rank parses (reference = a genome it owns)
__global__ void __launch_bounds__(256) expand_pairs_kernel(const uint64_t *__restrict__ pairs, const float *__restrict__ ani, uint64_t n_pairs,
                                                           const uint32_t *__restrict__ rank, const uint32_t *__restrict__ glen, int gbits,
                                                           uint32_t world, uint32_t me, const int32_t *__restrict__ slot_of_gid,
                                                           uint64_t *__restrict__ out_keys, uint32_t *__restrict__ out_cost,
                                                           unsigned long long *__restrict__ counts, uint8_t *__restrict__ is_ref)
{
    const int lane = threadIdx.x & 31;
    for (uint64_t i0 = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) & ~31ULL; i0 < n_pairs; i0 += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = i0 + lane;
        uint32_t r = 0, c = 0;
        bool a = false, b = false;
        float an = 0.9f;
        if (i < n_pairs) {
            const uint64_t k = pairs[i];
            r = (uint32_t)(k >> 32); c = (uint32_t)k;
            a = r % world == me; b = c % world == me;
            if (ani) an = ani[i];
        }
        uint64_t at_a = 2 * i, at_b = 2 * i + 1;
        if (world > 1) {                                      // compact append (order is restored by the sort that follows)
            const unsigned ma = __ballot_sync(0xffffffffu, a), mb = __ballot_sync(0xffffffffu, b);
            unsigned long long base = 0;
            if (lane == 0 && (ma | mb)) base = atomicAdd(&counts[0], (unsigned long long)(__popc(ma) + __popc(mb)));
            base = __shfl_sync(0xffffffffu, base, 0);
            at_a = base + __popc(ma & ((1u << lane) - 1));
            at_b = base + __popc(ma) + __popc(mb & ((1u << lane) - 1));
        }
        if (a) {
            out_keys[at_a] = ((uint64_t)rank[r] << gbits) | rank[c];
            out_cost[at_a] = __float_as_uint(parse_cost(an, glen[c]));
            is_ref[slot_of_gid[r]] = 1;
        }
        if (b) {
            out_keys[at_b] = ((uint64_t)rank[c] << gbits) | rank[r];
            out_cost[at_b] = __float_as_uint(parse_cost(an, glen[r]));
            is_ref[slot_of_gid[c]] = 1;
        }
    }
    if (world == 1 && blockIdx.x == 0 && threadIdx.x == 0) counts[0] = 2 * n_pairs;
}

// all-vs-all: every ordered pair (r, q), r != q, in LZ ids
__global__ void __launch_bounds__(256) all_pairs_kernel(uint32_t n, int gbits, uint64_t *__restrict__ out_keys, uint32_t *__restrict__ out_cost,
                                                        unsigned long long *__restrict__ counts)
{
    const uint64_t total = (uint64_t)n * (n - 1);
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t r = (uint32_t)(i / (n - 1)), t = (uint32_t)(i % (n - 1));
        out_keys[i] = ((uint64_t)r << gbits) | (t < r ? t : t + 1);
        out_cost[i] = 0;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) counts[0] = total;
}

// Longest-processing-time-first for the expensive tail: a parse costs 10-100x more for a divergent pair than for a
// near-identical one and warps take pairs from a shared cursor, so the (estimated) most expensive 1/8 of the pairs
// goes first; the rest keeps the by-reference order (L2 locality of the anchor tables).
constexpr int COST_CLASSES = 4096;          // top 13 bits of the positive float: monotone in the cost
__device__ __forceinline__ int cost_class(uint32_t bits) { return (int)(bits >> 19) & (COST_CLASSES - 1); }

__global__ void __launch_bounds__(256) cost_hist_kernel(const uint32_t *__restrict__ cost, const unsigned long long *__restrict__ counts,
                                                        uint32_t *__restrict__ hist)
{
    const uint64_t n = counts[0];
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        atomicAdd(&hist[cost_class(cost[i])], 1u);
}

// one block: the smallest suffix of classes holding >= n / heavy_div pairs becomes "heavy"; hist[c] := start of class c
// in the heavy list (classes in descending order); counts[1] = number of heavy pairs, counts[2] = the cut
__global__ void __launch_bounds__(1024) lpt_cut_kernel(uint32_t *__restrict__ hist, unsigned long long *__restrict__ counts, int heavy_div)
{
    // suffix sums over the 4 096 classes (4 per thread, descending class order), then the cut is found in parallel
    __shared__ uint32_t warp_tot[32];
    __shared__ int s_cut;
    const int t = threadIdx.x;
    uint32_t v[4], sum = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) { v[j] = hist[COST_CLASSES - 1 - (4 * t + j)]; sum += v[j]; }     // v[j]: class COST_CLASSES-1-(4t+j)
    const int lane = t & 31, w = t >> 5;
    uint32_t x = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) warp_tot[w] = x;
    if (t == 0) s_cut = COST_CLASSES;
    __syncthreads();
    if (w == 0) {
        uint32_t tt = warp_tot[lane], z = tt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, z, o); if (lane >= o) z += y; }
        warp_tot[lane] = z - tt;
    }
    __syncthreads();
    uint32_t before = warp_tot[w] + x - sum;          // pairs in classes ABOVE class COST_CLASSES-1-4t
    const unsigned long long n = counts[0];
    const unsigned long long want = n / heavy_div;
    // the cut: the highest class c such that the pairs in classes >= c reach `want` (none when n < 64)
    uint32_t run = before;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = COST_CLASSES - 1 - (4 * t + j);
        if (n >= 64 && run < want && run + v[j] >= want) s_cut = c;      // exactly one (thread, j) satisfies this when want > 0
        run += v[j];
    }
    __syncthreads();
    int cut = s_cut;
    if (n >= 64 && want == 0) cut = COST_CLASSES;
    // one huge class: do not reorder half the list
    __shared__ uint32_t s_acc;
    if (t == 0) s_acc = 0;
    __syncthreads();
    run = before;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = COST_CLASSES - 1 - (4 * t + j);
        if (c == cut) s_acc = run + v[j];
        run += v[j];
    }
    __syncthreads();
    if (cut < COST_CLASSES && (unsigned long long)s_acc > n / 2 && heavy_div > 2) ++cut;
    // starts of the heavy classes in the heavy list (descending class order): pairs in the classes above
    run = before;
    unsigned long long heavy_total = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = COST_CLASSES - 1 - (4 * t + j);
        if (c >= cut) { hist[c] = run; heavy_total = run + v[j]; }
        run += v[j];
    }
    // the number of heavy pairs = pairs in classes >= cut: written by the thread that owns class `cut` (or 0)
    if (cut >= COST_CLASSES) { if (t == 0) { counts[1] = 0; counts[2] = (unsigned long long)cut; } }
    else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (COST_CLASSES - 1 - (4 * t + j) == cut) { counts[1] = heavy_total; counts[2] = (unsigned long long)cut; }
    }
}

__global__ void __launch_bounds__(256) heavy_fill_kernel(const uint32_t *__restrict__ cost, const unsigned long long *__restrict__ counts,
                                                         uint32_t *__restrict__ class_cur, uint32_t *__restrict__ heavy, uint8_t *__restrict__ heavy_flag)
{
    const uint64_t n = counts[0];
    const int cut = (int)counts[2];
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const int c = cost_class(cost[i]);
        const bool hv = c >= cut;
        heavy_flag[i] = hv ? 1 : 0;
        if (hv) heavy[atomicAdd(&class_cur[c], 1u)] = (uint32_t)i;
    }
}

// the parse over a device-built schedule: work item w < n_heavy -> heavy[w], else the pair w - n_heavy unless it is heavy
template <int MINB>
__global__ void __launch_bounds__(128, MINB) parse_sched_kernel(const uint4 *__restrict__ grec, const uint64_t *__restrict__ gofs,
                                                                const uint32_t *__restrict__ glen, const RefDesc *__restrict__ refs,
                                                                const uint4 *__restrict__ ref_rec, const uint32_t *__restrict__ ht,
                                                                SchedView V, LzParams P, unsigned long long *__restrict__ cursor,
                                                                int32_t *__restrict__ stats)
{
    const int lane = threadIdx.x & 31;
    __shared__ uint32_t seed_masks[4][4][SEED_WORDS + 1];
    const unsigned long long n_dir = V.counts[0], n_heavy = V.counts[1];
    const uint64_t qmask = (1ULL << V.gbits) - 1;
    const RegionSink no_sink = {nullptr, nullptr, 0};
    for (;;) {
        unsigned long long w = 0;
        if (lane == 0) w = atomicAdd(cursor, 1ULL);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= n_heavy + n_dir) return;
        uint64_t idx;
        if (w < n_heavy) idx = V.heavy[w];
        else { idx = w - n_heavy; if (V.heavy_flag[idx]) continue; }
        const uint64_t key = V.keys[idx];
        const uint32_t rg = V.order[(uint32_t)(key >> V.gbits)], q = V.order[(uint32_t)(key & qmask)];
        const RefDesc d = refs[V.slot_of_gid[rg]];
        Text R = {ref_rec + d.rec_off, (int)d.n};
        Text Q = {grec + (gofs[q] >> 5), (int)glen[q] + P.mrd};
        int m, l, c;
        parse_pair<false>(Q, R, ht + d.ht_off, d.ht_cap, d.ht_cap + d.ht_tail, d.pos_bits, P, lane, seed_masks, m, l, c, no_sink, 0);
        if (lane == 0) { stats[3 * idx] = m; stats[3 * idx + 1] = l; stats[3 * idx + 2] = c; }
    }
}

__global__ void fill_sentinel_kernel(uint64_t *p, uint64_t lo, uint64_t hi)
{
    for (uint64_t i = lo + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < hi; i += (uint64_t)gridDim.x * blockDim.x) p[i] = ~0ULL;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------------------------------------------
// The reference side (texts + anchor tables) depends only on WHICH genomes are references, so it is launched first and
// the host builds and sorts the pair list while the GPU works on it (vb_align_job_begin ... vb_align_job_run).
struct RefBatch {
    std::vector<RefDesc> refs;
    uint64_t recs = 0, slots = 0, bytes = 0, max_len = 0;
    DevBuf<RefDesc> d_refs;
    DevBuf<uint4> ref_rec;
    DevBuf<uint32_t> ht, carry;
    DevBuf<uint2> ent;
    DevBuf<unsigned int> next_ref;
};

static uint64_t ref_table_slots(uint64_t len)
{
    // forward positions only (canonical k-mers); a multiple of 1024 keeps every table 16-byte aligned
    static const int quarter_slots = getenv("VB_ALIGN_TABLE_Q") ? atoi(getenv("VB_ALIGN_TABLE_Q")) : 16;  // slots per position * 4 (load 0.25 measured best)
    return std::max<uint64_t>(1024, ((uint64_t)quarter_slots * (len + 1) / 4 + 1023) / 1024 * 1024);
}

// spare slots behind the home range: probe chains run into them instead of wrapping; only a chain that is still not
// finished at the very end of the table (a genome that is mostly one repeat) continues at slot 0
static uint64_t ref_table_tail(uint64_t len) { (void)len; return 1024; }

static uint64_t ref_bytes(uint64_t len, int mrd)
{
    const uint64_t chunks = (2 * len + 3 * (uint64_t)mrd + 31) / 32 + 4;
    return chunks * 16 + (ref_table_slots(len) + ref_table_tail(len)) * 4;
}

static void ref_batch_add(RefBatch &b, const vb_genomes *g, uint32_t gid, int mrd)
{
    const uint64_t len = g->length(gid);
    const uint64_t nR = 2 * len + 3 * (uint64_t)mrd;
    const uint64_t chunks = (nR + 31) / 32 + 4;
    const uint64_t cap = ref_table_slots(len), tail = ref_table_tail(len);
    if (cap + tail >= (1ULL << 32)) throw vb_error(VB_ERR_ARG, "genome too long for the anchor table");
    uint32_t pos_bits = 1;
    while ((1ULL << pos_bits) <= nR) ++pos_bits;
    RefDesc d;
    d.rec_off = b.recs; d.ht_off = b.slots;
    d.ht_cap = (uint32_t)cap; d.ht_tail = (uint32_t)tail; d.pos_bits = pos_bits; d.n = (uint32_t)nR; d.len = (uint32_t)len; d.gid = gid;
    b.refs.push_back(d);
    b.recs += chunks + 2; b.slots += cap + tail; b.bytes += chunks * 16 + (cap + tail) * 4;
    b.max_len = std::max<uint64_t>(b.max_len, len);
}

// allocate, upload the descriptors, build texts and anchor tables (all asynchronous).  is_ref (device, optional): one
// flag per descriptor; references whose flag is 0 are skipped.
static void ref_batch_launch(vb_ctx *ctx, RefBatch &b, const DevGenomes &dg, const vb_align_params *ap, cudaStream_t st,
                             const uint8_t *d_is_ref = nullptr)
{
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, ctx->device);
    const uint32_t n_refs = (uint32_t)b.refs.size();
    const uint32_t iblocks = std::min<uint32_t>(n_refs, (uint32_t)n_sm * 5);
    const uint32_t carry_stride = (uint32_t)((b.max_len + 1 + 1023) / 1024 * 1024);
    if (!b.d_refs.p) {
        b.d_refs.alloc(n_refs);
        b.ref_rec.alloc(b.recs + 8);
        b.ht.alloc(b.slots);
        VB_CUDA(cudaMemcpyAsync(b.d_refs.p, b.refs.data(), sizeof(RefDesc) * n_refs, cudaMemcpyHostToDevice, st));
    }
    b.carry.alloc((size_t)iblocks * 2 * carry_stride);
    const uint32_t ent_stride = carry_stride + 2 * IDX_CHUNK;
    b.ent.alloc((size_t)iblocks * ent_stride);
    b.next_ref.alloc(1);
    VB_CUDA(cudaMemsetAsync(b.next_ref.p, 0, sizeof(unsigned int), st));
    dim3 grid_b(16, (unsigned)std::min<size_t>(n_refs, 32768));
    build_ref_text_kernel<<<grid_b, 256, 0, st>>>(dg.rec.p, dg.gofs.p, b.d_refs.p, n_refs, ap->mrd, b.ref_rec.p, d_is_ref);
    VB_LAUNCH_CHECK(ctx);
    build_ref_index_kernel<<<iblocks, IDX_THREADS, 0, st>>>(b.d_refs.p, n_refs, ap->mal, b.ref_rec.p, b.ht.p, d_is_ref, b.carry.p, carry_stride,
                                                            b.ent.p, ent_stride, b.next_ref.p);
    VB_LAUNCH_CHECK(ctx);
}

static void parse_launch(vb_ctx *ctx, const DevGenomes &dg, const RefBatch &b, const uint32_t *d_pref, const uint32_t *d_pqry,
                         uint32_t nb, const LzParams &P, unsigned int *d_cursor, int32_t *d_stats, cudaStream_t st,
                         const RegionSink *sink = nullptr, uint32_t pair_base = 0)
{
    int per_sm = 0;
    static const int minb = getenv("VB_PARSE_MINB") ? atoi(getenv("VB_PARSE_MINB")) : 7;
    auto kern = minb >= 8 ? parse_kernel<8, false> : (minb == 7 ? parse_kernel<7, false> : (minb == 6 ? parse_kernel<6, false> : parse_kernel<5, false>));
    if (sink) kern = parse_kernel<5, true>;
    RegionSink rs = sink ? *sink : RegionSink{nullptr, nullptr, 0};
    VB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, 0));
    int n_sm = 0;
    VB_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, ctx->device));
    int blocks = std::max(1, std::min<int>(per_sm * n_sm, (int)((nb + 3) / 4)));
    kern<<<blocks, 128, 0, st>>>(dg.rec.p, dg.gofs.p, dg.glen.p, b.d_refs.p, b.ref_rec.p, b.ht.p, d_pref,
                                 d_pqry, nb, P, d_cursor, d_stats, rs, pair_base);
    VB_LAUNCH_CHECK(ctx);
}

// device memory one batch of reference texts + anchor tables may take (no per-call memory query); VB_ALIGN_BUDGET_MB
// (test hook) forces small batches so that the multi-batch path runs on small inputs
static uint64_t ref_budget(const vb_ctx *ctx)
{
    const char *e = getenv("VB_ALIGN_BUDGET_MB");
    if (e && atoll(e) > 0) return (uint64_t)atoll(e) << 20;
    return (uint64_t)(ctx->mem_total * 0.4);
}

struct vb_align_job {
    vb_ctx *ctx;
    const vb_genomes *g;
    vb_align_params ap;
    cudaStream_t st;
    std::chrono::steady_clock::time_point h0;
    EventTimer t_all, t_up, t_idx;
    const DevGenomes *dg = nullptr;
    bool prebuilt = false;               // all references of the call fit one batch and are being indexed already
    std::vector<int32_t> slot_of_gid;    // prebuilt: index of a genome's RefDesc, -1 if it is not a reference
    RefBatch all;
    vb_align_job(vb_ctx *c, const vb_genomes *gg, const vb_align_params *p)
        : ctx(c), g(gg), ap(*p), st((cudaStream_t)c->stream), t_all(st), t_up(st), t_idx(st) {}
};

vb_align_job *vb_align_job_begin(vb_ctx *ctx, const vb_genomes *g, const vb_align_params *ap, const uint8_t *is_ref)
{
    if (ap->mal < 4 || ap->mal > 31 || ap->msl < 2 || ap->msl > ap->mal || ap->msl > 31)
        throw vb_error(VB_ERR_ARG, "need 2 <= msl <= mal <= 31");
    if (ap->aw < 1 || ap->aw > 32 || ap->ar < 1 || ap->ar > 32 || ap->am < 0)
        throw vb_error(VB_ERR_ARG, "need 1 <= aw <= 32, 1 <= ar <= 32, am >= 0");
    if (ap->mrd < 1 || ap->mrd > 4096 || ap->mqd < 0) throw vb_error(VB_ERR_ARG, "need 1 <= mrd <= 4096, mqd >= 0");
    VB_CUDA(cudaSetDevice(ctx->device));
    vb_align_job *job = new vb_align_job(ctx, g, ap);
    try {
        job->h0 = std::chrono::steady_clock::now();
        job->t_all.start();
        job->t_up.start();
        job->dg = &vb_get_dev_genomes(ctx, g, (uint32_t)ap->mrd + 128);
        job->t_up.stop();
        // reference texts + anchor tables of one batch may take up to 40 % of the device (no per-call memory query)
        const uint64_t budget = ref_budget(ctx);
        const uint32_t ng = g->count();
        uint64_t need = 0;
        uint32_t n_refs = 0;
        for (uint32_t i = 0; i < ng; ++i) if (is_ref[i]) { need += ref_bytes(g->length(i), ap->mrd); ++n_refs; }
        job->t_idx.start();
        if (n_refs && need <= budget) {
            job->slot_of_gid.assign(ng, -1);
            for (uint32_t i = 0; i < ng; ++i) if (is_ref[i]) { job->slot_of_gid[i] = (int32_t)job->all.refs.size(); ref_batch_add(job->all, g, i, ap->mrd); }
            ref_batch_launch(ctx, job->all, *job->dg, ap, job->st);
            job->prebuilt = true;
        }
        job->t_idx.stop();
    } catch (...) {
        delete job;
        throw;
    }
    return job;
}

void vb_align_job_end(vb_align_job *job) { delete job; }

void vb_align_job_run(vb_align_job *job, const uint32_t *ref, const uint32_t *qry, uint64_t n, int32_t *stats,
                      std::vector<int32_t> *regions, const float *cost)
{
    vb_ctx *ctx = job->ctx;
    const vb_genomes *g = job->g;
    const vb_align_params *ap = &job->ap;
    cudaStream_t st = job->st;
    const DevGenomes &dg = *job->dg;
    if (n >= (1ULL << 32)) throw vb_error(VB_ERR_ARG, "more than 2^32 pairs in one call");
    const uint32_t ng = g->count();
    for (uint64_t i = 0; i < n; ++i)
        if (ref[i] >= ng || qry[i] >= ng) throw vb_error(VB_ERR_ARG, "pair id out of range");
    LzParams P = {ap->mal, ap->msl, ap->mrd, ap->mqd, ap->reg, ap->aw, ap->am, ap->ar};
    double host_prep_ms = 0, host_post_ms = 0;
    double ms_index = 0, ms_parse = 0;

    // pairs grouped by reference (stable), so that one reference's index is built once and stays hot in L2
    std::vector<uint32_t> order(n);
    for (uint64_t i = 0; i < n; ++i) order[i] = (uint32_t)i;
    {   // already grouped (every reference in one contiguous run, as vb_align delivers it)?  then keep the order
        std::vector<uint8_t> seen(ng, 0);
        bool grouped = true;
        for (uint64_t i = 0; i < n && grouped; ++i)
            if (i == 0 || ref[i] != ref[i - 1]) { if (seen[ref[i]]) grouped = false; seen[ref[i]] = 1; }
        if (!grouped) std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return ref[a] < ref[b]; });
    }
    // Longest-processing-time-first for the expensive tail: a parse costs 10-100x more for a divergent pair than for a
    // near-identical one, and warps take pairs from a shared cursor in list order, so the (estimated) most expensive
    // 1/8 of the pairs goes to the front (fewer than one wave of warps, so their mutual order does not matter); the rest
    // keeps the by-reference grouping (L2 locality of the anchor tables).  Only when all references are indexed in one
    // batch (any order is valid then).
    static const bool lpt_off = getenv("VB_ALIGN_NO_LPT") != nullptr;
    if (cost && job->prebuilt && !lpt_off && n >= 64) {
        // counting sort on the top 13 bits of the (positive) float: 4096 cost classes, monotone in the cost -- O(n), and
        // a coarse order is all the scheduling needs
        constexpr int NB = 4096;
        auto cls = [&](uint32_t i) { uint32_t b; memcpy(&b, &cost[i], 4); return (int)(b >> 19) & (NB - 1); };
        std::vector<uint32_t> hist(NB + 1, 0);
        for (uint64_t i = 0; i < n; ++i) hist[cls((uint32_t)i)]++;
        static const int heavy_div = getenv("VB_ALIGN_HEAVY_DIV") ? std::max(1, atoi(getenv("VB_ALIGN_HEAVY_DIV"))) : 8;
        const uint64_t n_heavy = n / heavy_div;
        int cut = NB;                                    // classes >= cut are "heavy": the smallest suffix with >= n/8 pairs
        uint64_t acc = 0;
        while (cut > 0 && acc < n_heavy) acc += hist[--cut];
        if (acc > n / 2 && heavy_div > 2) { acc -= hist[cut]; ++cut; }   // one huge class: do not reorder half the list
        std::vector<uint32_t> start(NB + 1, 0);          // heavy classes in descending order
        uint64_t run = 0;
        for (int c = NB - 1; c >= cut; --c) { start[c] = (uint32_t)run; run += hist[c]; }
        std::vector<uint32_t> merged(n);
        uint64_t light_at = run;
        for (uint32_t i : order) {
            const int c = cls(i);
            if (c >= cut) merged[start[c]++] = i;
            else merged[light_at++] = i;
        }
        order.swap(merged);
    }
    const uint64_t budget = ref_budget(ctx);
    ctx->set_timing("align.hp4_sched_ms", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - job->h0).count());

    DevBuf<int32_t> d_stats(3 * std::max<uint64_t>(n, 1));
    DevBuf<uint32_t> d_pref(std::max<uint64_t>(n, 1)), d_pqry(std::max<uint64_t>(n, 1));
    DevBuf<unsigned int> d_cursor(1);

    uint64_t pos = 0;
    int n_batches = 0;
    while (pos < n) {
        // ---- the references of this batch: all of them (already launched), or as many as fit the memory budget
        RefBatch local;
        RefBatch &rb = job->prebuilt ? job->all : local;
        std::vector<uint32_t> b_ref, b_qry;
        uint64_t end = pos;
        if (job->prebuilt) {
            b_ref.resize(n); b_qry.resize(n);
            for (uint64_t i = 0; i < n; ++i) {
                const int32_t slot = job->slot_of_gid[ref[order[i]]];
                if (slot < 0) throw vb_error(VB_ERR_INTERNAL, "reference was not announced to vb_align_job_begin");
                b_ref[i] = (uint32_t)slot; b_qry[i] = qry[order[i]];
            }
            end = n;
        } else {
            while (end < n) {
                uint32_t r = ref[order[end]];
                if (local.refs.empty() || local.refs.back().gid != r) {
                    const uint64_t need = ref_bytes(g->length(r), ap->mrd);
                    if (!local.refs.empty() && local.bytes + need > budget) break;
                    if (local.refs.empty() && need > budget) throw vb_error(VB_ERR_MEM, "reference index does not fit device memory");
                    ref_batch_add(local, g, r, ap->mrd);
                }
                b_ref.push_back((uint32_t)local.refs.size() - 1);
                b_qry.push_back(qry[order[end]]);
                ++end;
            }
        }
        const uint32_t nb = (uint32_t)(end - pos);
        EventTimer t_idx(st), t_par(st);
        if (!job->prebuilt) {
            t_idx.start();
            ref_batch_launch(ctx, local, dg, ap, st);
            t_idx.stop();
        }
        VB_CUDA(cudaMemcpyAsync(d_pref.p, b_ref.data(), sizeof(uint32_t) * nb, cudaMemcpyHostToDevice, st));
        VB_CUDA(cudaMemcpyAsync(d_pqry.p, b_qry.data(), sizeof(uint32_t) * nb, cudaMemcpyHostToDevice, st));
        VB_CUDA(cudaMemsetAsync(d_cursor.p, 0, sizeof(unsigned int), st));
        if (n_batches == 0) host_prep_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - job->h0).count();
        t_par.start();
        if (!regions)
            parse_launch(ctx, dg, rb, d_pref.p, d_pqry.p, nb, P, d_cursor.p, d_stats.p, st);
        else {
            // regions are appended to a bounded buffer; the kernel always counts them, so one re-run with the exact
            // size suffices when the first guess was too small (the parse is deterministic)
            unsigned long long cap = std::max<unsigned long long>(1 << 16, 16ULL * nb), produced = 0;
            for (int attempt = 0; attempt < 2; ++attempt) {
                DevBuf<int32_t> rec(7 * cap);
                DevBuf<unsigned long long> cnt(1);
                VB_CUDA(cudaMemsetAsync(cnt.p, 0, sizeof(unsigned long long), st));
                VB_CUDA(cudaMemsetAsync(d_cursor.p, 0, sizeof(unsigned int), st));
                RegionSink sink = {rec.p, cnt.p, cap};
                parse_launch(ctx, dg, rb, d_pref.p, d_pqry.p, nb, P, d_cursor.p, d_stats.p, st, &sink, 0);
                VB_CUDA(cudaMemcpyAsync(&produced, cnt.p, sizeof(produced), cudaMemcpyDeviceToHost, st));
                VB_CUDA(cudaStreamSynchronize(st));
                if (produced <= cap) {
                    const size_t at = regions->size();
                    regions->resize(at + 7 * produced);
                    if (produced) {
                        VB_CUDA(cudaMemcpyAsync(regions->data() + at, rec.p, sizeof(int32_t) * 7 * produced, cudaMemcpyDeviceToHost, st));
                        VB_CUDA(cudaStreamSynchronize(st));
                    }
                    for (unsigned long long k = 0; k < produced; ++k)          // batch-local pair number -> caller's pair index
                        (*regions)[at + 7 * k] = (int32_t)order[pos + (uint32_t)(*regions)[at + 7 * k]];
                    break;
                }
                if (attempt == 1) throw vb_error(VB_ERR_INTERNAL, "region buffer overflow on the sized re-run");
                cap = produced;
            }
        }
        t_par.stop();
        std::vector<int32_t> tmp(3 * (size_t)nb);
        VB_CUDA(cudaMemcpyAsync(tmp.data(), d_stats.p, sizeof(int32_t) * 3 * nb, cudaMemcpyDeviceToHost, st));
        VB_CUDA(cudaStreamSynchronize(st));
        const auto h1 = std::chrono::steady_clock::now();
        for (uint32_t j = 0; j < nb; ++j) {
            uint64_t o = order[pos + j];
            stats[3 * o] = tmp[3 * j]; stats[3 * o + 1] = tmp[3 * j + 1]; stats[3 * o + 2] = tmp[3 * j + 2];
        }
        ms_index += job->prebuilt ? job->t_idx.ms() : t_idx.ms();
        ms_parse += t_par.ms();
        host_post_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h1).count();
        pos = end;
        ++n_batches;
    }
    job->t_all.stop();
    VB_CUDA(cudaStreamSynchronize(st));
    ctx->set_timing("align.total_ms", n ? job->t_all.ms() : 0.0);
    ctx->set_timing("align.upload_pack_ms", job->t_up.ms());
    ctx->set_timing("align.index_ms", ms_index);
    ctx->set_timing("align.parse_ms", ms_parse);
    ctx->set_timing("align.host_prep_ms", host_prep_ms);
    ctx->set_timing("align.host_post_ms", host_post_ms);
    ctx->set_timing("align.batches", n_batches);
    ctx->set_timing("align.pairs", (double)n);
}

void vb_align_pairs_impl(vb_ctx *ctx, const vb_genomes *g, const uint32_t *ref, const uint32_t *qry, uint64_t n,
                         const vb_align_params *ap, int32_t *stats)
{
    const uint32_t ng = g->count();
    std::vector<uint8_t> is_ref(ng, 0);
    for (uint64_t i = 0; i < n; ++i) {
        if (ref[i] >= ng || qry[i] >= ng) throw vb_error(VB_ERR_ARG, "pair id out of range");
        is_ref[ref[i]] = 1;
    }
    vb_align_job *job = vb_align_job_begin(ctx, g, ap, is_ref.data());
    try {
        vb_align_job_run(job, ref, qry, n, stats, nullptr, nullptr);
    } catch (...) {
        vb_align_job_end(job);
        throw;
    }
    vb_align_job_end(job);
}


// ---------------------------------------------------------------------------------------------------------------
// The fast path of vb_align / vb_shard_align: everything between the candidate list and the statistics stays on the
// device.  `meta` supplies lengths and the LZ-ANI order of ALL genomes, `store` their packed records; this rank parses
// the directed pairs whose reference it owns (owner(g) = g % world).  Returns false when the reference side (texts +
// anchor tables of all owned genomes) does not fit the memory budget -- the caller then takes the batched host path.
// ---------------------------------------------------------------------------------------------------------------
bool vb_align_fast(vb_ctx *ctx, const vb_genomes *meta, const DevGenomes &store, const vb_align_params *ap, const uint64_t *d_pairs,
                   const float *d_ani, uint64_t n_pairs, bool all_vs_all, uint32_t world, uint32_t me, AlignFastOut &out)
{
    if (ap->mal < 4 || ap->mal > 31 || ap->msl < 2 || ap->msl > ap->mal || ap->msl > 31)
        throw vb_error(VB_ERR_ARG, "need 2 <= msl <= mal <= 31");
    if (ap->aw < 1 || ap->aw > 32 || ap->ar < 1 || ap->ar > 32 || ap->am < 0)
        throw vb_error(VB_ERR_ARG, "need 1 <= aw <= 32, 1 <= ar <= 32, am >= 0");
    if (ap->mrd < 1 || ap->mrd > 4096 || ap->mqd < 0) throw vb_error(VB_ERR_ARG, "need 1 <= mrd <= 4096, mqd >= 0");
    cudaStream_t st = (cudaStream_t)ctx->stream;
    VB_CUDA(cudaSetDevice(ctx->device));
    const uint32_t ng = meta->count();
    const auto h0 = std::chrono::steady_clock::now();
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, ctx->device);

    // reference side: descriptors of every genome this rank owns (which of them are references is decided on the device)
    RefBatch rb;
    std::vector<int32_t> slot_of_gid(ng, -1);
    uint64_t need = 0;
    for (uint32_t g = me; g < ng; g += world) need += ref_bytes(meta->length(g), ap->mrd);
    if (need > ref_budget(ctx)) return false;
    for (uint32_t g = me; g < ng; g += world) { slot_of_gid[g] = (int32_t)rb.refs.size(); ref_batch_add(rb, meta, g, ap->mrd); }
    const uint32_t n_refs = (uint32_t)rb.refs.size();
    uint64_t n_cap = all_vs_all ? (uint64_t)ng * (ng > 0 ? ng - 1 : 0) : 2 * n_pairs;
    if (world > 1 && all_vs_all) throw vb_error(VB_ERR_ARG, "all-vs-all is not available in the multi-GPU pipeline");
    if (n_cap >= (1ULL << 32) - rsort::TILE) throw vb_error(VB_ERR_ARG, "more than 2^32 pairs in one call");
    out.n = 0;
    out.gbits = 1;
    while ((1ULL << out.gbits) < ng) out.gbits++;
    if (n_cap == 0 || n_refs == 0) {
        ctx->set_timing("align.pairs", 0.0);
        return true;
    }
    EventTimer t_all(st), t_list(st), t_idx(st), t_par(st);
    t_all.start();
    t_list.start();
    const std::vector<uint32_t> &order = vb_lz_order(meta), &rank = vb_lz_rank(meta);
    DevBuf<uint32_t> d_order(ng), d_rank(ng);
    DevBuf<int32_t> d_slot(ng);
    DevBuf<uint8_t> d_is_ref(n_refs);
    DevBuf<unsigned long long> counts(8);                    // [0] directed pairs, [1] heavy pairs, [2] class cut, [3] parse cursor
    VB_CUDA(cudaMemcpyAsync(d_order.p, order.data(), sizeof(uint32_t) * ng, cudaMemcpyHostToDevice, st));
    VB_CUDA(cudaMemcpyAsync(d_rank.p, rank.data(), sizeof(uint32_t) * ng, cudaMemcpyHostToDevice, st));
    VB_CUDA(cudaMemcpyAsync(d_slot.p, slot_of_gid.data(), sizeof(int32_t) * ng, cudaMemcpyHostToDevice, st));
    VB_CUDA(cudaMemsetAsync(d_is_ref.p, all_vs_all ? 1 : 0, n_refs, st));
    VB_CUDA(cudaMemsetAsync(counts.p, 0, counts.bytes(), st));
    rb.d_refs.alloc(n_refs);
    VB_CUDA(cudaMemcpyAsync(rb.d_refs.p, rb.refs.data(), sizeof(RefDesc) * n_refs, cudaMemcpyHostToDevice, st));

    // directed pair list, sorted by (reference, query) in LZ-ANI ids = the order of the result
    const uint64_t n_pad = (n_cap + rsort::TILE - 1) / rsort::TILE * rsort::TILE;
    DevBuf<uint64_t> ka(n_pad), kb(n_pad);
    DevBuf<uint32_t> ca(n_pad), cb(n_pad);
    auto grid = [&](uint64_t n) { return (int)std::max<uint64_t>(1, std::min<uint64_t>((n + 255) / 256, (uint64_t)n_sm * 16)); };
    if (all_vs_all) {
        all_pairs_kernel<<<grid(n_cap), 256, 0, st>>>(ng, out.gbits, ka.p, ca.p, counts.p);
        VB_LAUNCH_CHECK(ctx);
        if (n_pad > n_cap) { fill_sentinel_kernel<<<grid(n_pad - n_cap), 256, 0, st>>>(ka.p, n_cap, n_pad); VB_LAUNCH_CHECK(ctx); }
    } else {
        if (world > 1 || n_pad > n_cap) { fill_sentinel_kernel<<<grid(n_pad), 256, 0, st>>>(ka.p, world > 1 ? 0 : n_cap, n_pad); VB_LAUNCH_CHECK(ctx); }
        expand_pairs_kernel<<<grid(n_pairs), 256, 0, st>>>(d_pairs, d_ani, n_pairs, d_rank.p, store.glen.p, out.gbits, world, me, d_slot.p,
                                                           ka.p, ca.p, counts.p, d_is_ref.p);
        VB_LAUNCH_CHECK(ctx);
    }
    t_list.stop();
    // reference texts + anchor tables of the genomes that turned out to be references (overlaps nothing on the host:
    // everything here is enqueue-only)
    t_idx.start();
    rb.ref_rec.alloc(rb.recs + 8);
    rb.ht.alloc(rb.slots);
    ref_batch_launch(ctx, rb, store, ap, st, d_is_ref.p);
    t_idx.stop();
    EventTimer t_sched(st);
    t_sched.start();
    rsort::Workspace ws;
    const uint64_t *skeys = ka.p;
    const uint32_t *scost = ca.p;
    if (!all_vs_all) {                                       // (the all-vs-all list is generated in order)
        const bool in_b = rsort::sort_kv<8>(ctx, ka.p, ca.p, kb.p, cb.p, n_pad, 2 * out.gbits, ws);
        skeys = in_b ? kb.p : ka.p; scost = in_b ? cb.p : ca.p;
    }
    DevBuf<uint32_t> cls(COST_CLASSES), heavy(n_cap);
    DevBuf<uint8_t> heavy_flag(n_cap);
    static const bool lpt_off = getenv("VB_ALIGN_NO_LPT") != nullptr;
    static const int heavy_div = getenv("VB_ALIGN_HEAVY_DIV") ? std::max(1, atoi(getenv("VB_ALIGN_HEAVY_DIV"))) : 8;
    if (!all_vs_all && d_ani && !lpt_off) {
        VB_CUDA(cudaMemsetAsync(cls.p, 0, cls.bytes(), st));
        cost_hist_kernel<<<grid(n_cap), 256, 0, st>>>(scost, counts.p, cls.p);
        VB_LAUNCH_CHECK(ctx);
        lpt_cut_kernel<<<1, 1024, 0, st>>>(cls.p, counts.p, heavy_div);
        VB_LAUNCH_CHECK(ctx);
        heavy_fill_kernel<<<grid(n_cap), 256, 0, st>>>(scost, counts.p, cls.p, heavy.p, heavy_flag.p);
        VB_LAUNCH_CHECK(ctx);
    } else
        VB_CUDA(cudaMemsetAsync(heavy_flag.p, 0, n_cap, st));          // counts[1] stays 0: list order
    t_sched.stop();

    // the parse
    t_par.start();
    out.stats.alloc(3 * n_cap);
    LzParams P = {ap->mal, ap->msl, ap->mrd, ap->mqd, ap->reg, ap->aw, ap->am, ap->ar};
    SchedView V = {skeys, out.gbits, d_order.p, d_slot.p, heavy.p, heavy_flag.p, counts.p};
    static const int minb = getenv("VB_PARSE_MINB") ? atoi(getenv("VB_PARSE_MINB")) : 7;
    auto kern = minb >= 8 ? parse_sched_kernel<8> : (minb == 7 ? parse_sched_kernel<7> : (minb == 6 ? parse_sched_kernel<6> : parse_sched_kernel<5>));
    int per_sm = 0;
    VB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, 0));
    const int blocks = std::max(1, std::min<int>(per_sm * n_sm, (int)std::min<uint64_t>((n_cap + 3) / 4, 1u << 30)));
    kern<<<blocks, 128, 0, st>>>(store.rec.p, store.gofs.p, store.glen.p, rb.d_refs.p, rb.ref_rec.p, rb.ht.p, V, P, counts.p + 3, out.stats.p);
    VB_LAUNCH_CHECK(ctx);
    t_par.stop();
    // the sorted keys must outlive this scope: keep whichever buffer holds them
    out.keys = (skeys == ka.p) ? std::move(ka) : std::move(kb);
    t_all.stop();
    const double host_prep_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h0).count();
    unsigned long long n_dir = 0;
    VB_CUDA(cudaMemcpyAsync(&n_dir, counts.p, sizeof(n_dir), cudaMemcpyDeviceToHost, st));
    VB_CUDA(cudaStreamSynchronize(st));
    out.n = n_dir;
    ctx->set_timing("align.total_ms", t_all.ms());
    ctx->set_timing("align.upload_pack_ms", 0.0);
    ctx->set_timing("align.list_ms", t_list.ms() + t_sched.ms());
    ctx->set_timing("align.index_ms", t_idx.ms());
    ctx->set_timing("align.parse_ms", t_par.ms());
    ctx->set_timing("align.host_prep_ms", host_prep_ms);
    ctx->set_timing("align.batches", 1.0);
    ctx->set_timing("align.pairs", (double)n_dir);
    return true;
}
