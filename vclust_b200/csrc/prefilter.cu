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

__device__ __forceinline__ void table_add(uint64_t *__restrict__ tkeys, uint32_t *__restrict__ tvals, uint64_t cap_mask,
                                          uint64_t key, uint32_t inc, int *__restrict__ overflow)
{
    uint64_t h = fmix64(key) & cap_mask;
    for (uint64_t probes = 0; probes <= cap_mask; ++probes) {
        uint64_t cur = tkeys[h];
        if (cur == SLOT_EMPTY) {
            cur = atomicCAS((unsigned long long *)&tkeys[h], (unsigned long long)SLOT_EMPTY, (unsigned long long)key);
            if (cur == SLOT_EMPTY) cur = key;
        }
        if (cur == key) { atomicAdd(&tvals[h], inc); return; }
        h = (h + 1) & cap_mask;
    }
    *overflow = 1;
}

// k3b: duplicates per genome; every distinct (k-mer, genome) occurrence pairs with the distinct genomes before it in the run;
// row = the later (larger) genome id, col = the earlier one -- the lower triangle of all2all_sp.
__global__ void __launch_bounds__(256) segment_pairs_kernel(const uint64_t *__restrict__ keys,
                                                            const uint32_t *__restrict__ vals, uint64_t n,
                                                            uint32_t *__restrict__ dup_cnt, uint64_t *__restrict__ tkeys,
                                                            uint32_t *__restrict__ tvals, uint64_t cap_mask,
                                                            int *__restrict__ overflow)
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
                table_add(tkeys, tvals, cap_mask, ((uint64_t)g << 32) | gj, 1u, overflow);
                prev = gj;
            }
        }
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
    if (p->max_seqs > 0) throw vb_error(VB_ERR_ARG, "--max-seqs is not implemented on the GPU path yet");
    cudaStream_t st = (cudaStream_t)ctx->stream;
    VB_CUDA(cudaSetDevice(ctx->device));
    const uint32_t n = g->count();
    EventTimer t_all(st), t_up(st), t_ext(st), t_sort(st), t_seg(st), t_emit(st);

    t_all.start();
    t_up.start();
    DevGenomes dg_scratch;
    const DevGenomes &dg = vb_get_dev_genomes(ctx, g, /*u_is_t=*/true, 128, dg_scratch);
    t_up.stop();

    // ---- k1
    t_ext.start();
    const uint64_t n_slots = dg.total_slots;
    const uint64_t n_pad = ((n_slots + rsort::TILE - 1) / rsort::TILE) * rsort::TILE;
    DevBuf<uint64_t> keys_a(n_pad), keys_b(n_pad);
    DevBuf<uint32_t> vals_a(n_pad), vals_b(n_pad);
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
    extract_kernel<<<grid_for(n_pad), 256, 0, st>>>(dg.seq2.p, dg.inv.p, dg.tile_gid.p, n_slots, n_pad, ep, keys_a.p,
                                                   vals_a.p, valid_cnt);
    VB_LAUNCH_CHECK(ctx);
    t_ext.stop();

    // ---- k2 (bit 2k+shift is set only in the sentinel, so it sorts last)
    t_sort.start();
    rsort::Workspace ws;
    const int key_bits = 2 * p->k + ep.shift + 1;
    bool in_b = rsort::sort_kv<8>(ctx, keys_a.p, vals_a.p, keys_b.p, vals_b.p, n_pad, key_bits, ws);
    const uint64_t *skeys = in_b ? keys_b.p : keys_a.p;
    const uint32_t *svals = in_b ? vals_b.p : vals_a.p;
    t_sort.stop();

    // ---- k3
    t_seg.start();
    DevBuf<unsigned long long> scalars(4);
    VB_CUDA(cudaMemsetAsync(scalars.p, 0, scalars.bytes(), st));
    unsigned long long max_pairs = (unsigned long long)n * (n > 0 ? n - 1 : 0) / 2;
    unsigned long long n_inc = max_pairs;
    const bool count_first = max_pairs > (1ULL << 26);      // otherwise the dense bound N(N-1)/2 sizes the table
    if (count_first) {
        segment_count_kernel<<<grid_for(n_pad), 256, 0, st>>>(skeys, svals, n_pad, scalars.p);
        VB_LAUNCH_CHECK(ctx);
        VB_CUDA(cudaMemcpyAsync(&n_inc, scalars.p, sizeof(n_inc), cudaMemcpyDeviceToHost, st));
        VB_CUDA(cudaStreamSynchronize(st));
    }
    unsigned long long distinct_bound = std::min(n_inc, max_pairs);
    uint64_t cap = 1024;
    while (cap < 2 * distinct_bound) cap <<= 1;
    if (cap > (1ULL << 32)) throw vb_error(VB_ERR_MEM, "pair table would exceed 2^32 slots; split the input (--batch-size)");
    DevBuf<uint64_t> tkeys(cap);
    DevBuf<uint32_t> tvals(cap);
    DevBuf<int> overflow(1);
    VB_CUDA(cudaMemsetAsync(tkeys.p, 0xff, tkeys.bytes(), st));       // SLOT_EMPTY
    VB_CUDA(cudaMemsetAsync(tvals.p, 0, tvals.bytes(), st));
    VB_CUDA(cudaMemsetAsync(overflow.p, 0, sizeof(int), st));
    segment_pairs_kernel<<<grid_for(n_pad), 256, 0, st>>>(skeys, svals, n_pad, dup_cnt, tkeys.p, tvals.p, cap - 1, overflow.p);
    VB_LAUNCH_CHECK(ctx);
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
    emit_kernel<<<grid_for(cap), 256, 0, st>>>(tkeys.p, tvals.p, cap, totals, em, 0, scalars.p + 1, nullptr, nullptr);
    VB_LAUNCH_CHECK(ctx);
    unsigned long long n_emit = 0;
    int h_overflow = 0;
    VB_CUDA(cudaMemcpyAsync(&n_emit, scalars.p + 1, sizeof(n_emit), cudaMemcpyDeviceToHost, st));
    VB_CUDA(cudaMemcpyAsync(&h_overflow, overflow.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    VB_CUDA(cudaStreamSynchronize(st));
    if (h_overflow) throw vb_error(VB_ERR_INTERNAL, "pair table overflow");
    const uint64_t e_pad = ((n_emit + rsort::TILE - 1) / rsort::TILE) * rsort::TILE;
    std::vector<uint64_t> h_keys(n_emit);
    std::vector<uint32_t> h_vals(n_emit);
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
    }
    std::vector<uint32_t> h_tot(n);
    if (n) VB_CUDA(cudaMemcpyAsync(h_tot.data(), totals, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, st));
    t_emit.stop();
    t_all.stop();
    VB_CUDA(cudaStreamSynchronize(st));

    // ---- host: exact IEEE-double metric (params.cpp:28-32) and the two -min filters (sparse_filters.h:49-61)
    std::vector<uint64_t> keep;
    keep.reserve(n_emit);
    std::vector<double> ani(n_emit);
    const uint64_t cmask = (1ULL << em.gbits) - 1;
    for (uint64_t i = 0; i < n_emit; ++i) {
        uint32_t r = (uint32_t)(h_keys[i] >> em.gbits), c = (uint32_t)(h_keys[i] & cmask);
        if (partial) { ani[i] = 0; keep.push_back(i); continue; }
        ani[i] = vb_ani_shorter(h_vals[i], h_tot[r], h_tot[c], p->k);
        if (ani[i] >= p->min_ident) keep.push_back(i);
    }
    vb_pairs *res = vb_pairs_alloc(keep.size(), n);
    for (uint64_t o = 0; o < keep.size(); ++o) {
        uint64_t i = keep[o];
        res->row[o] = (uint32_t)(h_keys[i] >> em.gbits);
        res->col[o] = (uint32_t)(h_keys[i] & cmask);
        res->common[o] = h_vals[i];
        res->ani[o] = ani[i];
    }
    for (uint32_t i = 0; i < n; ++i) res->total_kmers[i] = h_tot[i];
    res->k = p->k;
    res->kmers_fraction = p->kmers_fraction;
    *out_pairs = res;

    ctx->set_timing("prefilter.total_ms", t_all.ms());
    ctx->set_timing("prefilter.upload_pack_ms", t_up.ms());
    ctx->set_timing("prefilter.extract_ms", t_ext.ms());
    ctx->set_timing("prefilter.sort_ms", t_sort.ms());
    ctx->set_timing("prefilter.segment_ms", t_seg.ms());
    ctx->set_timing("prefilter.emit_ms", t_emit.ms());
    ctx->set_timing("prefilter.tuples", (double)n_slots);
    ctx->set_timing("prefilter.pair_increments", (double)n_inc);
    ctx->set_timing("prefilter.table_slots", (double)cap);
    ctx->set_timing("prefilter.candidates", (double)n_emit);
}
