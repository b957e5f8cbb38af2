// Device-side plumbing shared by the .cu files: error checks, RAII buffers, the packed genome store.
#pragma once
#include <cuda_runtime.h>

#include <functional>
#include <string>
#include <vector>

#include "vb_internal.h"

#define VB_CUDA(expr)                                                                                      \
    do {                                                                                                   \
        cudaError_t _e = (expr);                                                                           \
        if (_e != cudaSuccess)                                                                             \
            throw vb_error(VB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + \
                                            ":" + std::to_string(__LINE__) + ")");                        \
    } while (0)

#define VB_LAUNCH_CHECK(ctx)             \
    do {                                 \
        (ctx)->launches++;               \
        VB_CUDA(cudaGetLastError());     \
    } while (0)

// Device memory for call-scoped temporaries comes from a per-context ARENA: one slab that only grows, handed out by
// bumping a pointer and popped in LIFO order (C++ scopes destroy DevBufs in reverse order), reset at the start of
// every top-level call.  After warm-up there is no cudaMalloc / cudaFree (and none of their implicit device
// synchronisations or page mapping) on the hot path.  Buffers that outlive a call (resident genomes) are allocated
// while vb_tls_arena == nullptr and use plain cudaMalloc.
struct vb_arena {
    struct Slab { char *base; size_t cap; size_t off; };
    std::vector<Slab> slabs;
    size_t peak = 0, live = 0;
    void *alloc(size_t bytes);
    void pop(void *p, size_t bytes);
    void reset(cudaStream_t st);          // call with no arena buffer alive
    void destroy();
};
extern thread_local vb_arena *vb_tls_arena;
extern thread_local cudaStream_t vb_tls_stream;
// true while buffers that OUTLIVE the call are being allocated (packed genome stores): they come from the device's
// stream-ordered pool (cudaMallocAsync on the context stream; release threshold set to "never" in vb_ctx_create), so
// that dropping one store and uploading the next -- every step of the host-buffer path -- costs no cudaMalloc/cudaFree
extern thread_local bool vb_tls_pool_alloc;

template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    vb_arena *arena = nullptr;
    cudaStream_t pool_stream = nullptr;      // non-null: allocated with cudaMallocAsync on this stream
    DevBuf() = default;
    explicit DevBuf(size_t count) { alloc(count); }
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n), arena(o.arena), pool_stream(o.pool_stream) { o.p = nullptr; o.n = 0; }
    DevBuf &operator=(DevBuf &&o) noexcept
    {
        if (this != &o) { release(); p = o.p; n = o.n; arena = o.arena; pool_stream = o.pool_stream; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    void alloc(size_t count)
    {
        release();
        n = count;
        if (!count) return;
        arena = nullptr; pool_stream = nullptr;
        cudaError_t e;
        if (vb_tls_pool_alloc && vb_tls_stream) {
            e = cudaMallocAsync((void **)&p, count * sizeof(T), vb_tls_stream);
            if (e == cudaSuccess) { pool_stream = vb_tls_stream; return; }
        } else {
            static const bool no_arena = getenv("VB_NO_ARENA") != nullptr;      // debugging: one cudaMalloc per buffer, so
            arena = no_arena ? nullptr : vb_tls_arena;                          // compute-sanitizer sees every overrun
            if (arena) { p = (T *)arena->alloc(count * sizeof(T)); return; }
            e = cudaMalloc((void **)&p, count * sizeof(T));
        }
        if (e != cudaSuccess) {
            p = nullptr; n = 0;
            cudaGetLastError();
            throw vb_error(VB_ERR_MEM, "cudaMalloc of " + std::to_string(count * sizeof(T)) + " bytes failed: " +
                                           cudaGetErrorString(e));
        }
    }
    void release()
    {
        if (p) {
            if (pool_stream) cudaFreeAsync(p, pool_stream);
            else if (arena) arena->pop(p, n * sizeof(T));
            else cudaFree(p);
        }
        p = nullptr; n = 0;
    }
    size_t bytes() const { return n * sizeof(T); }
};

// bytes this context could still allocate: free device memory + what the pool holds but is not using
uint64_t vb_device_available(vb_ctx *ctx);

// CUDA events are recycled through a per-thread free list: creating and destroying a dozen events per call cost more
// host time than the calls' own bookkeeping.  Events are device-bound: the list is keyed by the device the event was
// created on (remembered by the timer, not re-queried when it is returned).  The list destroys its events when the
// thread exits.
struct EventPool {
    std::vector<cudaEvent_t> free_list[64];
    ~EventPool()
    {
        for (auto &fl : free_list)
            for (cudaEvent_t e : fl) cudaEventDestroy(e);      // best effort: the context may already be gone at thread exit
    }
    cudaEvent_t get(int dev)
    {
        auto &fl = free_list[dev & 63];
        if (!fl.empty()) { cudaEvent_t e = fl.back(); fl.pop_back(); return e; }
        cudaEvent_t e = nullptr;
        cudaError_t rc = cudaEventCreate(&e);
        if (rc != cudaSuccess) {
            cudaGetLastError();
            throw vb_error(VB_ERR_CUDA, std::string("cudaEventCreate: ") + cudaGetErrorString(rc));
        }
        return e;
    }
    void put(int dev, cudaEvent_t e) { free_list[dev & 63].push_back(e); }
};
inline EventPool &vb_event_pool() { static thread_local EventPool p; return p; }

struct EventTimer {
    cudaEvent_t a = nullptr, b = nullptr;
    cudaStream_t s;
    int dev = 0;
    bool started = false, stopped = false;
    explicit EventTimer(cudaStream_t st) : s(st)
    {
        if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; }
        a = vb_event_pool().get(dev);
        try { b = vb_event_pool().get(dev); } catch (...) { vb_event_pool().put(dev, a); throw; }
    }
    EventTimer(const EventTimer &) = delete;
    EventTimer &operator=(const EventTimer &) = delete;
    ~EventTimer() { vb_event_pool().put(dev, a); vb_event_pool().put(dev, b); }
    void start() { cudaEventRecord(a, s); started = true; stopped = false; }
    void stop() { cudaEventRecord(b, s); stopped = true; }
    // 0 for a timer that was not started and stopped in this use (a recycled event still holds its previous record)
    double ms()
    {
        if (!started || !stopped) return 0.0;
        cudaEventSynchronize(b);
        float t = 0;
        if (cudaEventElapsedTime(&t, a, b) != cudaSuccess) { cudaGetLastError(); return 0.0; }
        return t;
    }
};

// ---------------------------------------------------------------------------------------------------------------
// Packed genome store in HBM.
//   Genome g occupies base slots [gofs[g], gofs[g] + glen[g]) of one global base axis; every genome starts at a
//   multiple of 128 bases and is followed by at least 128 slots of padding marked invalid, so a kernel may read up
//   to 64 bases past the end of a genome without a bounds check.
//   seq2 : 2 bits per base, 16 bases per uint32 word, base b of a word at bits [2b, 2b+1]  (A0 C1 G2 T3; U is stored as T)
//   inv_kdb : 1 bit per base, 32 bases per uint32 word, set when the base is not a valid symbol for kmer-db or is
//          padding.  The tools disagree on U: kmer-db reads it as T (alphabet.h:80-85), lz-ani as N
//          (seq_reservoir.h:243-247) -- the lz-ani flags live in `rec`; everything else is shared, so ONE upload serves
//          the prefilter and the align stage.
//   rec  : one uint4 per 32 slots = {bit 0 of the 32 codes, bit 1 of the 32 codes, inv_lz word, inv_kdb word}.  The LZ parse
//          compares texts 32 bases at a time: with bit PLANES a comparison is (lo^lo')|(hi^hi')|N|N' straight from two
//          128-bit loads per text, where the interleaved 2-bit form needs a 36-instruction bit squeeze per comparison
//   tile_gid : genome id owning each 128-base tile (0xffffffff for none)
// ---------------------------------------------------------------------------------------------------------------
struct DevGenomes {
    uint32_t n = 0;
    uint64_t total_slots = 0;        // multiple of 128
    DevBuf<uint32_t> seq2, inv_kdb;
    DevBuf<uint4> rec;               // the align stage's view: per 32 slots {low bit plane, high bit plane, inv_lz, inv_kdb}
    DevBuf<uint64_t> gofs;           // n entries
    DevBuf<uint32_t> glen;           // n entries
    DevBuf<uint32_t> tile_gid;       // total_slots / 128
    std::vector<uint64_t> h_gofs;
    std::vector<uint32_t> h_glen;
    uint32_t min_pad = 0;            // invalid slots guaranteed after every genome
};

// Candidate list of a prefilter call, kept on the device (stream-ordered pool) so that the align stage that follows
// builds its directed pair list without a round trip through the host.
struct DevPairs {
    uint64_t uid = 0;                // vb_pairs_uid of the host list this mirrors
    uint64_t n = 0;
    DevBuf<uint64_t> keys;           // row << 32 | col (input-order ids, row > col), sorted
    DevBuf<float> ani;               // device estimate of ani-shorter: scheduling cost model only
};

// Upload ASCII and pack on the device.  min_pad: invalid slots guaranteed after every genome (>= VB_STORE_PAD).
constexpr uint32_t VB_STORE_PAD = 128 + 64;      // covers the prefilter (128) and align with --mrd up to 64
// on_chunk(slot_lo, slot_hi), optional: called after the pack of each chunk of genomes has been enqueued -- the H2D of
// the next chunk runs on the context's copy stream meanwhile, so a consumer that enqueues its first pass over
// [slot_lo, slot_hi) from the callback overlaps that pass with the transfer.
struct DevGenomes;
using vb_chunk_fn = std::function<void(const DevGenomes &, uint64_t, uint64_t)>;
uint64_t vb_store_slots(const vb_genomes *g, uint32_t min_pad);    // base slots the packed store of g will have
// force_slots (0 = natural size): make the store exactly that many slots long (>= the natural size; the tail is invalid)
// -- the blocks of a multi-GPU run are padded to a common size so that they can be all-gathered
void vb_upload_genomes(vb_ctx *ctx, const vb_genomes *g, DevGenomes &out, uint32_t min_pad = VB_STORE_PAD,
                       const vb_chunk_fn *on_chunk = nullptr, uint64_t force_slots = 0);

// Returns the packed copy of g: the resident one when vb_genomes_make_resident was called for it, else the copy the
// previous call on this context uploaded (vclust prefilter followed by vclust align on the same set transfers the
// genomes once), else uploads now and keeps the copy for the next call (vb_genomes_evict drops it).
bool vb_has_dev_genomes(vb_ctx *ctx, const vb_genomes *g, uint32_t min_pad);
const DevGenomes &vb_get_dev_genomes(vb_ctx *ctx, const vb_genomes *g, uint32_t min_pad, bool *was_resident = nullptr,
                                     const vb_chunk_fn *on_chunk = nullptr);

// device-side result of vb_align_fast (align.cu): directed pairs sorted by (reference, query) in LZ-ANI ids, compact keys
// ref << gbits | query, and 3 ints per pair; the buffers live in the call's arena
struct AlignFastOut {
    DevBuf<uint64_t> keys;
    DevBuf<int32_t> stats;
    uint64_t n = 0;
    int gbits = 0;
};
bool vb_align_fast(vb_ctx *ctx, const vb_genomes *meta, const DevGenomes &store, const vb_align_params *ap, const uint64_t *d_pairs,
                   const float *d_ani, uint64_t n_pairs, bool all_vs_all, uint32_t world, uint32_t me, AlignFastOut &out);

#ifdef __CUDACC__
// 32 bases (64 bits) starting at base slot p of a 2-bit array; base p lands in bits [0,1].
__device__ __forceinline__ uint64_t fetch2(const uint32_t *__restrict__ w, uint64_t p)
{
    const uint32_t *q = w + (p >> 4);
    uint32_t a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    unsigned sh = (unsigned)(p & 15) * 2;
    uint32_t lo = __funnelshift_r(a, b, sh);
    uint32_t hi = __funnelshift_r(b, c, sh);
    return ((uint64_t)hi << 32) | lo;
}
// 32 flag bits starting at base slot p of a 1-bit array.
__device__ __forceinline__ uint32_t fetch1(const uint32_t *__restrict__ w, uint64_t p)
{
    const uint32_t *q = w + (p >> 5);
    uint32_t a = __ldg(q), b = __ldg(q + 1);
    return __funnelshift_r(a, b, (unsigned)(p & 31));
}
// squeeze the even bits of a 64-bit word (one flag per 2-bit base) into 32 bits
__device__ __forceinline__ uint32_t squeeze_even(uint64_t x)
{
    x &= 0x5555555555555555ULL;
    x = (x | (x >> 1)) & 0x3333333333333333ULL;
    x = (x | (x >> 2)) & 0x0f0f0f0f0f0f0f0fULL;
    x = (x | (x >> 4)) & 0x00ff00ff00ff00ffULL;
    x = (x | (x >> 8)) & 0x0000ffff0000ffffULL;
    x = (x | (x >> 16)) & 0x00000000ffffffffULL;
    return (uint32_t)x;
}
// per-base mismatch flags of two 32-base words
__device__ __forceinline__ uint32_t mismatch32(uint64_t a, uint64_t b)
{
    uint64_t x = a ^ b;
    return squeeze_even(x | (x >> 1));
}
#endif
