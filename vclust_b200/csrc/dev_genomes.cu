// Upload of a genome set and 2-bit packing on the device (one pass over the ASCII bytes).
#include "dev_util.cuh"

namespace {

// code table: 0..3 = ACGT, 4 = invalid.  Index 0: lz-ani rule (U invalid), 1: kmer-db rule (U == T).
__constant__ uint8_t c_code[2][256];

__global__ void pack_kernel(const uint8_t *__restrict__ ascii, const uint64_t *__restrict__ src_off,
                            const uint64_t *__restrict__ gofs, const uint32_t *__restrict__ glen,
                            const uint32_t *__restrict__ tile_gid, uint64_t c_lo, uint64_t c_hi,
                            uint32_t *__restrict__ seq2, uint32_t *__restrict__ inv_kdb, uint4 *__restrict__ rec)
{
    // one thread per 32 base slots [32 c, 32 c + 32), c in [c_lo, c_hi): two seq2 words, the kmer-db validity word and
    // the align stage's plane record
    for (uint64_t c = c_lo + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; c < c_hi;
         c += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t slot = c * 32;
        uint32_t gid = tile_gid[slot >> 7];
        uint32_t w0 = 0, w1 = 0, bad = 0xffffffffu, bad_lz = 0xffffffffu, p_lo = 0, p_hi = 0;
        if (gid != 0xffffffffu) {
            uint64_t local = slot - gofs[gid];
            uint32_t len = glen[gid];
            if (local < len) {
                const uint8_t *src = ascii + src_off[gid] + local;
                uint32_t m = (len - local) < 32 ? (uint32_t)(len - local) : 32u;
                bad = 0; bad_lz = 0;
#pragma unroll 8
                for (uint32_t j = 0; j < 32; ++j) {
                    uint32_t code = 4, code_lz = 4;
                    if (j < m) { code = c_code[1][src[j]]; code_lz = c_code[0][src[j]]; }
                    uint32_t two = code & 3;
                    if (code > 3) { bad |= 1u << j; two = 0; }
                    if (code_lz > 3) bad_lz |= 1u << j;
                    if (j < 16) w0 |= two << (2 * j);
                    else w1 |= two << (2 * (j - 16));
                    p_lo |= (two & 1u) << j;
                    p_hi |= (two >> 1) << j;
                }
            }
        }
        seq2[2 * c] = w0;
        seq2[2 * c + 1] = w1;
        inv_kdb[c] = bad;
        rec[c] = make_uint4(p_lo, p_hi, bad_lz, bad);
    }
}

// tile map from the store offsets: one warp per genome marks its 128-slot tiles (the array is pre-set to "none")
__global__ void tile_map_kernel(const uint64_t *__restrict__ gofs, const uint32_t *__restrict__ glen, uint32_t n,
                                uint32_t *__restrict__ tile_gid)
{
    const uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (g >= n) return;
    const uint64_t t0 = gofs[g] >> 7, t1 = (gofs[g] + glen[g] + 127) >> 7;       // genomes start on tile boundaries
    for (uint64_t t = t0 + lane; t < t1; t += 32) tile_gid[t] = g;
}

}  // namespace

uint64_t vb_store_slots(const vb_genomes *g, uint32_t min_pad)
{
    if (min_pad < VB_STORE_PAD) min_pad = VB_STORE_PAD;
    uint64_t slots = 0;
    for (uint32_t i = 0, n = g->count(); i < n; ++i) slots += ((g->length(i) + min_pad + 127) / 128) * 128;
    return slots + 128;
}

void vb_upload_genomes(vb_ctx *ctx, const vb_genomes *g, DevGenomes &out, uint32_t min_pad, const vb_chunk_fn *on_chunk,
                       uint64_t force_slots)
{
    if (min_pad < VB_STORE_PAD) min_pad = VB_STORE_PAD;
    if (g->skeleton) throw vb_error(VB_ERR_ARG, "this genome set holds names and lengths only (vb_genomes_skeleton): nothing to upload");
    cudaStream_t st = (cudaStream_t)ctx->stream;
    {   // one-time (per device) upload of the symbol table; contexts on several host threads may get here together
        static std::mutex table_mutex;
        static bool table_ready[64] = {false};
        std::lock_guard<std::mutex> lock(table_mutex);
        if (!table_ready[ctx->device & 63]) {
            uint8_t t[2][256];
            memset(t, 4, sizeof(t));
            const char *acgt = "ACGT";
            for (int r = 0; r < 2; ++r)
                for (int i = 0; i < 4; ++i) { t[r][(uint8_t)acgt[i]] = (uint8_t)i; t[r][(uint8_t)(acgt[i] | 0x20)] = (uint8_t)i; }
            t[1][(uint8_t)'U'] = 3; t[1][(uint8_t)'u'] = 3;
            VB_CUDA(cudaMemcpyToSymbol(c_code, t, sizeof(t)));
            table_ready[ctx->device & 63] = true;
        }
    }
    const uint32_t n = g->count();
    out.n = n;
    out.min_pad = min_pad;
    out.h_gofs.resize(n);
    out.h_glen.resize(n);
    uint64_t slots = 0;
    for (uint32_t i = 0; i < n; ++i) {
        uint64_t len = g->length(i);
        if (len > 0x7fff0000ULL) throw vb_error(VB_ERR_ARG, "genome longer than 2^31 bases: " + g->names[i]);
        out.h_gofs[i] = slots;
        out.h_glen[i] = (uint32_t)len;
        slots += ((len + min_pad + 127) / 128) * 128;  // >= min_pad invalid slots after every genome
    }
    slots += 128;
    if (force_slots) {
        if (force_slots < slots || force_slots % 128) throw vb_error(VB_ERR_INTERNAL, "vb_upload_genomes: bad forced store size");
        slots = force_slots;
    }
    out.total_slots = slots;

    out.seq2.alloc(slots / 16 + 8);
    out.inv_kdb.alloc(slots / 32 + 8);
    out.rec.alloc(slots / 32 + 8);
    out.gofs.alloc(n ? n : 1);
    out.glen.alloc(n ? n : 1);
    out.tile_gid.alloc(slots / 128);
    if (!g->pinned && !g->bases.empty() && ++g->uploads >= 2) {
        // Page-locking costs about as much as one pageable transfer (it touches every page), so it only pays when the
        // same host buffer is uploaded again: the first upload goes through the driver's staging path, from the second
        // one on the buffer is pinned and H2D runs at full PCIe rate.
        if (cudaHostRegister((void *)g->bases.data(), g->bases.size(), cudaHostRegisterDefault) == cudaSuccess) g->pinned = true;
        else cudaGetLastError();
    }
    const bool pool_mode = vb_tls_pool_alloc;
    vb_tls_pool_alloc = false;                   // the ASCII copy is call-scoped: arena
    DevBuf<uint8_t> ascii(g->bases.size() + 64);
    DevBuf<uint64_t> src_off(n + 1);
    vb_tls_pool_alloc = pool_mode;
    VB_CUDA(cudaMemsetAsync(out.seq2.p, 0, out.seq2.bytes(), st));
    VB_CUDA(cudaMemsetAsync(out.inv_kdb.p, 0xff, out.inv_kdb.bytes(), st));
    VB_CUDA(cudaMemsetAsync(out.rec.p, 0xff, out.rec.bytes(), st));          // slack records: all invalid
    VB_CUDA(cudaMemcpyAsync(src_off.p, g->offset.data(), sizeof(uint64_t) * (n + 1), cudaMemcpyHostToDevice, st));
    if (n) {
        VB_CUDA(cudaMemcpyAsync(out.gofs.p, out.h_gofs.data(), sizeof(uint64_t) * n, cudaMemcpyHostToDevice, st));
        VB_CUDA(cudaMemcpyAsync(out.glen.p, out.h_glen.data(), sizeof(uint32_t) * n, cudaMemcpyHostToDevice, st));
    }
    VB_CUDA(cudaMemsetAsync(out.tile_gid.p, 0xff, out.tile_gid.bytes(), st));
    if (n) {
        tile_map_kernel<<<(n * 32 + 255) / 256, 256, 0, st>>>(out.gofs.p, out.glen.p, n, out.tile_gid.p);
        VB_LAUNCH_CHECK(ctx);
    }
    // The genomes travel in up to 8 chunks (whole genomes, >= 4 MB each) on the copy stream; the pack of chunk c -- and
    // whatever the caller enqueues from on_chunk -- runs on the main stream while chunk c + 1 is in flight.
    cudaStream_t cs = (cudaStream_t)ctx->copy_stream;
    const uint64_t total_bytes = g->bases.size();
    const int max_chunks = (g->pinned && cs) ? (int)std::min<uint64_t>(8, std::max<uint64_t>(1, total_bytes >> 22)) : 1;
    cudaEvent_t ev0 = (cudaEvent_t)ctx->copy_events[8];
    if (max_chunks > 1) {                        // the copy stream may touch the fresh buffers only after this point of `st`
        VB_CUDA(cudaEventRecord(ev0, st));
        VB_CUDA(cudaStreamWaitEvent(cs, ev0, 0));
    }
    uint32_t g0 = 0;
    for (int c = 0; c < max_chunks && g0 < n; ++c) {
        uint32_t g1 = n;
        if (c + 1 < max_chunks) {
            const uint64_t want = total_bytes * (uint64_t)(c + 1) / max_chunks;
            g1 = (uint32_t)(std::lower_bound(g->offset.begin() + g0 + 1, g->offset.begin() + n, want) - g->offset.begin());
            g1 = std::min(std::max(g1, g0 + 1), n);
        }
        const uint64_t b0 = g->offset[g0], b1 = g->offset[g1];
        const uint64_t s0 = out.h_gofs[g0], s1 = g1 < n ? out.h_gofs[g1] : slots;
        if (b1 > b0) {
            if (max_chunks > 1) {
                VB_CUDA(cudaMemcpyAsync(ascii.p + b0, g->bases.data() + b0, b1 - b0, cudaMemcpyHostToDevice, cs));
                VB_CUDA(cudaEventRecord((cudaEvent_t)ctx->copy_events[c], cs));
                VB_CUDA(cudaStreamWaitEvent(st, (cudaEvent_t)ctx->copy_events[c], 0));
            } else
                VB_CUDA(cudaMemcpyAsync(ascii.p + b0, g->bases.data() + b0, b1 - b0, cudaMemcpyHostToDevice, st));
        }
        const uint64_t c_lo = s0 / 32, c_hi = s1 / 32;
        const int blocks = (int)std::max<uint64_t>(1, std::min<uint64_t>((c_hi - c_lo + 255) / 256, 148 * 16));
        pack_kernel<<<blocks, 256, 0, st>>>((const uint8_t *)ascii.p, src_off.p, out.gofs.p, out.glen.p, out.tile_gid.p, c_lo, c_hi,
                                            out.seq2.p, out.inv_kdb.p, out.rec.p);
        VB_LAUNCH_CHECK(ctx);
        if (on_chunk) (*on_chunk)(out, s0, s1);
        g0 = g1;
    }
    if (n == 0 && on_chunk) (*on_chunk)(out, 0, slots);
    // no synchronisation: the ASCII staging buffer is an arena block (re-used only by later work on this stream), the
    // small host arrays are pageable (staged by the runtime before cudaMemcpyAsync returns) or owned by g / out
}

static void drop_last(vb_ctx *ctx)
{
    if (!ctx->last.dev) return;
    cudaStream_t saved = vb_tls_stream;
    vb_tls_stream = (cudaStream_t)ctx->stream;
    delete ctx->last.dev;                         // stream-ordered frees: safe behind whatever still reads the buffers
    vb_tls_stream = saved;
    ctx->last = {nullptr, 0, 0, nullptr};
}

// upload into buffers that outlive the call (stream-ordered pool)
static DevGenomes *upload_persistent(vb_ctx *ctx, const vb_genomes *g, uint32_t min_pad, const vb_chunk_fn *on_chunk = nullptr,
                                     uint64_t force_slots = 0)
{
    auto *d = new DevGenomes();
    const bool saved = vb_tls_pool_alloc;
    vb_tls_pool_alloc = true;
    try { vb_upload_genomes(ctx, g, *d, min_pad, on_chunk, force_slots); } catch (...) { vb_tls_pool_alloc = saved; delete d; throw; }
    vb_tls_pool_alloc = saved;
    return d;
}

// would vb_get_dev_genomes find a device copy (no upload, on_chunk not called)?
bool vb_has_dev_genomes(vb_ctx *ctx, const vb_genomes *g, uint32_t min_pad)
{
    if (min_pad < VB_STORE_PAD) min_pad = VB_STORE_PAD;
    for (auto &r : ctx->resident)
        if (r.g == g && r.uid == g->uid && r.min_pad >= min_pad) return true;
    return ctx->last.dev && ctx->last.g == g && ctx->last.uid == g->uid && ctx->last.min_pad >= min_pad;
}

const DevGenomes &vb_get_dev_genomes(vb_ctx *ctx, const vb_genomes *g, uint32_t min_pad, bool *was_resident,
                                     const vb_chunk_fn *on_chunk)
{
    if (min_pad < VB_STORE_PAD) min_pad = VB_STORE_PAD;
    if (was_resident) *was_resident = true;
    for (auto &r : ctx->resident)
        if (r.g == g && r.uid == g->uid && r.min_pad >= min_pad) return *r.dev;
    if (ctx->last.dev && ctx->last.g == g && ctx->last.uid == g->uid && ctx->last.min_pad >= min_pad) return *ctx->last.dev;
    if (was_resident) *was_resident = false;
    drop_last(ctx);
    ctx->last = {g, g->uid, min_pad, upload_persistent(ctx, g, min_pad, on_chunk)};
    return *ctx->last.dev;
}

void vb_evict_impl(vb_ctx *ctx, const vb_genomes *g);

void vb_make_resident_impl(vb_ctx *ctx, const vb_genomes *g, uint32_t min_pad, uint64_t force_slots)
{
    if (min_pad < VB_STORE_PAD) min_pad = VB_STORE_PAD;
    if (force_slots) {                            // a block of a multi-GPU run: always a fresh store of exactly that size
        vb_evict_impl(ctx, g);
        ctx->resident.push_back({g, g->uid, min_pad, upload_persistent(ctx, g, min_pad, nullptr, force_slots)});
        return;
    }
    for (auto &r : ctx->resident)
        if (r.g == g && r.uid == g->uid && r.min_pad >= min_pad) return;
    if (ctx->last.dev && ctx->last.g == g && ctx->last.uid == g->uid && ctx->last.min_pad >= min_pad) {
        ctx->resident.push_back(ctx->last);       // promote the cached copy
        ctx->last = {nullptr, 0, 0, nullptr};
        return;
    }
    ctx->resident.push_back({g, g->uid, min_pad, upload_persistent(ctx, g, min_pad)});
}

void vb_evict_impl(vb_ctx *ctx, const vb_genomes *g)
{
    cudaStream_t saved = vb_tls_stream;
    vb_tls_stream = (cudaStream_t)ctx->stream;
    if (g == nullptr || ctx->last.g == g) drop_last(ctx);
    for (size_t i = 0; i < ctx->resident.size();) {
        if (g == nullptr || ctx->resident[i].g == g) {
            delete ctx->resident[i].dev;
            ctx->resident.erase(ctx->resident.begin() + i);
        }
        else ++i;
    }
    vb_tls_stream = saved;
}

void vb_unpin_genomes(const vb_genomes *g)
{
    if (g->pinned) { cudaHostUnregister((void *)g->bases.data()); g->pinned = false; }
}

thread_local vb_arena *vb_tls_arena = nullptr;
thread_local bool vb_tls_pool_alloc = false;

uint64_t vb_device_available(vb_ctx *ctx)
{
    size_t free_b = 0, total_b = 0;
    VB_CUDA(cudaMemGetInfo(&free_b, &total_b));
    cudaMemPool_t pool;
    uint64_t reserved = 0, used = 0;
    if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess) {
        cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved);
        cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used);
    }
    return (uint64_t)free_b + (reserved > used ? reserved - used : 0);
}

// ---------------------------------------------------------------------------------------------------------------
// arena
// ---------------------------------------------------------------------------------------------------------------
static size_t arena_round(size_t b) { return (b + 511) & ~(size_t)511; }

void *vb_arena::alloc(size_t bytes)
{
    bytes = arena_round(bytes);
    if (slabs.empty() || slabs.back().off + bytes > slabs.back().cap) {
        // grow: a new slab at least as large as everything so far (only during warm-up; reset() merges the slabs)
        size_t total = 0;
        for (auto &s : slabs) total += s.cap;
        size_t cap = std::max(bytes, std::max<size_t>(total, (size_t)256 << 20));
        char *base = nullptr;
        cudaError_t e = cudaMalloc((void **)&base, cap);
        if (e != cudaSuccess) {
            cudaGetLastError();
            cap = bytes;
            e = cudaMalloc((void **)&base, cap);
            if (e != cudaSuccess) {
                cudaGetLastError();
                throw vb_error(VB_ERR_MEM, "cudaMalloc of " + std::to_string(cap) + " bytes failed: " + cudaGetErrorString(e));
            }
        }
        slabs.push_back({base, cap, 0});
    }
    Slab &s = slabs.back();
    void *p = s.base + s.off;
    s.off += bytes;
    live += bytes;
    peak = std::max(peak, live);
    return p;
}

void vb_arena::pop(void *p, size_t bytes)
{
    bytes = arena_round(bytes);
    live -= std::min(live, bytes);
    for (size_t i = slabs.size(); i-- > 0;) {
        Slab &s = slabs[i];
        if ((char *)p + bytes == s.base + s.off) { s.off -= bytes; return; }     // top of this slab's stack
        if (s.off != 0) return;                                                  // not LIFO: space returns at reset()
    }
}

void vb_arena::reset(cudaStream_t st)
{
    if (slabs.size() > 1) {
        cudaStreamSynchronize(st);
        for (auto &s : slabs) cudaFree(s.base);
        slabs.clear();
        size_t want = std::max(arena_round(peak + peak / 8), (size_t)256 << 20);
        char *base = nullptr;
        if (cudaMalloc((void **)&base, want) == cudaSuccess) slabs.push_back({base, want, 0});
        else cudaGetLastError();
    }
    for (auto &s : slabs) s.off = 0;
    live = 0;
}

void vb_arena::destroy()
{
    for (auto &s : slabs) cudaFree(s.base);
    slabs.clear();
}

void *vb_pinned(vb_ctx *ctx, size_t bytes)
{
    if (bytes <= ctx->pin_cap) return ctx->pin_buf;
    if (ctx->pin_buf) { cudaFreeHost(ctx->pin_buf); ctx->pin_buf = nullptr; ctx->pin_cap = 0; }
    const size_t want = std::max(bytes + bytes / 4, (size_t)1 << 20);
    void *q = nullptr;
    if (cudaHostAlloc(&q, want, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        throw vb_error(VB_ERR_MEM, "cudaHostAlloc of " + std::to_string(want) + " bytes failed");
    }
    ctx->pin_buf = q; ctx->pin_cap = want;
    return q;
}
