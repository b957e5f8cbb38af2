// One genome set over the GPUs of one node: vb_shard_* (include/vclust_b200.h).  No reference counterpart -- kmer-db and
// lz-ani are single-process; the analogue of the split is the tiling of all2all-parts (console_all2all_parts.cpp:143-331)
// and lz-ani's per-reference work queue (lz_matcher.cpp:190-270).  The collectives are callbacks supplied by the host
// (torch.distributed over NCCL); everything they move is device memory owned by this library.
#include <algorithm>
#include <chrono>
#include <memory>
#include <numeric>

#include "dev_util.cuh"
#include "radix_sort.cuh"

void vb_make_resident_impl(vb_ctx *ctx, const vb_genomes *g, uint32_t min_pad, uint64_t force_slots = 0);
void vb_evict_impl(vb_ctx *ctx, const vb_genomes *g);
void vb_enter_call(vb_ctx *ctx);
vb_align_out *vb_align_out_alloc_impl(uint64_t total, uint32_t n);

struct vb_shard {
    vb_ctx *ctx = nullptr;
    vb_comm comm{};
    const vb_genomes *meta = nullptr, *local = nullptr;
    uint32_t first_id = 0;
    int mrd = 40;
    uint64_t block_slots = 0;            // slots of every rank's (padded) local store
    DevGenomes all;                      // rec / gofs / glen of ALL genomes (rec all-gathered); no seq2 / inv_kdb / tile map
    double est_kmers_all = 0;
    const vb_peer_xbuf *xbuf = nullptr;  // all-to-all #1 as peer copies (null: NCCL all-to-all); owned by the context
};

void vb_peer_xbuf_free(vb_peer_xbuf *x)
{
    if (!x) return;
    for (size_t r = 0; r < x->peer.size(); ++r)
        if (x->peer[r] && x->peer[r] != x->local) cudaIpcCloseMemHandle(x->peer[r]);
    if (x->local) cudaFree(x->local);
    delete x;
}

namespace {

void comm_check(int rc, const char *what)
{
    if (rc != 0) throw vb_error(VB_ERR_INTERNAL, std::string("collective failed: ") + what);
}

__global__ void iota_kernel(uint32_t *p, uint64_t n)
{
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = (uint32_t)i;
}
__global__ void fill_u64_kernel(uint64_t *p, uint64_t lo, uint64_t hi, uint64_t v)
{
    for (uint64_t i = lo + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < hi; i += (uint64_t)gridDim.x * blockDim.x) p[i] = v;
}
// out[i] = in[perm[i]] for records of three ints
__global__ void gather3_kernel(const int32_t *__restrict__ in, const uint32_t *__restrict__ perm, uint64_t n, int32_t *__restrict__ out)
{
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t j = perm[i];
        out[3 * i] = in[3 * j]; out[3 * i + 1] = in[3 * j + 1]; out[3 * i + 2] = in[3 * j + 2];
    }
}

// align results as 24-byte records {key, 3 ints, pad}: one all-to-all instead of two
struct ResRec { uint64_t key; int32_t st[3]; uint32_t pad; };
__global__ void pack_res_kernel(const uint64_t *__restrict__ keys, const int32_t *__restrict__ stats, uint64_t n, ResRec *__restrict__ rec)
{
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        rec[i] = {keys[i], {stats[3 * i], stats[3 * i + 1], stats[3 * i + 2]}, 0u};
}
__global__ void unpack_res_kernel(const ResRec *__restrict__ rec, uint64_t n, uint64_t *__restrict__ keys, int32_t *__restrict__ stats)
{
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const ResRec r = rec[i];
        keys[i] = r.key; stats[3 * i] = r.st[0]; stats[3 * i + 1] = r.st[1]; stats[3 * i + 2] = r.st[2];
    }
}

int grid_for(uint64_t n) { return (int)std::max<uint64_t>(1, std::min<uint64_t>((n + 255) / 256, 148 * 16)); }

#define VB_GUARD_BEGIN try {
#define VB_GUARD_END                                                        \
    }                                                                       \
    catch (const vb_error &e) { vb_set_error(e.what()); return e.code; }    \
    catch (const std::bad_alloc &) { vb_set_error("out of host memory"); return VB_ERR_MEM; } \
    catch (const std::exception &e) { vb_set_error(e.what()); return VB_ERR_INTERNAL; }       \
    return VB_OK;

}  // namespace

extern "C" {

int vb_shard_create(vb_ctx *ctx, const vb_comm *comm, const vb_genomes *meta, const vb_genomes *local, uint32_t first_id, int mrd,
                    vb_shard **out)
{
    VB_GUARD_BEGIN
    if (!ctx || !comm || !meta || !local || !out || !comm->all_to_all || !comm->all_gather || !comm->all_reduce_sum_u32)
        throw vb_error(VB_ERR_ARG, "vb_shard_create: bad arguments");
    if (comm->world < 1 || comm->rank < 0 || comm->rank >= comm->world) throw vb_error(VB_ERR_ARG, "vb_shard_create: bad rank / world");
    const uint32_t n = meta->count(), world = (uint32_t)comm->world;
    if ((uint64_t)first_id + local->count() > n) throw vb_error(VB_ERR_ARG, "vb_shard_create: the local block lies outside the set");
    for (uint32_t i = 0; i < local->count(); ++i)
        if (local->length(i) != meta->length(first_id + i)) throw vb_error(VB_ERR_ARG, "vb_shard_create: local genome lengths differ from the set's");
    vb_enter_call(ctx);
    cudaStream_t st = (cudaStream_t)ctx->stream;
    auto sh = std::make_unique<vb_shard>();
    sh->ctx = ctx; sh->comm = *comm; sh->meta = meta; sh->local = local; sh->first_id = first_id; sh->mrd = mrd;
    const uint32_t min_pad = std::max<uint32_t>((uint32_t)std::max(mrd, 0) + 128u, VB_STORE_PAD);

    // the blocks of all ranks: (first id, count), gathered
    DevBuf<uint32_t> d_mine(2), d_blocks(2 * (size_t)world);
    const uint32_t mine[2] = {first_id, local->count()};
    std::vector<uint32_t> blocks(2 * (size_t)world);
    VB_CUDA(cudaMemcpyAsync(d_mine.p, mine, sizeof(mine), cudaMemcpyHostToDevice, st));
    comm_check(comm->all_gather(comm->user, d_mine.p, d_blocks.p, sizeof(mine)), "all_gather(blocks)");
    VB_CUDA(cudaMemcpyAsync(blocks.data(), d_blocks.p, sizeof(uint32_t) * blocks.size(), cudaMemcpyDeviceToHost, st));
    VB_CUDA(cudaStreamSynchronize(st));
    uint32_t expect = 0;
    for (uint32_t r = 0; r < world; ++r) {
        if (blocks[2 * r] != expect) throw vb_error(VB_ERR_ARG, "vb_shard_create: the ranks' blocks must be contiguous and in rank order");
        expect += blocks[2 * r + 1];
    }
    if (expect != n) throw vb_error(VB_ERR_ARG, "vb_shard_create: the ranks' blocks do not cover the set");
    // store layout of every block (the formula of vb_upload_genomes), all blocks padded to the largest
    std::vector<uint64_t> gofs(n);
    std::vector<uint32_t> glen(n);
    uint64_t max_slots = 128;
    double total_len = 0;
    for (uint32_t r = 0; r < world; ++r) {
        uint64_t slots = 0;
        for (uint32_t g = blocks[2 * r]; g < blocks[2 * r] + blocks[2 * r + 1]; ++g) {
            const uint64_t len = meta->length(g);
            if (len > 0x7fff0000ULL) throw vb_error(VB_ERR_ARG, "genome longer than 2^31 bases: " + meta->names[g]);
            gofs[g] = slots; glen[g] = (uint32_t)len;
            slots += ((len + min_pad + 127) / 128) * 128;
            total_len += (double)len;
        }
        max_slots = std::max(max_slots, slots + 128);
    }
    sh->block_slots = max_slots;
    sh->est_kmers_all = total_len;
    for (uint32_t r = 0; r < world; ++r)
        for (uint32_t g = blocks[2 * r]; g < blocks[2 * r] + blocks[2 * r + 1]; ++g) gofs[g] += (uint64_t)r * max_slots;

    // local block: packed store of exactly max_slots slots, resident; then the align-stage records of all blocks
    vb_make_resident_impl(ctx, local, min_pad, max_slots);
    const DevGenomes &loc = vb_get_dev_genomes(ctx, local, min_pad);
    {
        const bool saved = vb_tls_pool_alloc;
        vb_tls_pool_alloc = true;
        try {
            sh->all.rec.alloc((size_t)world * (max_slots / 32) + 8);
            sh->all.gofs.alloc(std::max<uint32_t>(n, 1));
            sh->all.glen.alloc(std::max<uint32_t>(n, 1));
        } catch (...) { vb_tls_pool_alloc = saved; throw; }
        vb_tls_pool_alloc = saved;
    }
    sh->all.n = n;
    sh->all.total_slots = (uint64_t)world * max_slots;
    sh->all.min_pad = min_pad;
    sh->all.h_gofs = gofs; sh->all.h_glen = glen;
    VB_CUDA(cudaMemsetAsync(sh->all.rec.p + (size_t)world * (max_slots / 32), 0xff, 8 * sizeof(uint4), st));
    comm_check(comm->all_gather(comm->user, loc.rec.p, sh->all.rec.p, (max_slots / 32) * sizeof(uint4)), "all_gather(packed genomes)");
    if (n) {
        VB_CUDA(cudaMemcpyAsync(sh->all.gofs.p, gofs.data(), sizeof(uint64_t) * n, cudaMemcpyHostToDevice, st));
        VB_CUDA(cudaMemcpyAsync(sh->all.glen.p, glen.data(), sizeof(uint32_t) * n, cudaMemcpyHostToDevice, st));
    }
    VB_CUDA(cudaStreamSynchronize(st));
    ctx->set_timing("shard.block_slots", (double)max_slots);
    // Receive buffers for the tuple exchange, exported to all ranks (CUDA IPC): one pass moves at most ~10^9 tuples in all
    // (the prefilter's pass planner), 12 bytes each, plus the bucket histograms.  Every rank tries; the path is used only
    // if it works everywhere (VB_SHARD_NO_PEER=1 forces the NCCL all-to-all).
    if (world > 1) {
        const double per_rank = std::min(total_len, 1.0e9) / world;
        const uint64_t cap = ((uint64_t)(12.0 * 1.4 * per_rank) + (64ull << 20) + 4095) / 4096 * 4096;
        // (mapping costs ~0.2 s: the buffers are kept by the context and reused while they are large enough -- every rank
        // sees the same sizes, so every rank decides alike)
        if (ctx->xbuf && (ctx->xbuf->cap < cap || ctx->xbuf->peer.size() != world)) { vb_peer_xbuf_free(ctx->xbuf); ctx->xbuf = nullptr; }
        if (!ctx->xbuf && !getenv("VB_SHARD_NO_PEER")) {
            auto *x = new vb_peer_xbuf();
            uint32_t ok = 1u;
            cudaIpcMemHandle_t mine_h;
            memset(&mine_h, 0, sizeof(mine_h));
            if (cudaMalloc(&x->local, cap) != cudaSuccess) { cudaGetLastError(); x->local = nullptr; ok = 0; }
            if (ok && cudaIpcGetMemHandle(&mine_h, x->local) != cudaSuccess) { cudaGetLastError(); ok = 0; }
            static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
            DevBuf<uint32_t> d_h(17), d_all(17 * (size_t)world);
            uint32_t send[17];
            memcpy(send, &mine_h, 64);
            send[16] = ok;
            std::vector<uint32_t> all(17 * (size_t)world);
            VB_CUDA(cudaMemcpyAsync(d_h.p, send, sizeof(send), cudaMemcpyHostToDevice, st));
            comm_check(comm->all_gather(comm->user, d_h.p, d_all.p, sizeof(send)), "all_gather(IPC handles)");
            VB_CUDA(cudaMemcpyAsync(all.data(), d_all.p, sizeof(uint32_t) * all.size(), cudaMemcpyDeviceToHost, st));
            VB_CUDA(cudaStreamSynchronize(st));
            bool everyone = true;
            for (uint32_t r = 0; r < world; ++r) everyone = everyone && all[17 * r + 16] == 1;
            x->peer.assign(world, nullptr);
            if (everyone) {
                for (uint32_t r = 0; r < world && ok; ++r) {
                    if (r == (uint32_t)comm->rank) { x->peer[r] = x->local; continue; }
                    cudaIpcMemHandle_t h;
                    memcpy(&h, &all[17 * r], 64);
                    if (cudaIpcOpenMemHandle(&x->peer[r], h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); x->peer[r] = nullptr; ok = 0; }
                }
            } else ok = 0;
            // did every rank map every buffer?
            DevBuf<uint32_t> d_ok(1);
            VB_CUDA(cudaMemcpyAsync(d_ok.p, &ok, sizeof(ok), cudaMemcpyHostToDevice, st));
            comm_check(comm->all_reduce_sum_u32(comm->user, d_ok.p, 1), "all_reduce(peer mapping)");
            uint32_t n_ok = 0;
            VB_CUDA(cudaMemcpyAsync(&n_ok, d_ok.p, sizeof(n_ok), cudaMemcpyDeviceToHost, st));
            VB_CUDA(cudaStreamSynchronize(st));
            if (n_ok == world) { x->cap = cap; ctx->xbuf = x; }
            else vb_peer_xbuf_free(x);
        }
        sh->xbuf = getenv("VB_SHARD_NO_PEER") ? nullptr : ctx->xbuf;
        ctx->set_timing("shard.peer_exchange", sh->xbuf ? 1.0 : 0.0);
    }
    *out = sh.release();
    VB_GUARD_END
}

int vb_shard_prefilter(vb_shard *sh, const vb_prefilter_params *p, vb_pairs **out)
{
    VB_GUARD_BEGIN
    if (!sh || !p || !out) throw vb_error(VB_ERR_ARG, "vb_shard_prefilter: bad arguments");
    vb_enter_call(sh->ctx);
    vb_prefilter_job job;
    job.g = sh->local;
    job.gid_base = sh->first_id;
    job.n_total = sh->meta->count();
    job.est_kmers_all = sh->est_kmers_all;
    job.comm = &sh->comm;
    job.xbuf = sh->xbuf;
    job.keep_dev = true;
    vb_prefilter_run(sh->ctx, job, p, out);
    VB_GUARD_END
}

int vb_shard_align(vb_shard *sh, const vb_align_params *p, vb_align_out **out)
{
    VB_GUARD_BEGIN
    if (!sh || !p || !out) throw vb_error(VB_ERR_ARG, "vb_shard_align: bad arguments");
    vb_ctx *ctx = sh->ctx;
    if (p->mrd + 128 > (int)sh->all.min_pad) throw vb_error(VB_ERR_ARG, "vb_shard_align: --mrd larger than the one given to vb_shard_create");
    vb_enter_call(ctx);
    ctx->clear_timings("align.");
    cudaStream_t st = (cudaStream_t)ctx->stream;
    const vb_comm &cm = sh->comm;
    const uint32_t world = (uint32_t)cm.world, rank = (uint32_t)cm.rank, n = sh->meta->count();
    const DevPairs *dp = ctx->dev_pairs;
    AlignFastOut fo;
    if (!vb_align_fast(ctx, sh->meta, sh->all, p, dp ? dp->keys.p : nullptr, dp ? dp->ani.p : nullptr, dp ? dp->n : 0, false, world, rank, fo))
        throw vb_error(VB_ERR_MEM, "the reference indexes of this rank's genomes do not fit device memory");
    EventTimer t_g(st);
    t_g.start();
    // gather on rank 0: sizes, then keys and statistics; rank 0 merges the ranks' sorted runs with one sort
    DevBuf<unsigned long long> d_cnt(1), d_all(world);
    const unsigned long long my_n = fo.n;
    VB_CUDA(cudaMemcpyAsync(d_cnt.p, &my_n, sizeof(my_n), cudaMemcpyHostToDevice, st));
    comm_check(cm.all_gather(cm.user, d_cnt.p, d_all.p, sizeof(unsigned long long)), "all_gather(result sizes)");
    std::vector<unsigned long long> cnts(world);
    VB_CUDA(cudaMemcpyAsync(cnts.data(), d_all.p, sizeof(unsigned long long) * world, cudaMemcpyDeviceToHost, st));
    VB_CUDA(cudaStreamSynchronize(st));
    uint64_t total = 0;
    std::vector<uint64_t> scv(world, 0), rcv(world, 0);
    scv[0] = fo.n;
    if (rank == 0) for (uint32_t s = 0; s < world; ++s) { rcv[s] = cnts[s]; total += cnts[s]; }
    if (total >= (1ULL << 32) - rsort::TILE) throw vb_error(VB_ERR_ARG, "more than 2^32 directed pairs");
    const uint64_t n_pad = (total + rsort::TILE - 1) / rsort::TILE * rsort::TILE;
    DevBuf<uint64_t> ka(n_pad + 1), kb(n_pad + 1);
    DevBuf<uint32_t> va(n_pad + 1), vbuf(n_pad + 1);
    DevBuf<int32_t> st_all(3 * total + 3), st_sorted(3 * total + 3);
    DevBuf<ResRec> send_rec(fo.n + 1), recv_rec(total + 1);
    static_assert(sizeof(ResRec) == 24, "result record size");
    if (fo.n) { pack_res_kernel<<<grid_for(fo.n), 256, 0, st>>>(fo.keys.p, fo.stats.p, fo.n, send_rec.p); VB_LAUNCH_CHECK(ctx); }
    comm_check(cm.all_to_all(cm.user, send_rec.p, scv.data(), recv_rec.p, rcv.data(), 24), "all_to_all(results)");
    if (total) { unpack_res_kernel<<<grid_for(total), 256, 0, st>>>(recv_rec.p, total, ka.p, st_all.p); VB_LAUNCH_CHECK(ctx); }
    vb_align_out *res = nullptr;
    if (rank == 0) {
        char *pin = (char *)vb_pinned(ctx, 20 * total + 16);
        const uint64_t *keys = (const uint64_t *)pin;
        const int32_t *stats = (const int32_t *)(pin + 8 * total);
        if (total) {
            rsort::Workspace ws;
            iota_kernel<<<grid_for(n_pad), 256, 0, st>>>(va.p, n_pad);
            VB_LAUNCH_CHECK(ctx);
            if (n_pad > total) { fill_u64_kernel<<<grid_for(n_pad - total), 256, 0, st>>>(ka.p, total, n_pad, ~0ULL); VB_LAUNCH_CHECK(ctx); }
            const bool in_b = rsort::sort_kv<8>(ctx, ka.p, va.p, kb.p, vbuf.p, n_pad, 2 * fo.gbits, ws);
            gather3_kernel<<<grid_for(total), 256, 0, st>>>(st_all.p, in_b ? vbuf.p : va.p, total, st_sorted.p);
            VB_LAUNCH_CHECK(ctx);
            VB_CUDA(cudaMemcpyAsync(pin, in_b ? kb.p : ka.p, sizeof(uint64_t) * total, cudaMemcpyDeviceToHost, st));
            VB_CUDA(cudaMemcpyAsync(pin + 8 * total, st_sorted.p, sizeof(int32_t) * 3 * total, cudaMemcpyDeviceToHost, st));
        }
        t_g.stop();
        res = vb_align_out_alloc_impl(total, n);            // (host allocation overlaps the transfer)
        VB_CUDA(cudaStreamSynchronize(st));
        const std::vector<uint32_t> &order = vb_lz_order(sh->meta);
        std::copy(order.begin(), order.end(), res->order);
        const uint64_t qmask = (1ULL << fo.gbits) - 1;
        const int gbits = fo.gbits;
        vb_parallel_for(total, 65536, 16, [&](uint64_t lo, uint64_t hi) {
            for (uint64_t i = lo; i < hi; ++i) {
                res->ref[i] = (uint32_t)(keys[i] >> gbits); res->qry[i] = (uint32_t)(keys[i] & qmask);
                res->sym_in_matches[i] = stats[3 * i]; res->sym_in_literals[i] = stats[3 * i + 1]; res->no_components[i] = stats[3 * i + 2];
            }
        });
    } else {
        t_g.stop();
        VB_CUDA(cudaStreamSynchronize(st));
        res = vb_align_out_alloc_impl(0, n);
        const std::vector<uint32_t> &order = vb_lz_order(sh->meta);
        std::copy(order.begin(), order.end(), res->order);
    }
    ctx->set_timing("align.gather_ms", t_g.ms());
    *out = res;
    VB_GUARD_END
}

void vb_shard_destroy(vb_shard *sh)
{
    if (!sh) return;
    cudaSetDevice(sh->ctx->device);
    cudaStream_t saved = vb_tls_stream;
    vb_tls_stream = (cudaStream_t)sh->ctx->stream;
    vb_evict_impl(sh->ctx, sh->local);
    delete sh;                                    // (its pool buffers are freed in stream order)
    vb_tls_stream = saved;
}

// Exercises the callbacks on small device buffers: all_gather of rank-stamped words, all_to_all with uneven counts
// (rank r sends (r + p + 1) records of 12 bytes to peer p), all_reduce of ones.
int vb_comm_selftest(vb_ctx *ctx, const vb_comm *comm)
{
    VB_GUARD_BEGIN
    if (!ctx || !comm) throw vb_error(VB_ERR_ARG, "vb_comm_selftest: bad arguments");
    vb_enter_call(ctx);
    cudaStream_t st = (cudaStream_t)ctx->stream;
    const uint32_t world = (uint32_t)comm->world, rank = (uint32_t)comm->rank;
    // all_gather
    DevBuf<uint32_t> g_in(4), g_out(4 * (size_t)world);
    const uint32_t stamp[4] = {rank, rank * 7 + 1, 0xabcd0000u + rank, 42};
    VB_CUDA(cudaMemcpyAsync(g_in.p, stamp, sizeof(stamp), cudaMemcpyHostToDevice, st));
    comm_check(comm->all_gather(comm->user, g_in.p, g_out.p, sizeof(stamp)), "all_gather");
    std::vector<uint32_t> got(4 * (size_t)world);
    VB_CUDA(cudaMemcpyAsync(got.data(), g_out.p, sizeof(uint32_t) * got.size(), cudaMemcpyDeviceToHost, st));
    VB_CUDA(cudaStreamSynchronize(st));
    for (uint32_t r = 0; r < world; ++r)
        if (got[4 * r] != r || got[4 * r + 1] != r * 7 + 1 || got[4 * r + 2] != 0xabcd0000u + r || got[4 * r + 3] != 42)
            throw vb_error(VB_ERR_INTERNAL, "all_gather returned wrong data");
    // all_to_all, 12-byte records {source, destination, index}
    std::vector<uint64_t> sc(world), rc(world);
    uint64_t ns = 0, nr = 0;
    for (uint32_t pr = 0; pr < world; ++pr) { sc[pr] = rank + pr + 1; rc[pr] = pr + rank + 1; ns += sc[pr]; nr += rc[pr]; }
    std::vector<uint32_t> send(3 * ns), recv(3 * nr);
    uint64_t at = 0;
    for (uint32_t pr = 0; pr < world; ++pr)
        for (uint64_t i = 0; i < sc[pr]; ++i, ++at) { send[3 * at] = rank; send[3 * at + 1] = pr; send[3 * at + 2] = (uint32_t)i; }
    DevBuf<uint32_t> d_send(3 * ns), d_recv(3 * nr);
    VB_CUDA(cudaMemcpyAsync(d_send.p, send.data(), sizeof(uint32_t) * send.size(), cudaMemcpyHostToDevice, st));
    comm_check(comm->all_to_all(comm->user, d_send.p, sc.data(), d_recv.p, rc.data(), 12), "all_to_all");
    VB_CUDA(cudaMemcpyAsync(recv.data(), d_recv.p, sizeof(uint32_t) * recv.size(), cudaMemcpyDeviceToHost, st));
    VB_CUDA(cudaStreamSynchronize(st));
    at = 0;
    for (uint32_t pr = 0; pr < world; ++pr)
        for (uint64_t i = 0; i < rc[pr]; ++i, ++at)
            if (recv[3 * at] != pr || recv[3 * at + 1] != rank || recv[3 * at + 2] != (uint32_t)i)
                throw vb_error(VB_ERR_INTERNAL, "all_to_all returned wrong data");
    // all_reduce
    std::vector<uint32_t> ones(1000);
    for (uint32_t i = 0; i < 1000; ++i) ones[i] = i + rank;
    DevBuf<uint32_t> d_r(1000);
    VB_CUDA(cudaMemcpyAsync(d_r.p, ones.data(), sizeof(uint32_t) * 1000, cudaMemcpyHostToDevice, st));
    comm_check(comm->all_reduce_sum_u32(comm->user, d_r.p, 1000), "all_reduce");
    VB_CUDA(cudaMemcpyAsync(ones.data(), d_r.p, sizeof(uint32_t) * 1000, cudaMemcpyDeviceToHost, st));
    VB_CUDA(cudaStreamSynchronize(st));
    for (uint32_t i = 0; i < 1000; ++i)
        if (ones[i] != i * world + world * (world - 1) / 2) throw vb_error(VB_ERR_INTERNAL, "all_reduce returned wrong data");
    VB_GUARD_END
}

}  // extern "C"
