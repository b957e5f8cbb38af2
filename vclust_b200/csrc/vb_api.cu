// C ABI of libvclust_b200.so (see include/vclust_b200.h).  Everything here is glue: argument checks, exception ->
// status translation, and the host-side bookkeeping LZ-ANI does around its matching loop (genome re-ordering,
// filter symmetrisation: seq_reservoir.cpp:215-251, filter.cpp:80-81,253-345, lz_matcher.cpp:172-277).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <numeric>

#include "dev_util.cuh"

vb_genomes *vb_genomes_load_impl(const char *const *paths, int n_paths, int multisample, vb_fasta_flavor flavor, int sep_len);
vb_genomes *vb_genomes_from_memory_impl(const char *const *names, const char *const *seqs, const uint64_t *lens, uint32_t n);
void vb_write_filter_impl(const vb_genomes *g, const vb_pairs *pr, const char *path);
vb_pairs *vb_read_filter_impl(const char *path, double thr, const vb_genomes *g);
void vb_pairs_free_impl(vb_pairs *p);
void vb_write_ani_impl(const vb_genomes *g, const vb_align_out *res, const char *ani_path, const char *ids_path,
                       const char *const *columns, int n_columns, const double out_filters[5]);

void vb_make_resident_impl(vb_ctx *ctx, const vb_genomes *g, uint32_t min_pad, uint64_t force_slots = 0);
void vb_peer_xbuf_free(vb_peer_xbuf *x);
void vb_evict_impl(vb_ctx *ctx, const vb_genomes *g);
void vb_unpin_genomes(const vb_genomes *g);

thread_local cudaStream_t vb_tls_stream = nullptr;
uint64_t vb_next_uid()
{
    static std::atomic<uint64_t> next{1};
    return next.fetch_add(1);
}
static thread_local std::string g_last_error;
void vb_set_error(const std::string &msg) { g_last_error = msg; }

#define VB_GUARD_BEGIN try {
#define VB_GUARD_END                                                        \
    }                                                                       \
    catch (const vb_error &e) { vb_set_error(e.what()); return e.code; }    \
    catch (const std::bad_alloc &) { vb_set_error("out of host memory"); return VB_ERR_MEM; } \
    catch (const std::exception &e) { vb_set_error(e.what()); return VB_ERR_INTERNAL; }       \
    return VB_OK;

vb_pairs *vb_pairs_alloc(uint64_t n, uint32_t n_genomes);

// every device-touching entry point: select the device, bind this thread to the context's stream and arena, and
// start with an empty arena (all temporaries of the previous call are dead by now)
static void vb_enter(vb_ctx *ctx)
{
    VB_CUDA(cudaSetDevice(ctx->device));
    vb_tls_stream = (cudaStream_t)ctx->stream;
    vb_tls_arena = ctx->arena;
    ctx->arena->reset((cudaStream_t)ctx->stream);
}

void vb_enter_call(vb_ctx *ctx) { vb_enter(ctx); }

// LZ-ANI order: length descending, then name ascending (stable) -- seq_reservoir.cpp:229-236.  Computed once per set.
const std::vector<uint32_t> &vb_lz_order(const vb_genomes *g)
{
    if (g->lz_order.size() != g->count() || g->lz_rank.size() != g->count()) {
        std::vector<uint32_t> order(g->count());
        std::iota(order.begin(), order.end(), 0u);
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
            uint32_t la = (uint32_t)g->length(a) - 2u, lb = (uint32_t)g->length(b) - 2u;   // len - 2*no_parts, unsigned (sic)
            if (la != lb) return la > lb;
            return g->names[a] < g->names[b];
        });
        std::vector<uint32_t> rank(g->count());
        for (uint32_t i = 0; i < g->count(); ++i) rank[order[i]] = i;
        g->lz_order.swap(order);
        g->lz_rank.swap(rank);
    }
    return g->lz_order;
}
const std::vector<uint32_t> &vb_lz_rank(const vb_genomes *g) { vb_lz_order(g); return g->lz_rank; }

static vb_align_out *vb_align_out_alloc(uint64_t total, uint32_t n)
{
    if (total >= (1ULL << 32)) throw vb_error(VB_ERR_ARG, "more than 2^32 directed pairs in one call");
    auto *res = (vb_align_out *)calloc(1, sizeof(vb_align_out));
    if (!res) throw vb_error(VB_ERR_MEM, "out of host memory");
    res->n = total;
    res->n_genomes = n;
    res->ref = (uint32_t *)malloc(sizeof(uint32_t) * std::max<uint64_t>(total, 1));
    res->qry = (uint32_t *)malloc(sizeof(uint32_t) * std::max<uint64_t>(total, 1));
    res->sym_in_matches = (int32_t *)calloc(std::max<uint64_t>(total, 1), sizeof(int32_t));
    res->sym_in_literals = (int32_t *)calloc(std::max<uint64_t>(total, 1), sizeof(int32_t));
    res->no_components = (int32_t *)calloc(std::max<uint64_t>(total, 1), sizeof(int32_t));
    res->order = (uint32_t *)malloc(sizeof(uint32_t) * std::max<uint32_t>(n, 1));
    if (!res->ref || !res->qry || !res->sym_in_matches || !res->sym_in_literals || !res->no_components || !res->order) {
        vb_align_out_free(res);
        throw vb_error(VB_ERR_MEM, "out of host memory for " + std::to_string(total) + " directed pairs");
    }
    return res;
}

vb_align_out *vb_align_out_alloc_impl(uint64_t total, uint32_t n) { return vb_align_out_alloc(total, n); }

extern "C" {

int vb_version(char *buf, size_t n)
{
    static const char v[] = "vclust-b200 0.1.0 (prefilter: Kmer-db 2.3.1 semantics; align: LZ-ANI 1.2.3 semantics; sm_100a)";
    if (!buf || n == 0) return VB_ERR_ARG;
    snprintf(buf, n, "%s", v);
    return VB_OK;
}

int vb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char *vb_last_error(void) { return g_last_error.c_str(); }

int vb_ctx_create(int device, vb_ctx **out) { return vb_ctx_create_on_stream(device, nullptr, out); }

int vb_ctx_create_on_stream(int device, void *cuda_stream, vb_ctx **out)
{
    VB_GUARD_BEGIN
    if (!out) throw vb_error(VB_ERR_ARG, "vb_ctx_create: null out pointer");
    int n = vb_device_count();
    if (n <= 0) throw vb_error(VB_ERR_CUDA, "no CUDA device available: libvclust_b200 has no CPU fallback");
    if (device < 0 || device >= n) throw vb_error(VB_ERR_ARG, "vb_ctx_create: device index out of range");
    VB_CUDA(cudaSetDevice(device));
    auto *ctx = new vb_ctx();
    ctx->device = device;
    if (cuda_stream) { ctx->stream = cuda_stream; ctx->owns_stream = false; }
    else {
        cudaStream_t st;
        VB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        ctx->stream = (void *)st;
    }
    cudaStream_t cs;
    VB_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    ctx->copy_stream = (void *)cs;
    for (auto &e : ctx->copy_events) { cudaEvent_t ev; VB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)); e = (void *)ev; }
    ctx->arena = new vb_arena();
    {   // packed genome stores live in the device's stream-ordered pool: never hand freed blocks back to the driver
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t thr = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
        cudaGetLastError();
    }
    size_t free_b = 0, total_b = 0;
    VB_CUDA(cudaMemGetInfo(&free_b, &total_b));
    ctx->mem_total = total_b;
    *out = ctx;
    VB_GUARD_END
}

void vb_ctx_destroy(vb_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    vb_tls_stream = (cudaStream_t)ctx->stream;
    vb_tls_arena = nullptr;
    vb_drop_dev_pairs(ctx);
    vb_evict_impl(ctx, nullptr);
    cudaStreamSynchronize((cudaStream_t)ctx->stream);
    if (ctx->arena) { ctx->arena->destroy(); delete ctx->arena; }
    if (ctx->pin_buf) cudaFreeHost(ctx->pin_buf);
    if (ctx->xbuf) vb_peer_xbuf_free(ctx->xbuf);
    for (auto &e : ctx->events) if (e) cudaEventDestroy((cudaEvent_t)e);
    for (auto &e : ctx->copy_events) if (e) cudaEventDestroy((cudaEvent_t)e);
    if (ctx->copy_stream) { cudaStreamSynchronize((cudaStream_t)ctx->copy_stream); cudaStreamDestroy((cudaStream_t)ctx->copy_stream); }
    if (ctx->stream && ctx->owns_stream) cudaStreamDestroy((cudaStream_t)ctx->stream);
    delete ctx;
}

int vb_ctx_mark(vb_ctx *ctx, int slot)
{
    VB_GUARD_BEGIN
    if (!ctx || slot < 0 || slot >= 8) throw vb_error(VB_ERR_ARG, "vb_ctx_mark: slot must be 0..7");
    VB_CUDA(cudaSetDevice(ctx->device));
    if (!ctx->events[slot]) { cudaEvent_t e; VB_CUDA(cudaEventCreate(&e)); ctx->events[slot] = (void *)e; }
    VB_CUDA(cudaEventRecord((cudaEvent_t)ctx->events[slot], (cudaStream_t)ctx->stream));
    VB_GUARD_END
}

int vb_ctx_elapsed_ms(vb_ctx *ctx, int slot_a, int slot_b, double *ms)
{
    VB_GUARD_BEGIN
    if (!ctx || !ms || slot_a < 0 || slot_a >= 8 || slot_b < 0 || slot_b >= 8 || !ctx->events[slot_a] || !ctx->events[slot_b])
        throw vb_error(VB_ERR_ARG, "vb_ctx_elapsed_ms: unknown slot");
    VB_CUDA(cudaEventSynchronize((cudaEvent_t)ctx->events[slot_b]));
    float t = 0;
    VB_CUDA(cudaEventElapsedTime(&t, (cudaEvent_t)ctx->events[slot_a], (cudaEvent_t)ctx->events[slot_b]));
    *ms = t;
    VB_GUARD_END
}

int vb_genomes_make_resident(vb_ctx *ctx, const vb_genomes *g, vb_fasta_flavor rule, int mrd)
{
    VB_GUARD_BEGIN
    if (!ctx || !g) throw vb_error(VB_ERR_ARG, "vb_genomes_make_resident: bad arguments");
    vb_enter(ctx);
    (void)rule;     // one packed store serves both stages (two validity planes)
    vb_make_resident_impl(ctx, g, (uint32_t)std::max(mrd, 0) + 128u);
    VB_GUARD_END
}

int vb_genomes_evict(vb_ctx *ctx, const vb_genomes *g)
{
    VB_GUARD_BEGIN
    if (!ctx) throw vb_error(VB_ERR_ARG, "vb_genomes_evict: bad arguments");
    vb_enter(ctx);
    vb_evict_impl(ctx, g);
    VB_GUARD_END
}

int vb_ctx_timing(const vb_ctx *ctx, const char *key, double *ms)
{
    if (!ctx || !key || !ms) return VB_ERR_ARG;
    for (auto &t : ctx->timings)
        if (t.key == key) { *ms = t.ms; return VB_OK; }
    return VB_ERR_ARG;
}

uint64_t vb_ctx_launches(const vb_ctx *ctx) { return ctx ? ctx->launches : 0; }

int vb_genomes_load(const char *const *paths, int n_paths, int multisample, vb_fasta_flavor flavor, int sep_len,
                    vb_genomes **out)
{
    VB_GUARD_BEGIN
    if (!paths || n_paths <= 0 || !out) throw vb_error(VB_ERR_ARG, "vb_genomes_load: bad arguments");
    *out = vb_genomes_load_impl(paths, n_paths, multisample, flavor, sep_len);
    VB_GUARD_END
}

int vb_genomes_from_memory(const char *const *names, const char *const *seqs, const uint64_t *lens, uint32_t n,
                           vb_genomes **out)
{
    VB_GUARD_BEGIN
    if (!names || !seqs || !lens || !out) throw vb_error(VB_ERR_ARG, "vb_genomes_from_memory: bad arguments");
    *out = vb_genomes_from_memory_impl(names, seqs, lens, n);
    VB_GUARD_END
}

int vb_genomes_skeleton(const char *const *names, const uint64_t *lens, uint32_t n, vb_genomes **out)
{
    VB_GUARD_BEGIN
    if (!out || (n && (!names || !lens))) throw vb_error(VB_ERR_ARG, "vb_genomes_skeleton: bad arguments");
    auto *g = new vb_genomes();
    g->skeleton = true;
    g->names.reserve(n);
    g->offset.assign(1, 0);
    for (uint32_t i = 0; i < n; ++i) { g->names.emplace_back(names[i]); g->offset.push_back(g->offset.back() + lens[i]); }
    *out = g;
    VB_GUARD_END
}

uint32_t vb_genomes_count(const vb_genomes *g) { return g ? g->count() : 0; }
const char *vb_genomes_name(const vb_genomes *g, uint32_t i) { return (g && i < g->count()) ? g->names[i].c_str() : ""; }
uint64_t vb_genomes_length(const vb_genomes *g, uint32_t i) { return (g && i < g->count()) ? g->length(i) : 0; }
uint64_t vb_genomes_total_bases(const vb_genomes *g) { return g ? g->offset.back() : 0; }
const char *vb_genomes_sequence(const vb_genomes *g, uint32_t i) { return (g && !g->skeleton && i < g->count()) ? g->bases.data() + g->offset[i] : nullptr; }
void vb_genomes_free(vb_genomes *g) { if (g) { vb_unpin_genomes(g); delete g; } }

int vb_prefilter(vb_ctx *ctx, const vb_genomes *g, const vb_prefilter_params *p, vb_pairs **out)
{
    VB_GUARD_BEGIN
    if (!ctx || !g || !p || !out) throw vb_error(VB_ERR_ARG, "vb_prefilter: bad arguments");
    // Inputs beyond what one pass holds (~10^9 k-mer tuples) run in several passes over disjoint slices of the k-mer hash
    // space into one accumulator (prefilter.cu) -- what vclust's --batch-size is for (the reference builds partial
    // databases and runs all2all-parts); the output does not depend on it.
    vb_enter(ctx);
    vb_prefilter_job job;
    job.g = g;
    vb_prefilter_run(ctx, job, p, out);
    VB_GUARD_END
}

int vb_prefilter_partial(vb_ctx *ctx, const vb_genomes *g, const vb_prefilter_params *p, uint32_t shard_index,
                         uint32_t shard_count, vb_pairs **out)
{
    VB_GUARD_BEGIN
    if (!ctx || !g || !p || !out) throw vb_error(VB_ERR_ARG, "vb_prefilter_partial: bad arguments");
    vb_enter(ctx);
    vb_prefilter_job job;
    job.g = g;
    job.shard_index = shard_index;
    job.shard_count = shard_count;
    job.keep_dev = false;
    vb_prefilter_run(ctx, job, p, out);
    VB_GUARD_END
}

int vb_pairs_merge(const uint32_t *row, const uint32_t *col, const uint32_t *common, uint64_t n,
                   const uint32_t *total_kmers, uint32_t n_genomes, const vb_prefilter_params *p, vb_pairs **out)
{
    VB_GUARD_BEGIN
    if (!p || !out || !total_kmers || (n && (!row || !col || !common))) throw vb_error(VB_ERR_ARG, "vb_pairs_merge: bad arguments");
    std::vector<uint64_t> idx(n);
    std::iota(idx.begin(), idx.end(), 0ULL);
    std::sort(idx.begin(), idx.end(), [&](uint64_t a, uint64_t b) {
        return row[a] != row[b] ? row[a] < row[b] : col[a] < col[b];
    });
    std::vector<uint32_t> r, c, v;
    std::vector<double> a;
    for (uint64_t i = 0; i < n;) {
        uint64_t j = i;
        uint64_t sum = 0;
        while (j < n && row[idx[j]] == row[idx[i]] && col[idx[j]] == col[idx[i]]) sum += common[idx[j++]];
        uint32_t rr = row[idx[i]], cc = col[idx[i]];
        if (rr >= n_genomes || cc >= n_genomes) throw vb_error(VB_ERR_ARG, "vb_pairs_merge: id out of range");
        if (sum > 0 && sum >= (uint64_t)std::max(p->min_kmers, 0)) {           // sparse_filters.h:49-61
            double ani = vb_ani_shorter((uint32_t)sum, total_kmers[rr], total_kmers[cc], p->k);
            if (ani >= p->min_ident) { r.push_back(rr); c.push_back(cc); v.push_back((uint32_t)sum); a.push_back(ani); }
        }
        i = j;
    }
    if (p->max_seqs > 0) vb_sample_rows(n_genomes, (uint32_t)p->max_seqs, r, c, v, a);
    vb_pairs *res = vb_pairs_alloc(r.size(), n_genomes);
    for (size_t i = 0; i < r.size(); ++i) { res->row[i] = r[i]; res->col[i] = c[i]; res->common[i] = v[i]; res->ani[i] = a[i]; }
    for (uint32_t i = 0; i < n_genomes; ++i) res->total_kmers[i] = total_kmers[i];
    res->k = p->k;
    res->kmers_fraction = p->kmers_fraction;
    *out = res;
    VB_GUARD_END
}

int vb_align_out_from_pairs(const vb_genomes *g, const uint32_t *ref, const uint32_t *qry, const int32_t *stats, uint64_t n,
                            vb_align_out **out)
{
    VB_GUARD_BEGIN
    if (!g || !out || (n && (!ref || !qry || !stats))) throw vb_error(VB_ERR_ARG, "vb_align_out_from_pairs: bad arguments");
    const uint32_t ng = g->count();
    std::vector<uint32_t> order = vb_lz_order(g);
    std::vector<uint32_t> rank(ng);
    for (uint32_t i = 0; i < ng; ++i) rank[order[i]] = i;
    std::vector<uint64_t> idx(n);
    std::iota(idx.begin(), idx.end(), 0ULL);
    for (uint64_t i = 0; i < n; ++i)
        if (ref[i] >= ng || qry[i] >= ng) throw vb_error(VB_ERR_ARG, "vb_align_out_from_pairs: id out of range");
    std::stable_sort(idx.begin(), idx.end(), [&](uint64_t a, uint64_t b) {
        uint32_t ra = rank[ref[a]], rb = rank[ref[b]];
        return ra != rb ? ra < rb : rank[qry[a]] < rank[qry[b]];
    });
    vb_align_out *res = vb_align_out_alloc(n, ng);
    std::copy(order.begin(), order.end(), res->order);
    for (uint64_t o = 0; o < n; ++o) {
        uint64_t i = idx[o];
        res->ref[o] = rank[ref[i]]; res->qry[o] = rank[qry[i]];
        res->sym_in_matches[o] = stats[3 * i]; res->sym_in_literals[o] = stats[3 * i + 1]; res->no_components[o] = stats[3 * i + 2];
    }
    *out = res;
    VB_GUARD_END
}

int vb_write_filter(const vb_genomes *g, const vb_pairs *pairs, const char *path)
{
    VB_GUARD_BEGIN
    if (!g || !pairs || !path) throw vb_error(VB_ERR_ARG, "vb_write_filter: bad arguments");
    vb_write_filter_impl(g, pairs, path);
    VB_GUARD_END
}

int vb_read_filter(const char *path, double thr, const vb_genomes *g, vb_pairs **out)
{
    VB_GUARD_BEGIN
    if (!g || !path || !out) throw vb_error(VB_ERR_ARG, "vb_read_filter: bad arguments");
    *out = vb_read_filter_impl(path, thr, g);
    VB_GUARD_END
}

void vb_pairs_free(vb_pairs *p) { vb_pairs_free_impl(p); }

int vb_align_pairs(vb_ctx *ctx, const vb_genomes *g, const uint32_t *ref, const uint32_t *qry, uint64_t n,
                   const vb_align_params *p, int32_t *stats)
{
    VB_GUARD_BEGIN
    if (!ctx || !g || !p || (n && (!ref || !qry || !stats))) throw vb_error(VB_ERR_ARG, "vb_align_pairs: bad arguments");
    vb_enter(ctx);
    vb_align_pairs_impl(ctx, g, ref, qry, n, p, stats);
    VB_GUARD_END
}

// regions from the kernel's 7-int records (pair index first), sorted by pair, then like calc_regions
static vb_regions *vb_regions_build(const std::vector<int32_t> &rec, const uint32_t *ref, const uint32_t *qry, int mrd)
{
    const uint64_t n = rec.size() / 7;
    std::vector<uint64_t> idx(n);
    std::iota(idx.begin(), idx.end(), 0ULL);
    std::sort(idx.begin(), idx.end(), [&](uint64_t a, uint64_t b) {
        const int32_t *x = &rec[7 * a], *y = &rec[7 * b];
        if (x[0] != y[0]) return (uint32_t)x[0] < (uint32_t)y[0];
        const int lx = x[2] - x[1], ly = y[2] - y[1];
        if (lx != ly) return lx > ly;                                   // parser.cpp:826-833
        return x[1] < y[1];
    });
    vb_regions *r = (vb_regions *)calloc(1, sizeof(vb_regions));
    if (!r) throw vb_error(VB_ERR_MEM, "out of memory");
    r->n = n; r->mrd = mrd;
    const size_t m = std::max<uint64_t>(n, 1);
    r->ref = (uint32_t *)malloc(m * 4); r->qry = (uint32_t *)malloc(m * 4);
    r->q_start = (int32_t *)malloc(m * 4); r->q_end = (int32_t *)malloc(m * 4);
    r->r_start = (int32_t *)malloc(m * 4); r->r_end = (int32_t *)malloc(m * 4);
    r->matches = (int32_t *)malloc(m * 4); r->mismatches = (int32_t *)malloc(m * 4);
    if (!r->ref || !r->qry || !r->q_start || !r->q_end || !r->r_start || !r->r_end || !r->matches || !r->mismatches) {
        vb_regions_free(r);
        throw vb_error(VB_ERR_MEM, "out of memory");
    }
    for (uint64_t o = 0; o < n; ++o) {
        const int32_t *x = &rec[7 * idx[o]];
        r->ref[o] = ref[(uint32_t)x[0]]; r->qry[o] = qry[(uint32_t)x[0]];
        r->q_start[o] = x[1]; r->q_end[o] = x[2]; r->r_start[o] = x[3]; r->r_end[o] = x[4];
        r->matches[o] = x[5]; r->mismatches[o] = x[6];
    }
    return r;
}

void vb_regions_free(vb_regions *r)
{
    if (!r) return;
    free(r->ref); free(r->qry); free(r->q_start); free(r->q_end); free(r->r_start); free(r->r_end);
    free(r->matches); free(r->mismatches); free(r);
}

int vb_align_pairs_regions(vb_ctx *ctx, const vb_genomes *g, const uint32_t *ref, const uint32_t *qry, uint64_t n,
                           const vb_align_params *p, int32_t *stats, vb_regions **regions)
{
    VB_GUARD_BEGIN
    if (!ctx || !g || !p || !regions || (n && (!ref || !qry || !stats))) throw vb_error(VB_ERR_ARG, "vb_align_pairs_regions: bad arguments");
    vb_enter(ctx);
    const uint32_t ng = g->count();
    std::vector<uint8_t> is_ref(ng, 0);
    for (uint64_t i = 0; i < n; ++i) {
        if (ref[i] >= ng || qry[i] >= ng) throw vb_error(VB_ERR_ARG, "pair id out of range");
        is_ref[ref[i]] = 1;
    }
    std::vector<int32_t> rec;
    vb_align_job *job = vb_align_job_begin(ctx, g, p, is_ref.data());
    try {
        vb_align_job_run(job, ref, qry, n, stats, &rec);
    } catch (...) {
        vb_align_job_end(job);
        throw;
    }
    vb_align_job_end(job);
    *regions = vb_regions_build(rec, ref, qry, p->mrd);
    VB_GUARD_END
}

int vb_write_aln(const vb_genomes *g, const vb_regions *regions, const char *path, const double out_filters[5])
{
    VB_GUARD_BEGIN
    if (!g || !regions || !path) throw vb_error(VB_ERR_ARG, "vb_write_aln: bad arguments");
    vb_write_aln_impl(g, regions, path, out_filters);
    VB_GUARD_END
}

static void vb_align_common(vb_ctx *ctx, const vb_genomes *g, const vb_pairs *pairs, const vb_align_params *p, vb_align_out **out,
                            vb_regions **regions);

int vb_align(vb_ctx *ctx, const vb_genomes *g, const vb_pairs *pairs, const vb_align_params *p, vb_align_out **out)
{
    VB_GUARD_BEGIN
    if (!ctx || !g || !p || !out) throw vb_error(VB_ERR_ARG, "vb_align: bad arguments");
    vb_align_common(ctx, g, pairs, p, out, nullptr);
    VB_GUARD_END
}

int vb_align_regions(vb_ctx *ctx, const vb_genomes *g, const vb_pairs *pairs, const vb_align_params *p, vb_align_out **out,
                     vb_regions **regions)
{
    VB_GUARD_BEGIN
    if (!ctx || !g || !p || (!out && !regions)) throw vb_error(VB_ERR_ARG, "vb_align_regions: bad arguments");
    vb_align_common(ctx, g, pairs, p, out, regions);
    VB_GUARD_END
}

// device result of vb_align_fast -> vb_align_out (keys are ref << gbits | query in LZ-ANI ids, already sorted)
static vb_align_out *align_out_from_fast(vb_ctx *ctx, const vb_genomes *g, const AlignFastOut &fo)
{
    cudaStream_t st = (cudaStream_t)ctx->stream;
    char *pin = (char *)vb_pinned(ctx, 20 * fo.n + 16);
    const uint64_t *keys = (const uint64_t *)pin;
    const int32_t *stats = (const int32_t *)(pin + 8 * fo.n);
    if (fo.n) {
        VB_CUDA(cudaMemcpyAsync(pin, fo.keys.p, sizeof(uint64_t) * fo.n, cudaMemcpyDeviceToHost, st));
        VB_CUDA(cudaMemcpyAsync(pin + 8 * fo.n, fo.stats.p, sizeof(int32_t) * 3 * fo.n, cudaMemcpyDeviceToHost, st));
    }
    vb_align_out *res = vb_align_out_alloc(fo.n, g->count());     // (host allocation overlaps the transfer)
    VB_CUDA(cudaStreamSynchronize(st));
    const std::vector<uint32_t> &order = vb_lz_order(g);
    std::copy(order.begin(), order.end(), res->order);
    const uint64_t qmask = (1ULL << fo.gbits) - 1;
    vb_parallel_for(fo.n, 65536, 8, [&](uint64_t lo, uint64_t hi) {
        for (uint64_t i = lo; i < hi; ++i) {
            res->ref[i] = (uint32_t)(keys[i] >> fo.gbits); res->qry[i] = (uint32_t)(keys[i] & qmask);
            res->sym_in_matches[i] = stats[3 * i]; res->sym_in_literals[i] = stats[3 * i + 1]; res->no_components[i] = stats[3 * i + 2];
        }
    });
    return res;
}

static void vb_align_common(vb_ctx *ctx, const vb_genomes *g, const vb_pairs *pairs, const vb_align_params *p, vb_align_out **out,
                            vb_regions **regions)
{
    {
    vb_enter(ctx);
    ctx->clear_timings("align.");
    const auto h0 = std::chrono::steady_clock::now();
    const uint32_t n = g->count();
    if (!regions && out && !getenv("VB_ALIGN_HOST_LIST")) {
        // fast path: the directed pair list, its order and the schedule are built on the device -- from the candidate list
        // the preceding vb_prefilter left there, else from an upload of (row, col, ani)
        cudaStream_t st = (cudaStream_t)ctx->stream;
        const DevGenomes &dg = vb_get_dev_genomes(ctx, g, (uint32_t)p->mrd + 128);
        const uint64_t np = pairs ? pairs->n_pairs : 0;
        const uint64_t *d_keys = nullptr;
        const float *d_ani = nullptr;
        DevBuf<uint64_t> up_keys;
        DevBuf<float> up_ani;
        std::vector<uint64_t> h_keys;
        std::vector<float> h_ani;
        if (pairs && ctx->dev_pairs && ctx->dev_pairs->uid == vb_pairs_uid(pairs) && ctx->dev_pairs->n == np) {
            d_keys = ctx->dev_pairs->keys.p; d_ani = ctx->dev_pairs->ani.p;
        } else if (pairs && np) {
            h_keys.resize(np); h_ani.resize(np);
            for (uint64_t i = 0; i < np; ++i) {
                if (pairs->row[i] >= n || pairs->col[i] >= n) throw vb_error(VB_ERR_ARG, "vb_align: pair id out of range");
                h_keys[i] = ((uint64_t)pairs->row[i] << 32) | pairs->col[i];
                h_ani[i] = pairs->ani ? (float)pairs->ani[i] : 0.9f;
            }
            up_keys.alloc(np); up_ani.alloc(np);
            VB_CUDA(cudaMemcpyAsync(up_keys.p, h_keys.data(), sizeof(uint64_t) * np, cudaMemcpyHostToDevice, st));
            VB_CUDA(cudaMemcpyAsync(up_ani.p, h_ani.data(), sizeof(float) * np, cudaMemcpyHostToDevice, st));
            d_keys = up_keys.p; d_ani = up_ani.p;
        }
        AlignFastOut fo;
        if (vb_align_fast(ctx, g, dg, p, d_keys, d_ani, np, pairs == nullptr && n >= 2, 1, 0, fo)) {
            const auto h1 = std::chrono::steady_clock::now();
            *out = align_out_from_fast(ctx, g, fo);
            ctx->set_timing("align.host_post_ms", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h1).count());
            ctx->set_timing("align.api_prep_ms", 0.0);
            return;
        }
        vb_enter(ctx);                     // the references do not fit in one batch: host-built list, batches of references
    }
    // which genomes are references?  (every genome that occurs in a pair: the list is symmetrised)  The reference side
    // of the work is launched now; the pair list is built on the host while the GPU indexes the references.
    std::vector<uint8_t> is_ref(n, pairs ? 0 : 1);
    if (pairs)
        for (uint64_t i = 0; i < pairs->n_pairs; ++i) {
            uint32_t a = pairs->row[i], b = pairs->col[i];
            if (a >= n || b >= n) throw vb_error(VB_ERR_ARG, "vb_align: pair id out of range");
            is_ref[a] = 1; is_ref[b] = 1;
        }
    if (!pairs && n < 2) std::fill(is_ref.begin(), is_ref.end(), 0);
    auto lap = [&](const char *key) {
        ctx->set_timing(key, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h0).count());
    };
    vb_align_job *job = vb_align_job_begin(ctx, g, p, is_ref.data());
    lap("align.hp1_begin_ms");
    vb_align_out *res = nullptr;
    try {
        std::vector<uint32_t> order = vb_lz_order(g);
        lap("align.hp2_order_ms");
        std::vector<uint32_t> rank(n);
        for (uint32_t i = 0; i < n; ++i) rank[order[i]] = i;
        // directed pair list in LZ-ANI ids, grouped by reference with the queries ascending (results rows are sorted,
        // :253); built as a CSR (count, prefix, fill, sort each row)
        std::vector<uint64_t> start(n + 1, 0);
        if (!pairs) {
            for (uint32_t r = 0; r < n; ++r) start[r + 1] = start[r] + (n - 1);
        } else {
            for (uint64_t i = 0; i < pairs->n_pairs; ++i) {
                start[rank[pairs->row[i]] + 1]++;        // filter[i].push(id); filter[id].push(i)
                start[rank[pairs->col[i]] + 1]++;
            }
            for (uint32_t r = 0; r < n; ++r) start[r + 1] += start[r];
        }
        const uint64_t total = start[n];
        std::vector<float> cost;
        res = vb_align_out_alloc(total, n);
        std::copy(order.begin(), order.end(), res->order);
        if (!pairs) {
            uint64_t w = 0;
            for (uint32_t r = 0; r < n; ++r)
                for (uint32_t q = 0; q < n; ++q) if (q != r) { res->ref[w] = r; res->qry[w] = q; ++w; }
        } else {
            // directed pairs sorted by (reference, query) with two stable counting passes (by query, then by reference)
            // instead of one small sort per row.  The estimated cost rides along; cost model: a parse scans the query
            // once (extension, ~0.25 instructions per base) and pays ~1000 instructions per seed event; events happen
            // where the sliding-window rule (more than 7 mismatches in 15) fires, i.e. at a rate of about
            // C(15,8) p^8 (1-p)^7 per base at divergence p
            auto rate = [](double ani) {
                const double p = std::min(std::max(1.0 - ani, 0.0), 0.5), q = 1.0 - p;
                const double p2 = p * p, p4 = p2 * p2, q2 = q * q, q4 = q2 * q2;
                return 0.25 + 1000.0 * 6435.0 * (p4 * p4) * (q4 * q2 * q);
            };
            struct Dir { uint32_t ref, qry; float cost; };
            std::vector<Dir> by_qry(total);
            {
                std::vector<uint64_t> fill(start.begin(), start.end() - 1);      // the list is symmetric: same counts
                for (uint64_t i = 0; i < pairs->n_pairs; ++i) {
                    const uint32_t a = rank[pairs->row[i]], b = rank[pairs->col[i]];
                    const double r = pairs->ani ? rate(pairs->ani[i]) : 1.0;
                    by_qry[fill[b]++] = {a, b, (float)(r * (double)g->length(pairs->col[i]))};   // reference a, query b
                    by_qry[fill[a]++] = {b, a, (float)(r * (double)g->length(pairs->row[i]))};   // reference b, query a
                }
            }
            cost.resize(total);
            std::vector<uint64_t> fill(start.begin(), start.end() - 1);
            for (uint64_t i = 0; i < total; ++i) {                                // by_qry is grouped by query, ascending
                const Dir &d = by_qry[i];
                const uint64_t w = fill[d.ref]++;
                res->ref[w] = d.ref; res->qry[w] = d.qry; cost[w] = d.cost;
            }
        }
        lap("align.hp3_csr_ms");
        std::vector<uint32_t> in_ref(total), in_qry(total);
        for (uint64_t w = 0; w < total; ++w) { in_ref[w] = order[res->ref[w]]; in_qry[w] = order[res->qry[w]]; }
        std::vector<int32_t> stats(3 * std::max<uint64_t>(total, 1));
        const double api_prep_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h0).count();
        std::vector<int32_t> rec;
        vb_align_job_run(job, in_ref.data(), in_qry.data(), total, stats.data(), regions ? &rec : nullptr,
                         cost.empty() ? nullptr : cost.data());
        ctx->set_timing("align.api_prep_ms", api_prep_ms);
        lap("align.hp6_run_done_ms");
        if (regions) *regions = vb_regions_build(rec, in_ref.data(), in_qry.data(), p->mrd);
        for (uint64_t i = 0; i < total; ++i) {
            res->sym_in_matches[i] = stats[3 * i];
            res->sym_in_literals[i] = stats[3 * i + 1];
            res->no_components[i] = stats[3 * i + 2];
        }
    } catch (...) {
        vb_align_job_end(job);
        if (res) vb_align_out_free(res);
        throw;
    }
    vb_align_job_end(job);
    if (out) *out = res; else vb_align_out_free(res);
    lap("align.hp7_end_ms");
    }
}

int vb_write_ani(const vb_genomes *g, const vb_align_out *res, const char *ani_path, const char *ids_path,
                 const char *const *columns, int n_columns, const double out_filters[5])
{
    VB_GUARD_BEGIN
    if (!g || !res || !ani_path || !ids_path || (n_columns && !columns)) throw vb_error(VB_ERR_ARG, "vb_write_ani: bad arguments");
    vb_write_ani_impl(g, res, ani_path, ids_path, columns, n_columns, out_filters);
    VB_GUARD_END
}

void vb_align_out_free(vb_align_out *r)
{
    if (!r) return;
    free(r->ref); free(r->qry); free(r->sym_in_matches); free(r->sym_in_literals); free(r->no_components); free(r->order);
    free(r);
}

}  // extern "C"
