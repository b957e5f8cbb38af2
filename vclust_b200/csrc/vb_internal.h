// Internal declarations shared by the translation units of libvclust_b200.so (not part of the C ABI).
#pragma once
#include <algorithm>
#include <cstdint>
#include <new>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/vclust_b200.h"

struct vb_error : std::runtime_error {
    int code;
    vb_error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

void vb_set_error(const std::string &msg);

uint64_t vb_next_uid();

// growable byte buffer WITHOUT value-initialisation: a 400 MB FASTA is read straight into it and compacted in place, so
// every page is touched once (std::vector<char>::resize would zero-fill the pages first)
struct vb_bytes {
    char *p = nullptr;
    size_t n = 0, cap = 0;
    vb_bytes() = default;
    vb_bytes(const vb_bytes &) = delete;
    vb_bytes &operator=(const vb_bytes &) = delete;
    ~vb_bytes() { free(p); }
    char *data() { return p; }
    const char *data() const { return p; }
    size_t size() const { return n; }
    bool empty() const { return n == 0; }
    void reserve(size_t c)
    {
        if (c <= cap) return;
        char *q = (char *)realloc(p, c);
        if (!q) throw std::bad_alloc();
        p = q; cap = c;
    }
    void resize(size_t m) { if (m > cap) reserve(std::max(m, cap + cap / 2)); n = m; }
    void append(const char *s, size_t len) { resize(n + len); if (len) memcpy(p + n - len, s, len); }
    void swap(vb_bytes &o) { std::swap(p, o.p); std::swap(n, o.n); std::swap(cap, o.cap); }
    void shrink_to_fit()
    {
        if (n == cap || !p) return;
        char *q = (char *)realloc(p, n ? n : 1);
        if (q) { p = q; cap = n ? n : 1; }
    }
};

// Host-side genome set.  Sequences are kept as ASCII exactly as read (separators between the records of one
// file already inserted as 'N' bytes); symbol coding happens on the device, per stage, because kmer-db and
// lz-ani disagree on 'U' (kmer-db alphabet.h:80-85 vs lz-ani seq_reservoir.h:243-247).
struct vb_genomes {
    std::vector<std::string> names;
    std::vector<uint64_t> offset;   // n+1 offsets into bases
    vb_bytes bases;                 // concatenated ASCII
    vb_fasta_flavor flavor = VB_FASTA_KMERDB;
    // One vb_genomes may be used by one thread at a time (the lazily filled caches below are not synchronised across
    // concurrent calls on the SAME set; different sets / contexts are independent).
    mutable bool pinned = false;    // bases page-locked with cudaHostRegister (done lazily by the second upload)
    mutable int uploads = 0;
    bool skeleton = false;          // names + lengths only (vb_genomes_skeleton): `bases` is empty
    uint64_t uid = vb_next_uid();   // distinguishes a new set that re-uses the address of a freed one (device-copy cache)
    // LZ-ANI order (seq_reservoir.cpp:215-251): computed once per set (a string sort of all names), used by every vb_align
    mutable std::vector<uint32_t> lz_order, lz_rank;
    uint32_t count() const { return (uint32_t)names.size(); }
    uint64_t length(uint32_t i) const { return offset[i + 1] - offset[i]; }
};
const std::vector<uint32_t> &vb_lz_order(const vb_genomes *g);     // order[i] = input id of the genome with LZ-ANI id i
const std::vector<uint32_t> &vb_lz_rank(const vb_genomes *g);      // rank[input id] = LZ-ANI id

// fn(lo, hi) over [0, n) on up to `max_threads` host threads (inline when the range is small)
void vb_parallel_for(uint64_t n, uint64_t min_per_thread, unsigned max_threads, const std::function<void(uint64_t, uint64_t)> &fn);

// vb_pairs as allocated by the library: the public struct plus an identity, so that a context can recognise the list
// its last prefilter call produced and use the copy it kept on the device
struct vb_pairs_box {
    vb_pairs pub;
    uint64_t uid;
};
inline uint64_t vb_pairs_uid(const vb_pairs *p) { return p ? ((const vb_pairs_box *)p)->uid : 0; }

// ---- text formats (host_format.cpp) ---------------------------------------------------------------------------
int vb_fmt_fixed6(double v, char *out);                 // kmer-db conversion.h:167-219 Double2PChar(v, 6)
int vb_fmt_real(double v, int prec, char *out);         // refresh numeric_conversions.h:229-299 real_to_pchar
void vb_write_aln_impl(const vb_genomes *g, const vb_regions *regions, const char *path, const double out_filters[5]);
double vb_ani_shorter(uint32_t common, uint32_t cnt1, uint32_t cnt2, int k);   // kmer-db params.cpp:28-32
// kmer-db -sample-rows ani-shorter:N on a list of passing pairs (row > col); rewrites the list, sorted by (row, item)
void vb_sample_rows(uint32_t n_genomes, uint32_t max_items, std::vector<uint32_t> &row, std::vector<uint32_t> &col,
                    std::vector<uint32_t> &common, std::vector<double> &ani);

// ---- device side (declared here, defined in the .cu files) -----------------------------------------------------
struct vb_timing { std::string key; double ms; };

struct DevGenomes;
struct DevPairs;
struct vb_arena;
struct vb_resident {                 // a genome set kept packed in HBM across calls (vb_genomes_make_resident)
    const vb_genomes *g;
    uint64_t uid;
    uint32_t min_pad;
    DevGenomes *dev;
};

// Receive buffers of all-to-all #1, one per rank, mapped into every rank's address space (CUDA IPC, shard.cu): a rank
// copies each destination's slice of its level-1 buffer straight into the owner's buffer over NVLink (one peer
// cudaMemcpyAsync per destination; a fused store from the partition kernel was slower: tiny NVLink writes).
struct vb_peer_xbuf {
    void *local = nullptr;               // this rank's buffer (cudaMalloc)
    uint64_t cap = 0;                    // bytes, the same on every rank
    std::vector<void *> peer;            // peer[r]: rank r's buffer as seen from this process (peer[rank] == local)
};

struct vb_ctx {
    int device = 0;
    void *stream = nullptr;          // cudaStream_t
    bool owns_stream = true;         // false: the caller's stream (vb_ctx_create_on_stream)
    void *events[8] = {nullptr};     // cudaEvent_t, vb_ctx_mark / vb_ctx_elapsed_ms
    void *copy_stream = nullptr;     // cudaStream_t: chunked H2D of a genome upload, overlapped with the kernels on `stream`
    void *copy_events[9] = {nullptr};
    uint64_t launches = 0;
    vb_arena *arena = nullptr;       // call-scoped device temporaries (dev_util.cuh)
    uint64_t mem_total = 0;          // device memory, queried once (cudaMemGetInfo is slow and synchronising)
    std::vector<vb_resident> resident;
    vb_resident last = {nullptr, 0, 0, nullptr};   // the most recent upload that was not made resident (implicit cache)
    DevPairs *dev_pairs = nullptr;   // candidate list of the last vb_prefilter, kept on the device for vb_align
    uint64_t pair_hint_uid = 0, pair_hint_entries = 0;    // distinct pairs the last hashed-table prefilter of set `uid` produced
    uint32_t pair_hint_n = 0;
    int pair_hint_k = 0;
    struct vb_peer_xbuf *xbuf = nullptr;   // exchange buffers mapped between the ranks (kept across vb_shard_create calls)
    void *pin_buf = nullptr;         // page-locked staging buffer for result read-backs (grown on demand, kept)
    size_t pin_cap = 0;
    std::vector<vb_timing> timings;
    void set_timing(const std::string &k, double ms) {
        for (auto &t : timings) if (t.key == k) { t.ms = ms; return; }
        timings.push_back({k, ms});
    }
    void add_timing(const std::string &k, double ms) {
        for (auto &t : timings) if (t.key == k) { t.ms += ms; return; }
        timings.push_back({k, ms});
    }
    void clear_timings(const std::string &prefix) {            // at the start of every top-level call of that stage
        timings.erase(std::remove_if(timings.begin(), timings.end(), [&](const vb_timing &t) { return t.key.compare(0, prefix.size(), prefix) == 0; }),
                      timings.end());
    }
};

// The prefilter pipeline (prefilter.cu).  One rank of `comm` (nullptr: the whole job) works on the genomes of `g`, whose
// global ids start at gid_base; n_total / total_slots_all describe the whole set.  shard_index / shard_count: the legacy
// k-mer shard of vb_prefilter_partial (no thresholds).  keep_dev: leave the final candidate list on the device
// (ctx->dev_pairs) for the align stage.
struct vb_prefilter_job {
    const vb_genomes *g = nullptr;
    const vb_peer_xbuf *xbuf = nullptr;  // several ranks: peer copies into mapped receive buffers instead of the tuple all-to-all
    uint32_t gid_base = 0, n_total = 0;
    double est_kmers_all = 0;            // k-mers of the whole set (all ranks), for the pass / bucket plan
    const vb_comm *comm = nullptr;
    uint32_t shard_index = 0, shard_count = 1;
    bool keep_dev = true;
};
void vb_prefilter_run(vb_ctx *ctx, const vb_prefilter_job &job, const vb_prefilter_params *p, vb_pairs **out);
void vb_drop_dev_pairs(vb_ctx *ctx);
// page-locked host memory of at least `bytes` bytes, owned by the context, valid until the next call of vb_pinned
void *vb_pinned(vb_ctx *ctx, size_t bytes);
// vb_align in two steps: _begin uploads the genomes and launches the reference texts + anchor tables of the genomes
// flagged in is_ref[n_genomes] (asynchronously); _run takes the directed pairs (every reference must have been
// flagged) and returns the statistics; _end releases the device buffers (LIFO after everything _run allocated).
struct vb_align_job;
vb_align_job *vb_align_job_begin(vb_ctx *ctx, const vb_genomes *g, const vb_align_params *p, const uint8_t *is_ref);
// regions != nullptr: also collect the alignment regions (lz-ani --out-alignment), 7 ints each:
// pair index (into ref/qry), q_start, q_end, r_start, r_end (0-based half-open, reference text coordinates), matches, mismatches
// cost (optional, n entries): estimated relative cost of each parse, used only to schedule expensive pairs first
void vb_align_job_run(vb_align_job *job, const uint32_t *ref, const uint32_t *qry, uint64_t n, int32_t *stats,
                      std::vector<int32_t> *regions, const float *cost = nullptr);
void vb_align_job_end(vb_align_job *job);
void vb_align_pairs_impl(vb_ctx *ctx, const vb_genomes *g, const uint32_t *ref, const uint32_t *qry, uint64_t n,
                         const vb_align_params *p, int32_t *stats);
