/*
 * vclust_b200.h -- C ABI of libvclust_b200.so: the B200-native replacement for the two compute stages of Vclust,
 * `vclust prefilter` (Kmer-db build + all2all-sp + distance) and `vclust align` (LZ-ANI all2all).
 *
 * The reference has no in-process API for this path: vclust.py shells out to the kmer-db / lz-ani binaries and
 * exchanges files (reference vclust.py:915-1181 build the argv, :762-807 runs it).  Each entry point below
 * therefore names the reference *invocation* (or the function inside the tool) that it replaces.
 *
 * Conventions
 *   - every function returns 0 on success and a negative vb_status on failure; the message is available from
 *     vb_last_error() (thread-local).  Nothing throws or exits across this boundary, nothing is printed.
 *   - inputs are borrowed for the duration of the call; every vb_* object returned through an out-pointer is
 *     owned by the library and released with the matching *_free.
 *   - plain pointers and sizes only; no CUDA / torch types appear in any signature.
 *   - the library needs an sm_100 CUDA device; there is NO CPU fallback (vb_ctx_create fails without a device).
 */
#ifndef VCLUST_B200_H
#define VCLUST_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum vb_status {
    VB_OK = 0,
    VB_ERR_ARG = -1,      /* bad argument */
    VB_ERR_IO = -2,       /* file cannot be opened / parsed */
    VB_ERR_CUDA = -3,     /* CUDA runtime error or no usable device */
    VB_ERR_MEM = -4,      /* out of host or device memory */
    VB_ERR_MISMATCH = -5, /* filter file and FASTA disagree (lz-ani lz_matcher.cpp:43-75) */
    VB_ERR_INTERNAL = -6
} vb_status;

typedef struct vb_ctx vb_ctx;         /* one CUDA device + stream + scratch */
typedef struct vb_genomes vb_genomes; /* host-side genome set (names + sequences) */

/* How FASTA text is turned into genomes.  The two reference tools read the same file slightly differently:
 * VB_FASTA_KMERDB follows kmer-db (genome_input_file.h:287-338, loader_ex.cpp:150-257): split at every '>',
 *   directory mode pools all records of a file into one sample (k-mers never span records);
 * VB_FASTA_LZANI follows lz-ani (seq_reservoir.cpp:90-210): line based, directory mode joins the records of a
 *   file with `mrd` N symbols, and in multi-FASTA mode an unterminated last line is ignored. */
typedef enum vb_fasta_flavor { VB_FASTA_KMERDB = 0, VB_FASTA_LZANI = 1 } vb_fasta_flavor;

/* ---- library / device ------------------------------------------------------------------------------------- */
int vb_version(char *buf, size_t n);                 /* replaces `kmer-db -version` / `lz-ani --version` (vclust.py:1296-1334) */
int vb_device_count(void);                           /* number of CUDA devices visible (0 => nothing below can run) */
const char *vb_last_error(void);

int vb_ctx_create(int device, vb_ctx **out);
/* Same, but every kernel and copy of the context is enqueued on the caller's CUDA stream (`cuda_stream` is a
 * cudaStream_t passed as a plain pointer; NULL = vb_ctx_create's own non-blocking stream).  A host that drives
 * collectives on a stream of its own (torch.distributed / NCCL) hands that stream in, so the library's kernels and the
 * collectives are ordered by the stream and need no host synchronisation in between. */
int vb_ctx_create_on_stream(int device, void *cuda_stream, vb_ctx **out);
void vb_ctx_destroy(vb_ctx *ctx);
/* Timings of the last vb_prefilter / vb_align call on this context, in milliseconds (CUDA events on the
 * context's stream).  Keys are fixed strings; unknown key => returns VB_ERR_ARG. */
int vb_ctx_timing(const vb_ctx *ctx, const char *key, double *ms);
/* Number of kernel launches issued by the library on this context since creation. */
uint64_t vb_ctx_launches(const vb_ctx *ctx);
/* CUDA events on the context's stream, for callers that time a sequence of calls on the device:
 * vb_ctx_mark records event `slot` (0..7); vb_ctx_elapsed_ms waits for slot_b and returns slot_a -> slot_b. */
int vb_ctx_mark(vb_ctx *ctx, int slot);
int vb_ctx_elapsed_ms(vb_ctx *ctx, int slot_a, int slot_b, double *ms);

/* ---- genomes ------------------------------------------------------------------------------------------------ */
/* Replaces the FASTA ingest of both tools.  paths: one multi-FASTA file (multisample != 0) or the files of a
 * directory, one genome per file (multisample == 0; the genome is named after the file, extension included).
 * gzip input is detected by magic.  sep_len: number of N symbols between the records of one file in directory
 * mode (lz-ani: --mrd; ignored for VB_FASTA_KMERDB, which uses a single k-mer-breaking separator). */
int vb_genomes_load(const char *const *paths, int n_paths, int multisample, vb_fasta_flavor flavor, int sep_len,
                    vb_genomes **out);
/* Same, from memory: n sequences of ASCII bases (used by the bench and the tests; no file I/O). */
int vb_genomes_from_memory(const char *const *names, const char *const *seqs, const uint64_t *lens, uint32_t n,
                           vb_genomes **out);
/* Names and lengths only, no sequence data: what the ranks of a multi-GPU run know about the genomes they do not hold
 * (vb_shard_create), and all that vb_write_filter / vb_write_ani need. */
int vb_genomes_skeleton(const char *const *names, const uint64_t *lens, uint32_t n, vb_genomes **out);
/* Keep the 2-bit packed copy of g in HBM on this context until vb_genomes_evict / vb_ctx_destroy, so that later
 * vb_prefilter (rule VB_FASTA_KMERDB: U == T) or vb_align* (rule VB_FASTA_LZANI: U == N, padded for `mrd`) calls
 * on the same g start with their input already on the device.  Without it every call uploads g itself.
 * g must outlive its residency; call vb_genomes_evict(ctx, g) (g == NULL: all) before freeing it. */
int vb_genomes_make_resident(vb_ctx *ctx, const vb_genomes *g, vb_fasta_flavor rule, int mrd);
int vb_genomes_evict(vb_ctx *ctx, const vb_genomes *g);
uint32_t vb_genomes_count(const vb_genomes *g);
const char *vb_genomes_name(const vb_genomes *g, uint32_t i);
uint64_t vb_genomes_length(const vb_genomes *g, uint32_t i);
uint64_t vb_genomes_total_bases(const vb_genomes *g);
/* the ASCII sequence of genome i as the stage will see it (records of a file already joined); borrowed, not NUL-terminated */
const char *vb_genomes_sequence(const vb_genomes *g, uint32_t i);
void vb_genomes_free(vb_genomes *g);

/* ---- prefilter ---------------------------------------------------------------------------------------------- */
/* Flags of `vclust prefilter` (vclust.py:192-262) that reach kmer-db (vclust.py:915-1055). */
typedef struct vb_prefilter_params {
    int32_t k;              /* -k, 15..30            -> kmer-db build -k */
    int32_t min_kmers;      /* --min-kmers           -> all2all-sp -min num-kmers:X  (sparse_filters.h:24-29) */
    double min_ident;       /* --min-ident           -> -min ani-shorter:Y and distance -min Y (params.cpp:28-32) */
    double kmers_fraction;  /* --kmers-fraction      -> build -f (filter.h:33-146) */
    int32_t max_seqs;       /* --max-seqs            -> -sample-rows ani-shorter:N (sampler.h:45-121); 0 = off */
    int32_t batch_size;     /* --batch-size          -> tiling hint only; the result never depends on it */
} vb_prefilter_params;

/* Sparse lower-triangular result: for every kept pair row > col (input order ids), the number of distinct
 * canonical k-mers the two genomes share and the ani-shorter value; plus total_kmers per genome (the set sizes,
 * kmer_db.h:129).  Sorted by (row, col) -- the order of the reference's all2all.txt rows. */
typedef struct vb_pairs {
    uint64_t n_pairs;
    uint32_t *row;
    uint32_t *col;
    uint32_t *common;
    double *ani;            /* ani-shorter computed in IEEE double exactly as params.cpp:28-32 */
    uint32_t n_genomes;
    uint32_t *total_kmers;
    int32_t k;
    double kmers_fraction;
} vb_pairs;

/* Replaces: kmer-db build [-multisample-fasta] -k K -f F  +  kmer-db all2all-sp -sparse -min num-kmers:X
 *           -min ani-shorter:Y [-sample-rows ani-shorter:N]   (console_build.cpp:33, similarity_calculator.cpp:442). */
int vb_prefilter(vb_ctx *ctx, const vb_genomes *g, const vb_prefilter_params *p, vb_pairs **out);
/* Multi-GPU building blocks (no reference counterpart; the reference is single-process):
 * vb_prefilter_partial counts only the k-mers of one hash shard (fmix64(kmer) % shard_count == shard_index) and
 * returns EVERY pair with a non-zero partial count (no thresholds, ani = 0) plus the shard's part of total_kmers;
 * summing the parts of all shards gives exactly vb_prefilter's integers.  vb_pairs_merge does that sum for a
 * concatenation of partial triples (any order, duplicates allowed), applies the two -min filters with the exact
 * double metric and returns the same vb_pairs vb_prefilter would. */
int vb_prefilter_partial(vb_ctx *ctx, const vb_genomes *g, const vb_prefilter_params *p, uint32_t shard_index,
                         uint32_t shard_count, vb_pairs **out);
int vb_pairs_merge(const uint32_t *row, const uint32_t *col, const uint32_t *common, uint64_t n,
                   const uint32_t *total_kmers, uint32_t n_genomes, const vb_prefilter_params *p, vb_pairs **out);
/* Replaces: kmer-db distance ani-shorter -sparse -min Y (console_distance.cpp:7-213): writes the filter text. */
int vb_write_filter(const vb_genomes *g, const vb_pairs *pairs, const char *path);
/* Replaces: CFilter::load_filter (lz-ani filter.cpp:20-298): reads a filter text, keeps entries >= thr, checks
 * that the header names equal the genome names (lz_matcher.cpp:43-75).  Output rows/cols as in the file. */
int vb_read_filter(const char *path, double thr, const vb_genomes *g, vb_pairs **out);
void vb_pairs_free(vb_pairs *p);

/* ---- align -------------------------------------------------------------------------------------------------- */
/* Flags of `vclust align` that reach lz-ani (vclust.py:1142-1179; defaults lz-ani params.h:38-45). */
typedef struct vb_align_params {
    int32_t mal, msl, mrd, mqd, reg, aw, am, ar;
} vb_align_params;

/* Per directed pair: query `qry` parsed against reference `ref` (ids in LZ-ANI order: length desc, name asc --
 * seq_reservoir.cpp:215-251) and the three integers of results_t (lz-ani defs.h:48-65). order[i] = input id of
 * the genome with LZ-ANI id i. Sorted by (ref, qry). */
typedef struct vb_align_out {
    uint64_t n;
    uint32_t *ref;
    uint32_t *qry;
    int32_t *sym_in_matches;
    int32_t *sym_in_literals;
    int32_t *no_components;
    uint32_t n_genomes;
    uint32_t *order;
} vb_align_out;

/* Replaces: lz-ani all2all ... [--flt-kmerdb F thr]  up to and including do_matching (lz_matcher.cpp:172-277;
 * parser.cpp:16-50,482-783).  pairs == NULL => all-vs-all (every ordered pair).  pairs use INPUT-order ids and
 * are symmetrised exactly like filter.cpp:80-81,253-289 (duplicates kept). */
int vb_align(vb_ctx *ctx, const vb_genomes *g, const vb_pairs *pairs, const vb_align_params *p, vb_align_out **out);
/* Low-level variant used by the bench/tests: explicit directed pair list in INPUT-order ids, no reordering;
 * stats[3*i..] = (sym_in_matches, sym_in_literals, no_components) of pair i. */
int vb_align_pairs(vb_ctx *ctx, const vb_genomes *g, const uint32_t *ref, const uint32_t *qry, uint64_t n,
                   const vb_align_params *p, int32_t *stats);
/* Assemble the result object of vb_align [LZ-ANI ids, sorted by (ref, qry)] from directed pairs in INPUT-order ids and their
 * 3-int statistics, e.g. gathered from several GPUs. */
int vb_align_out_from_pairs(const vb_genomes *g, const uint32_t *ref, const uint32_t *qry, const int32_t *stats, uint64_t n,
                            vb_align_out **out);
/* Replaces: CLZMatcher::store_results (lz_matcher.cpp:280-579): writes <ani_path> and <ids_path>.
 * columns: the --out-format list (vclust.py:38-47); out_filters: minimum tani, gani, ani, qcov, rcov (0 = off). */
int vb_write_ani(const vb_genomes *g, const vb_align_out *res, const char *ani_path, const char *ids_path,
                 const char *const *columns, int n_columns, const double out_filters[5]);
void vb_align_out_free(vb_align_out *r);

/* ---- alignment regions (vclust align --out-aln -> lz-ani --out-alignment) ------------------------------------ */
/* One row per local alignment (a component of the parse at least `reg` symbols long; parser.cpp:786-837 calc_regions).
 * ref / qry are INPUT-order genome ids; coordinates are 0-based half-open: [q_start, q_end) in the query and
 * [r_start, r_end) in the reference TEXT (forward genome, 2*mrd separators, reverse complement), exactly the
 * region_t fields of lz-ani defs.h:67-142.  Rows are grouped by directed pair (in the order of the call's result),
 * inside a pair sorted like calc_regions: longer first, then smaller q_start. */
typedef struct vb_regions {
    uint64_t n;
    uint32_t *ref;
    uint32_t *qry;
    int32_t *q_start, *q_end, *r_start, *r_end, *matches, *mismatches;
    int32_t mrd;            /* needed to map reverse-strand coordinates (lz_matcher.cpp:113,158-162) */
} vb_regions;

/* vb_align that also returns the regions of every directed pair (either out pointer may be NULL). */
int vb_align_regions(vb_ctx *ctx, const vb_genomes *g, const vb_pairs *pairs, const vb_align_params *p, vb_align_out **out,
                     vb_regions **regions);
/* vb_align_pairs that also returns the regions (explicit directed pairs, INPUT-order ids). */
int vb_align_pairs_regions(vb_ctx *ctx, const vb_genomes *g, const uint32_t *ref, const uint32_t *qry, uint64_t n,
                           const vb_align_params *p, int32_t *stats, vb_regions **regions);
/* Replaces: CLZMatcher::store_alignment (lz_matcher.cpp:102-169): writes the alignment TSV (header lz_matcher.h:68).
 * out_filters as in vb_write_ani; only gani, ani and qcov apply here, computed per directed pair from its regions.
 * The reference writes the pairs in thread-completion order; this writer uses the deterministic order of `regions`. */
int vb_write_aln(const vb_genomes *g, const vb_regions *regions, const char *path, const double out_filters[5]);
void vb_regions_free(vb_regions *r);

/* ---- one genome set over the GPUs of one node ------------------------------------------------------------------
 * No reference counterpart (kmer-db and lz-ani are single-process; their analogue of the split is the tiling of
 * all2all-parts, console_all2all_parts.cpp:143-331, and lz-ani's per-reference work queue, lz_matcher.cpp:190-270).
 * One process per GPU.  The genomes are block-partitioned: rank r loads genomes [first_id, first_id + count(local)) and
 * only knows names and lengths of the rest (`meta`).  The library does the compute and owns every device buffer; the
 * HOST supplies the collectives as callbacks (torch.distributed over NCCL in vclust_b200/distributed.py), all of them
 * on DEVICE memory and enqueued on the context's stream (vb_ctx_create_on_stream):
 *   all_to_all   records of `width` bytes; send_counts[p] records go to peer p (the send buffer is ordered by peer),
 *                recv_counts[p] records arrive from peer p (both sides already know the counts)
 *   all_gather   `bytes` bytes from every rank, concatenated in rank order
 *   all_reduce_sum_u32   in place
 * Each returns 0 on success.  Data path of vb_shard_prefilter + vb_shard_align:
 *   setup     pack the local genomes, all-gather the packed align-stage records (every GPU can read every genome)
 *   extract   k-mers of the LOCAL genomes -> (hash, genome) tuples, partitioned by hash range
 *   exchange  all-to-all #1: every tuple travels to the rank that owns its hash range
 *   count     grouping + pair counting on the owner's range -> partial common-k-mer counts per genome pair
 *   reduce    all-reduce of total-kmers; all-to-all #2: partial counts travel to the owners of both genomes
 *             (owner(g) = g mod world), which sum them and apply the -min filters
 *   align     every owner parses (reference = own genome, query) for its candidate pairs; results are gathered on rank 0 */
typedef struct vb_comm {
    int32_t rank, world;
    void *user;
    int (*all_to_all)(void *user, const void *send, const uint64_t *send_counts, void *recv, const uint64_t *recv_counts,
                      uint32_t width);
    int (*all_gather)(void *user, const void *send, void *recv, uint64_t bytes);
    int (*all_reduce_sum_u32)(void *user, void *buf, uint64_t n);
} vb_comm;

typedef struct vb_shard vb_shard;
/* Collective.  meta: all genomes (names + lengths; vb_genomes_skeleton or a full set); local: this rank's block, with
 * sequences; comm is copied (its callbacks must stay valid until vb_shard_destroy).  mrd: the --mrd the align stage
 * will use (padding of the packed store). */
int vb_shard_create(vb_ctx *ctx, const vb_comm *comm, const vb_genomes *meta, const vb_genomes *local, uint32_t first_id,
                    int mrd, vb_shard **out);
/* Collective.  On rank 0 *out is the complete result of vb_prefilter on the whole set; on the other ranks it holds no
 * pairs.  Every rank keeps its share of the candidate list on the device for vb_shard_align. */
int vb_shard_prefilter(vb_shard *sh, const vb_prefilter_params *p, vb_pairs **out);
/* Collective, after vb_shard_prefilter.  On rank 0 *out is the complete result of vb_align; elsewhere it is empty. */
int vb_shard_align(vb_shard *sh, const vb_align_params *p, vb_align_out **out);
void vb_shard_destroy(vb_shard *sh);
/* Exercises the three callbacks of comm on small device buffers and checks what comes back (used by the tests). */
int vb_comm_selftest(vb_ctx *ctx, const vb_comm *comm);

#ifdef __cplusplus
}
#endif
#endif /* VCLUST_B200_H */
