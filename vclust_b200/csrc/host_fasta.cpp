// FASTA ingest for both stages (host side).  Mirrors the *reading rules* of the two reference tools, which
// differ in corner cases; see vb_fasta_flavor in include/vclust_b200.h.
//   kmer-db : genome_input_file.h:80-136 (whole-file read, gzip by magic), :287-338 (record splitting),
//             loader_ex.cpp:150-257 (directory mode: one sample per file, named by file name)
//   lz-ani  : seq_reservoir.cpp:90-153 (directory mode), :156-210 (multi-FASTA), file_wrapper.h:917-950 (getline)
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <exception>
#include <filesystem>
#include <thread>
#include <vector>

#include "vb_internal.h"

namespace {

// Appends the whole (decompressed) file to `out` and returns the number of bytes added; gzip is detected by its magic
// bytes (kmer-db genome_input_file.h:98-118; lz-ani reads through the same kind of wrapper), plain files are one fread.
size_t read_whole(const char *path, vb_bytes &out)
{
    FILE *fp = fopen(path, "rb");
    if (!fp) throw vb_error(VB_ERR_IO, std::string("Cannot open file: ") + path);
    unsigned char magic[2] = {0, 0};
    const size_t got = fread(magic, 1, 2, fp);
    const size_t start = out.size();
    size_t used = 0;
    if (got == 2 && magic[0] == 0x1f && magic[1] == 0x8b) {
        fclose(fp);
        gzFile f = gzopen(path, "rb");
        if (!f) throw vb_error(VB_ERR_IO, std::string("Cannot open file: ") + path);
        gzbuffer(f, 1 << 20);
        out.resize(start + (16 << 20));
        for (;;) {
            if (out.size() - start - used < (4u << 20)) out.resize(start + 2 * (out.size() - start));
            int n = gzread(f, out.data() + start + used, (unsigned)std::min<size_t>(out.size() - start - used, 1u << 30));
            if (n < 0) { gzclose(f); out.resize(start); throw vb_error(VB_ERR_IO, std::string("Cannot read file: ") + path); }
            if (n == 0) break;
            used += (size_t)n;
        }
        gzclose(f);
        out.resize(start + used);
        return used;
    }
    std::error_code ec;
    const auto fsize = std::filesystem::file_size(path, ec);
    rewind(fp);
    if (!ec && fsize >= (64u << 20)) {
        // a big regular file: the copy out of the page cache and the first-touch faults of the destination are the
        // whole cost of ingest, so several threads pread() disjoint ranges straight into the store
        out.resize(start + (size_t)fsize);
        char *dst = out.data() + start;
        const int fd = fileno(fp);
        std::atomic<bool> short_read{false};
        vb_parallel_for((uint64_t)fsize, 32u << 20, 8, [&](uint64_t lo, uint64_t hi) {
            while (lo < hi) {
                const ssize_t n = pread(fd, dst + lo, (size_t)std::min<uint64_t>(hi - lo, 1u << 30), (off_t)lo);
                if (n <= 0) { short_read = true; return; }
                lo += (uint64_t)n;
            }
        });
        fclose(fp);
        if (short_read) { out.resize(start); throw vb_error(VB_ERR_IO, std::string("Cannot read file: ") + path); }
        return (size_t)fsize;
    } else if (!ec && fsize > 0) {
        out.resize(start + (size_t)fsize);
        used = fread(out.data() + start, 1, (size_t)fsize, fp);
    } else {                                            // not a regular file (pipe ...): read until EOF
        out.resize(start + (1 << 20));
        for (;;) {
            if (start + used == out.size()) out.resize(start + 2 * (out.size() - start));
            size_t n = fread(out.data() + start + used, 1, out.size() - start - used, fp);
            if (n == 0) break;
            used += n;
        }
    }
    const bool bad = ferror(fp) != 0;
    fclose(fp);
    out.resize(start + used);
    if (bad) { out.resize(start); throw vb_error(VB_ERR_IO, std::string("Cannot read file: ") + path); }
    return used;
}

// Receives the records of one file in order and writes the sequence bytes straight into vb_genomes::bases (no
// per-record strings).  A multi-FASTA is read INTO bases and compacted in place -- the write position never passes
// the read position, hence memmove -- so a 400 MB input is one fread plus one memchr/memmove pass over the same pages.
//   multisample : every kept record is a genome.
//   otherwise   : the file is ONE genome named by its file name; kmer-db pools the records and k-mers must not span
//                 them -> one 'N' between consecutive records (loader_ex.cpp:150-257); lz-ani joins them with sep_len
//                 N symbols whenever the sequence so far is non-empty (seq_reservoir.cpp:90-153).
struct Sink {
    vb_genomes *g;
    bool multisample, kmerdb;
    size_t sep_len;
    char *base = nullptr;          // g->bases.data() after reserve_for()
    size_t w = 0;                  // write position in g->bases
    size_t file_start = 0, rec_start = 0;
    bool first_record = true;
    std::string rec_name;
    std::vector<std::string> *names_out = nullptr;      // where kept records are listed (default: the genome set itself;
    std::vector<uint64_t> *ends_out = nullptr;          // a chunk of a big file parsed by its own thread: local lists)

    // in place (multisample): the file's bytes already sit at bases[w ...); otherwise room for bytes + separators
    void start_file(size_t file_bytes, size_t max_records_hint, bool in_place)
    {
        if (!in_place) g->bases.resize(w + file_bytes + (max_records_hint + 1) * std::max<size_t>(sep_len, 1));
        base = g->bases.data();
        file_start = w;
        first_record = true;
    }
    void begin_record(const char *name, size_t name_len)
    {
        if (multisample) rec_name.assign(name, name_len);
        else if (kmerdb) { if (!first_record) base[w++] = 'N'; }
        else if (w > file_start) { memset(base + w, 'N', sep_len); w += sep_len; }
        first_record = false;
        rec_start = w;
    }
    inline void append(const char *p, size_t n) { memmove(base + w, p, n); w += n; }
    void end_record(bool keep)
    {
        if (!multisample) return;
        if (!keep) { w = rec_start; return; }
        size_t sp = rec_name.find(' ');
        if (sp != std::string::npos) rec_name.resize(sp);
        (names_out ? names_out : &g->names)->push_back(rec_name);
        (ends_out ? ends_out : &g->offset)->push_back(w);
    }
    void end_file(const char *path)
    {
        if (!multisample) {
            g->names.push_back(std::filesystem::path(path).filename().string());
            g->offset.push_back(w);
        }
        g->bases.resize(w);
    }
};

// append [b, e) without '\n' and '\r'
inline void append_stripped(Sink &s, const char *b, const char *e)
{
    while (b < e) {
        const char *nl = (const char *)memchr(b, '\n', (size_t)(e - b));
        const char *le = nl ? nl : e;
        const char *cr = (const char *)memchr(b, '\r', (size_t)(le - b));
        if (!cr) s.append(b, (size_t)(le - b));
        else
            for (const char *p = b; p < le; ++p) if (*p != '\r') s.base[s.w++] = *p;       // rare: CR inside a line / CRLF
        b = nl ? nl + 1 : e;
    }
}

// kmer-db rule: a record starts at every '>' byte; header ends at '\n' (a preceding '\r' is dropped) and is cut at
// the first space; the sequence is everything up to the next '>' without '\n' and '\r'.
void split_kmerdb(const char *d, size_t n, Sink &s)
{
    const char *end = d + n;
    const char *pos = (const char *)memchr(d, '>', n);
    while (pos) {
        const char *eol = (const char *)memchr(pos, '\n', (size_t)(end - pos));
        if (!eol) eol = end;
        const char *hend = eol;
        if (hend > pos + 1 && hend[-1] == '\r') --hend;
        s.begin_record(pos + 1, (size_t)(hend - pos - 1));
        const char *nxt = eol < end ? (const char *)memchr(eol, '>', (size_t)(end - eol)) : nullptr;
        const char *bend = nxt ? nxt : end;
        const char *bbeg = std::min(eol + 1, bend);
        append_stripped(s, bbeg, bend);
        s.end_record(true);
        pos = nxt;
    }
}

// lz-ani rule: line based; '>' counts only in column 0; every line loses one trailing '\r'; empty lines are skipped.
// Directory mode keeps a last line without '\n', multi-FASTA mode drops it; directory mode also keeps records with an
// empty name and a header-less leading sequence (load_fasta ignores names), multi-FASTA mode drops both.
void split_lzani(const char *d, size_t n, bool dir_mode, Sink &s)
{
    const char *end = d + n;
    const char *pos = d;
    bool have = false, open = false, named = false;      // a header was seen / a record is open / its name is non-empty
    auto close = [&]() {
        if (!open) return;
        const bool keep = dir_mode ? true : (have && named);
        s.end_record(keep);
        open = false;
    };
    auto handle = [&](const char *b, const char *e) {
        if (e > b && e[-1] == '\r') --e;
        if (e == b) return;
        if (*b == '>') {
            close();
            s.begin_record(b + 1, (size_t)(e - b - 1));
            have = true; open = true; named = e - b - 1 > 0;
        } else {
            if (!open) { s.begin_record(b, 0); open = true; named = false; }    // header-less leading sequence
            s.append(b, (size_t)(e - b));
        }
    };
    while (pos < end) {
        const char *eol = (const char *)memchr(pos, '\n', (size_t)(end - pos));
        if (!eol) {
            if (dir_mode) handle(pos, end);
            break;
        }
        handle(pos, eol);
        pos = eol + 1;
    }
    close();
}

// A big multi-FASTA (already read into bases at [s.w, s.w + n)) parsed by several threads.  The file is cut at '>' bytes
// that start a line -- a record start under both tools' rules, and the only place where a chunk can begin without
// knowing what came before -- every thread compacts its chunk in place and lists its records, and the gaps between the
// compacted chunks are then closed, again by all threads at once.
void split_parallel(const char *d, size_t n, bool kmerdb, unsigned n_threads, Sink &s)
{
    struct Chunk { size_t lo, hi, w_end; std::vector<std::string> names; std::vector<uint64_t> ends; };
    std::vector<Chunk> chunks(n_threads);
    const char *end = d + n;
    size_t prev = 0;
    for (unsigned c = 0; c < n_threads; ++c) {
        size_t hi = n;
        if (c + 1 < n_threads) {
            const char *p = d + std::max(prev, n * (size_t)(c + 1) / n_threads);
            hi = n;
            while (p < end && (p = (const char *)memchr(p, '>', (size_t)(end - p)))) {
                if (p > d && p[-1] == '\n') { hi = (size_t)(p - d); break; }
                ++p;
            }
        }
        chunks[c].lo = prev; chunks[c].hi = hi;
        prev = hi;
    }
    const size_t file_w = s.w;                          // the file's bytes start here (in-place mode)
    char *base = s.base;
    std::vector<std::thread> pool;
    auto work = [&](unsigned c) {
        Chunk &ch = chunks[c];
        Sink ls{s.g, true, kmerdb, s.sep_len};
        ls.base = base;
        ls.w = ls.file_start = file_w + ch.lo;
        ls.names_out = &ch.names; ls.ends_out = &ch.ends;
        if (ch.hi > ch.lo) {
            if (kmerdb) split_kmerdb(d + ch.lo, ch.hi - ch.lo, ls);
            else split_lzani(d + ch.lo, ch.hi - ch.lo, false, ls);
        }
        ch.w_end = ls.w;
    };
    std::vector<std::exception_ptr> failed(n_threads);          // (an exception must not leave a thread: std::terminate)
    auto guarded = [&](unsigned c) { try { work(c); } catch (...) { failed[c] = std::current_exception(); } };
    for (unsigned c = 1; c < n_threads; ++c) pool.emplace_back(guarded, c);
    guarded(0);
    for (auto &th : pool) th.join();
    pool.clear();
    for (auto &f : failed) if (f) std::rethrow_exception(f);
    // close the gaps between the compacted chunks, in place.  Chunk c moves left by shift[c] = the slack of the chunks
    // before it -- far less than its length -- so its destination overlaps only the last shift[c-1] bytes of chunk c-1's
    // compacted bytes: those tails are set aside first (a few MB), then all chunks move at the same time.
    std::vector<size_t> dst(n_threads + 1, file_w);
    for (unsigned c = 0; c < n_threads; ++c) dst[c + 1] = dst[c] + (chunks[c].w_end - (file_w + chunks[c].lo));
    auto src_of = [&](unsigned c) { return file_w + chunks[c].lo; };
    auto len_of = [&](unsigned c) { return dst[c + 1] - dst[c]; };
    bool together = true;
    for (unsigned c = 0; c + 1 < n_threads; ++c)
        if (src_of(c) - dst[c] > len_of(c)) together = false;          // (degenerate: a chunk shorter than the slack before it)
    if (together) {
        std::vector<std::string> tails(n_threads);
        for (unsigned c = 0; c + 1 < n_threads; ++c) {
            const size_t ts = src_of(c) - dst[c];
            tails[c].assign(base + src_of(c) + len_of(c) - ts, ts);
        }
        auto move = [&](unsigned c) {
            const size_t ts = tails[c].size();
            if (dst[c] != src_of(c)) memmove(base + dst[c], base + src_of(c), len_of(c) - ts);
            if (ts) memcpy(base + dst[c] + len_of(c) - ts, tails[c].data(), ts);
        };
        for (unsigned c = 1; c < n_threads; ++c) pool.emplace_back(move, c);
        move(0);
        for (auto &th : pool) th.join();
    } else {
        for (unsigned c = 0; c < n_threads; ++c)
            if (dst[c] != src_of(c)) memmove(base + dst[c], base + src_of(c), len_of(c));
    }
    for (unsigned c = 0; c < n_threads; ++c) {
        const size_t shift = file_w + chunks[c].lo - dst[c];
        for (auto &nm : chunks[c].names) s.g->names.push_back(std::move(nm));
        for (uint64_t e : chunks[c].ends) s.g->offset.push_back(e - shift);
    }
    s.w = dst[n_threads];
}

}  // namespace

vb_genomes *vb_genomes_load_impl(const char *const *paths, int n_paths, int multisample, vb_fasta_flavor flavor,
                                 int sep_len)
{
    auto *g = new vb_genomes();
    g->flavor = flavor;
    g->offset.push_back(0);
    try {
        Sink s{g, multisample != 0, flavor == VB_FASTA_KMERDB, (size_t)std::max(sep_len, 0)};
        vb_bytes tmp;
        for (int i = 0; i < n_paths; ++i) {
            const char *d;
            size_t n;
            if (multisample) {                          // read into the store, compact in place
                n = read_whole(paths[i], g->bases);
                d = g->bases.data() + s.w;
                s.start_file(n, 0, true);
            } else {                                    // separators may outgrow the headers they replace: two buffers
                tmp.resize(0);
                n = read_whole(paths[i], tmp);
                d = tmp.data();
                size_t hdrs = 0;                        // at most one separator per header line
                for (const char *p = d, *e = d + n; (p = (const char *)memchr(p, '>', (size_t)(e - p))); ++p) ++hdrs;
                s.start_file(n, hdrs, false);
            }
            // big multi-FASTA files are parsed by several threads (VB_FASTA_PAR_MIN: bytes from which on, test hook)
            const size_t par_min = getenv("VB_FASTA_PAR_MIN") ? (size_t)strtoull(getenv("VB_FASTA_PAR_MIN"), nullptr, 10) : ((size_t)64 << 20);
            const unsigned par_thr = getenv("VB_FASTA_THREADS") ? (unsigned)std::clamp(atoi(getenv("VB_FASTA_THREADS")), 1, 64) : 8u;
            const unsigned n_thr = std::min({par_thr, std::max(1u, std::thread::hardware_concurrency()),
                                             (unsigned)std::max<size_t>(1, n / std::max<size_t>(par_min / 4, 1))});
            if (multisample && n >= par_min && n_thr > 1) split_parallel(d, n, flavor == VB_FASTA_KMERDB, n_thr, s);
            else if (flavor == VB_FASTA_KMERDB) split_kmerdb(d, n, s);
            else split_lzani(d, n, !multisample, s);
            s.end_file(paths[i]);
        }
        g->bases.shrink_to_fit();
    } catch (...) {
        delete g;
        throw;
    }
    return g;
}

vb_genomes *vb_genomes_from_memory_impl(const char *const *names, const char *const *seqs, const uint64_t *lens,
                                        uint32_t n)
{
    auto *g = new vb_genomes();
    g->offset.push_back(0);
    uint64_t total = 0;
    for (uint32_t i = 0; i < n; ++i) total += lens[i];
    g->bases.reserve(total);
    for (uint32_t i = 0; i < n; ++i) {
        g->names.emplace_back(names[i]);
        g->bases.append(seqs[i], lens[i]);
        g->offset.push_back(g->bases.size());
    }
    return g;
}
