// FASTA ingest for both stages (host side).  Mirrors the *reading rules* of the two reference tools, which
// differ in corner cases; see vb_fasta_flavor in include/vclust_b200.h.
//   kmer-db : genome_input_file.h:80-136 (whole-file read, gzip by magic), :287-338 (record splitting),
//             loader_ex.cpp:150-257 (directory mode: one sample per file, named by file name)
//   lz-ani  : seq_reservoir.cpp:90-153 (directory mode), :156-210 (multi-FASTA), file_wrapper.h:917-950 (getline)
#include <zlib.h>

#include <algorithm>
#include <filesystem>

#include "vb_internal.h"

namespace {

std::string read_whole(const char *path)
{
    gzFile f = gzopen(path, "rb");          // transparently reads plain files too
    if (!f) throw vb_error(VB_ERR_IO, std::string("Cannot open file: ") + path);
    gzbuffer(f, 1 << 20);
    std::string data;
    std::vector<char> buf(8 << 20);
    for (;;) {
        int n = gzread(f, buf.data(), (unsigned)buf.size());
        if (n < 0) { gzclose(f); throw vb_error(VB_ERR_IO, std::string("Cannot read file: ") + path); }
        if (n == 0) break;
        data.append(buf.data(), (size_t)n);
    }
    gzclose(f);
    return data;
}

struct record { std::string name; std::string seq; };

// kmer-db rule: a record starts at every '>' byte; header ends at '\n' (a preceding '\r' is dropped) and is cut at
// the first space; the sequence is everything up to the next '>' without '\n' and '\r'.
void split_kmerdb(const std::string &d, std::vector<record> &out)
{
    size_t pos = d.find('>');
    while (pos != std::string::npos) {
        size_t eol = d.find('\n', pos);
        if (eol == std::string::npos) eol = d.size();
        size_t hend = eol;
        if (hend > pos + 1 && d[hend - 1] == '\r') --hend;
        record r;
        r.name.assign(d, pos + 1, hend - pos - 1);
        size_t sp = r.name.find(' ');
        if (sp != std::string::npos) r.name.resize(sp);
        size_t nxt = d.find('>', eol);
        size_t bend = (nxt == std::string::npos) ? d.size() : nxt;
        size_t bbeg = std::min(eol + 1, bend);
        r.seq.reserve(bend - bbeg);
        for (size_t i = bbeg; i < bend; ++i) {
            char c = d[i];
            if (c != '\n' && c != '\r') r.seq.push_back(c);
        }
        out.push_back(std::move(r));
        pos = nxt;
    }
}

// lz-ani rule: line based; '>' counts only in column 0; every line loses one trailing '\r'; empty lines are skipped.
// keep_unterminated_tail: directory mode keeps a last line without '\n', multi-FASTA mode drops it.
// dir_mode additionally keeps records with an empty name and header-less leading sequence (load_fasta ignores names).
void split_lzani(const std::string &d, bool dir_mode, std::vector<record> &out)
{
    const bool keep_unterminated_tail = dir_mode;
    record cur;
    bool have = false;
    size_t pos = 0;
    auto handle = [&](size_t b, size_t e) {
        if (e > b && d[e - 1] == '\r') --e;
        if (e == b) return;
        if (d[b] == '>') {
            if ((have && !cur.name.empty()) || (dir_mode && (have || !cur.seq.empty()))) out.push_back(std::move(cur));
            cur = record();
            cur.name.assign(d, b + 1, e - b - 1);
            have = true;
        } else
            cur.seq.append(d, b, e - b);
    };
    while (pos < d.size()) {
        size_t eol = d.find('\n', pos);
        if (eol == std::string::npos) {
            if (keep_unterminated_tail) handle(pos, d.size());
            break;
        }
        handle(pos, eol);
        pos = eol + 1;
    }
    if ((have && !cur.name.empty()) || (dir_mode && (have || !cur.seq.empty()))) out.push_back(std::move(cur));
    for (auto &r : out) {
        size_t sp = r.name.find(' ');
        if (sp != std::string::npos) r.name.resize(sp);
    }
}

void push_genome(vb_genomes *g, const std::string &name, const std::string &seq)
{
    g->names.push_back(name);
    g->bases.insert(g->bases.end(), seq.begin(), seq.end());
    g->offset.push_back(g->bases.size());
}

}  // namespace

vb_genomes *vb_genomes_load_impl(const char *const *paths, int n_paths, int multisample, vb_fasta_flavor flavor,
                                 int sep_len)
{
    auto *g = new vb_genomes();
    g->flavor = flavor;
    g->offset.push_back(0);
    try {
        for (int i = 0; i < n_paths; ++i) {
            std::string data = read_whole(paths[i]);
            std::vector<record> recs;
            if (flavor == VB_FASTA_KMERDB) split_kmerdb(data, recs);
            else split_lzani(data, !multisample, recs);
            if (multisample) {
                for (auto &r : recs) push_genome(g, r.name, r.seq);
            } else {
                // one genome per file.  kmer-db: records pooled, k-mers must not span records -> one 'N' between
                // them.  lz-ani: records joined by sep_len N symbols whenever the sequence so far is non-empty.
                std::string joined;
                size_t sep = (flavor == VB_FASTA_KMERDB) ? 1 : (size_t)std::max(sep_len, 0);
                bool first = true;
                for (auto &r : recs) {
                    if (flavor == VB_FASTA_KMERDB) { if (!first) joined.append(sep, 'N'); }
                    else if (!joined.empty()) joined.append(sep, 'N');
                    joined += r.seq;
                    first = false;
                }
                push_genome(g, std::filesystem::path(paths[i]).filename().string(), joined);
            }
        }
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
        g->bases.insert(g->bases.end(), seqs[i], seqs[i] + lens[i]);
        g->offset.push_back(g->bases.size());
    }
    return g;
}
