// Text formats of the hot path's outputs (host side): the kmer-db filter file and LZ-ANI's ani.tsv / ids.tsv.
// Number formatting reproduces the reference byte for byte:
//   kmer-db  conversion.h:167-219      Double2PChar(v, 6): x = (u64)(v*1e6 + 0.5), printed d.dddddd
//   lz-ani   numeric_conversions.h:229-299,342-390  real_to_pchar: shortest round-trip decimal (dragonbox), then
//            round-half-up to `prec` significant digits, no zero padding; scientific when the exponent is large.
#include <algorithm>
#include <charconv>
#include <cmath>
#include <fstream>
#include <memory>
#include <thread>

#include "vb_internal.h"

// Output file whose every write is checked: a full disk or an I/O error throws VB_ERR_IO and removes the partial
// file (a silently truncated filter would silently drop candidate pairs in the align stage).
struct OutFile {
    FILE *f = nullptr;
    std::string path;
    OutFile(const char *p, const char *what) : path(p)
    {
        f = fopen(p, "wb");
        if (!f) throw vb_error(VB_ERR_IO, std::string(what) + p);
    }
    OutFile(const OutFile &) = delete;
    OutFile &operator=(const OutFile &) = delete;
    void fail()
    {
        if (f) { fclose(f); f = nullptr; }
        remove(path.c_str());
        throw vb_error(VB_ERR_IO, "write error on " + path);
    }
    void write(const char *data, size_t n) { if (n && fwrite(data, 1, n, f) != n) fail(); }
    void write(const std::string &s) { write(s.data(), s.size()); }
    void close() { FILE *g = f; f = nullptr; if (g && fclose(g) != 0) { remove(path.c_str()); throw vb_error(VB_ERR_IO, "write error on " + path); } }
    ~OutFile() { if (f) { fclose(f); remove(path.c_str()); } }       // left open = an exception is in flight: drop the partial file
};

int vb_fmt_fixed6(double v, char *out)
{
    char *p = out;
    if (v < 0) { *p++ = '-'; v = -v; }
    uint64_t x = (uint64_t)(v * 1000000.0 + 0.5);
    p += sprintf(p, "%llu.%06llu", (unsigned long long)(x / 1000000ULL), (unsigned long long)(x % 1000000ULL));
    return (int)(p - out);
}

static int put_u64(uint64_t v, char *out)
{
    auto r = std::to_chars(out, out + 24, v);
    return (int)(r.ptr - out);
}

int vb_fmt_real(double v, int prec, char *out)
{
    char *p = out;
    if (v == 0) { *p++ = '0'; return 1; }
    if (std::isnan(v)) { memcpy(p, "nan", 3); return 3; }
    if (std::isinf(v)) { if (v < 0) { memcpy(p, "-inf", 4); return 4; } memcpy(p, "inf", 3); return 3; }
    prec = std::clamp(prec, 1, 15);
    // shortest round-trip digits: d.ddddde[+-]XX  (std::to_chars without precision is shortest, like dragonbox)
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::scientific);
    *r.ptr = 0;
    const char *s = buf;
    bool neg = false;
    if (*s == '-') { neg = true; ++s; }
    uint64_t sig = 0;
    int n_dig = 0;
    const char *e = strchr(s, 'e');
    for (const char *q = s; q < e; ++q)
        if (*q != '.') { sig = sig * 10 + (uint64_t)(*q - '0'); ++n_dig; }
    int exponent = atoi(e + 1) - (n_dig - 1);       // value = sig * 10^exponent
    while (sig % 10 == 0) { sig /= 10; ++exponent; --n_dig; }
    static const uint64_t p10[] = {1ULL, 10ULL, 100ULL, 1000ULL, 10000ULL, 100000ULL, 1000000ULL, 10000000ULL,
                                   100000000ULL, 1000000000ULL, 10000000000ULL, 100000000000ULL, 1000000000000ULL,
                                   10000000000000ULL, 100000000000000ULL, 1000000000000000ULL, 10000000000000000ULL,
                                   100000000000000000ULL, 1000000000000000000ULL};
    if (n_dig > prec) {
        sig += p10[n_dig - prec] / 2;
        sig /= p10[n_dig - prec];
        exponent += n_dig - prec;
        n_dig = prec;
        if (sig >= p10[prec]) { sig /= 10; ++exponent; }
    }
    if (neg) *p++ = '-';
    char digs[32];
    int nd = put_u64(sig, digs);        // nd == n_dig
    if (exponent == 0) {
        memcpy(p, digs, nd); p += nd;
    } else if (exponent > 0 || -exponent >= n_dig + 4) {
        if (n_dig == 1) *p++ = digs[0];
        else {
            *p++ = digs[0]; *p++ = '.';
            memcpy(p, digs + 1, nd - 1); p += nd - 1;
            exponent += n_dig - 1;
        }
        *p++ = 'e';
        int ex = exponent;
        if (ex < 0) { *p++ = '-'; ex = -ex; } else *p++ = '+';
        p += sprintf(p, "%02d", ex);
    } else if (-exponent < n_dig) {
        int ip = n_dig + exponent;
        memcpy(p, digs, ip); p += ip;
        *p++ = '.';
        memcpy(p, digs + ip, nd - ip); p += nd - ip;
    } else {
        *p++ = '0'; *p++ = '.';
        for (int i = 0; i < -exponent - n_dig; ++i) *p++ = '0';
        memcpy(p, digs, nd); p += nd;
    }
    return (int)(p - out);
}

double vb_ani_shorter(uint32_t common, uint32_t cnt1, uint32_t cnt2, int k)
{
    double j = (double)common / std::min(cnt1, cnt2);
    double d = (j == 0) ? 1.0 : (-1.0 / k) * std::log((2 * j) / (j + 1));
    return 1.0 - d;
}

// ---------------------------------------------------------------------------------------------------------------
// filter file
// ---------------------------------------------------------------------------------------------------------------
void vb_write_filter_impl(const vb_genomes *g, const vb_pairs *pr, const char *path)
{
    OutFile f(path, "Cannot open output file: ");
    std::string out;
    out.reserve(1 << 20);
    char num[64];
    // header: "kmer-length: K fraction: F ,name1,name2,...,"   (console_distance.cpp:37-42; F printed by ostream<<double)
    snprintf(num, sizeof(num), "%g", pr->kmers_fraction);
    out += "kmer-length: " + std::to_string(pr->k) + " fraction: " + num + " ,";
    for (auto &nm : g->names) { out += nm; out += ','; }
    out += '\n';
    uint64_t e = 0;
    for (uint32_t r = 0; r < g->count(); ++r) {
        out += g->names[r];
        out += ',';
        for (; e < pr->n_pairs && pr->row[e] == r; ++e) {
            int n = put_u64((uint64_t)pr->col[e] + 1, num);       // 1-based column ids (array.h:625-637)
            out.append(num, n);
            out += ':';
            n = vb_fmt_fixed6(pr->ani[e], num);
            out.append(num, n);
            out += ',';
        }
        out += '\n';
        if (out.size() > (1 << 20)) { f.write(out); out.clear(); }
    }
    f.write(out);
    f.close();
}

// kmer-db `all2all-sp -sample-rows ani-shorter:N` (vclust --max-seqs N).  Every pair (i, j), i > j, that passed the -min
// filters is offered to row i as item j and -- through the transposed pass -- to row j as item i (array.h:450-543); a
// row keeps its N best items, where the heap of sampler.h:45-66 always evicts the item that is worst by
// (score ascending, item id descending), i.e. the survivors are the top N by (score desc, item asc) whatever the
// insertion order; rows are written with the items ascending (sampler.h:123-139).  The result therefore has entries
// on both sides of the diagonal, and a pair kept by both of its rows appears twice.
void vb_sample_rows(uint32_t n_genomes, uint32_t max_items, std::vector<uint32_t> &row, std::vector<uint32_t> &col,
                    std::vector<uint32_t> &common, std::vector<double> &ani)
{
    const size_t n = row.size();
    std::vector<uint64_t> start((size_t)n_genomes + 1, 0);
    for (size_t i = 0; i < n; ++i) { start[row[i] + 1]++; start[col[i] + 1]++; }
    for (uint32_t r = 0; r < n_genomes; ++r) start[r + 1] += start[r];
    struct Item { uint32_t item, common; double score; };
    std::vector<Item> items(2 * n);
    std::vector<uint64_t> fill(start.begin(), start.end() - 1);
    for (size_t i = 0; i < n; ++i) {
        items[fill[row[i]]++] = {col[i], common[i], ani[i]};
        items[fill[col[i]]++] = {row[i], common[i], ani[i]};
    }
    row.clear(); col.clear(); common.clear(); ani.clear();
    for (uint32_t r = 0; r < n_genomes; ++r) {
        Item *b = items.data() + start[r], *e = items.data() + start[r + 1];
        if ((uint64_t)(e - b) > max_items) {
            std::partial_sort(b, b + max_items, e, [](const Item &x, const Item &y) {
                return x.score != y.score ? x.score > y.score : x.item < y.item;
            });
            e = b + max_items;
        }
        std::sort(b, e, [](const Item &x, const Item &y) { return x.item < y.item; });
        for (Item *p = b; p != e; ++p) { row.push_back(r); col.push_back(p->item); common.push_back(p->common); ani.push_back(p->score); }
    }
}

void vb_pairs_free_impl(vb_pairs *p);

static std::vector<std::string> split_keep(const std::string &s, char sep)
{   // lz-ani utils.cpp:15-36: empty middle tokens kept, empty trailing token dropped
    std::vector<std::string> parts;
    std::string cur;
    for (char c : s) {
        if (c == sep) { parts.push_back(cur); cur.clear(); }
        else cur.push_back(c);
    }
    if (!cur.empty()) parts.push_back(cur);
    return parts;
}

vb_pairs *vb_pairs_alloc(uint64_t n, uint32_t n_genomes)
{
    auto *box = (vb_pairs_box *)calloc(1, sizeof(vb_pairs_box));
    if (!box) throw vb_error(VB_ERR_MEM, "out of host memory");
    box->uid = vb_next_uid();
    vb_pairs *p = &box->pub;
    p->n_pairs = n;
    p->n_genomes = n_genomes;
    p->row = (uint32_t *)malloc(sizeof(uint32_t) * std::max<uint64_t>(n, 1));
    p->col = (uint32_t *)malloc(sizeof(uint32_t) * std::max<uint64_t>(n, 1));
    p->common = (uint32_t *)calloc(std::max<uint64_t>(n, 1), sizeof(uint32_t));
    p->ani = (double *)calloc(std::max<uint64_t>(n, 1), sizeof(double));
    p->total_kmers = (uint32_t *)calloc(std::max<uint32_t>(n_genomes, 1), sizeof(uint32_t));
    if (!p->row || !p->col || !p->common || !p->ani || !p->total_kmers) {
        vb_pairs_free_impl(p);
        throw vb_error(VB_ERR_MEM, "out of host memory for " + std::to_string(n) + " pairs");
    }
    return p;
}

void vb_pairs_free_impl(vb_pairs *p)
{
    if (!p) return;
    free(p->row); free(p->col); free(p->common); free(p->ani); free(p->total_kmers);
    free(p);                      // == the vb_pairs_box (pub is its first member)
}

void vb_parallel_for(uint64_t n, uint64_t min_per_thread, unsigned max_threads, const std::function<void(uint64_t, uint64_t)> &fn)
{
    if (!n) return;
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const unsigned t = (unsigned)std::min<uint64_t>(std::min(hw, std::max(1u, max_threads)), std::max<uint64_t>(1, n / std::max<uint64_t>(1, min_per_thread)));
    if (t <= 1) { fn(0, n); return; }
    std::vector<std::thread> pool;
    for (unsigned i = 1; i < t; ++i) pool.emplace_back([&, i]() { fn(n * i / t, n * (i + 1) / t); });
    fn(0, n / t);
    for (auto &th : pool) th.join();
}

vb_pairs *vb_read_filter_impl(const char *path, double thr, const vb_genomes *g)
{
    std::ifstream in(path, std::ios::binary);
    if (!in) throw vb_error(VB_ERR_IO, std::string("Cannot open file: ") + path);
    std::string line;
    auto getline_cr = [&](std::string &l) {
        if (!std::getline(in, l)) return false;
        if (!l.empty() && l.back() == '\r') l.pop_back();
        return true;
    };
    if (!getline_cr(line)) throw vb_error(VB_ERR_IO, "Incorrect kmer-db filter file");
    auto names = split_keep(line, ',');
    if (names.size() <= 2) throw vb_error(VB_ERR_IO, "Incorrect kmer-db filter file");
    names.erase(names.begin());
    if (names.size() != g->names.size() || names != g->names)
        throw vb_error(VB_ERR_MISMATCH, names.size() != g->names.size()
                                            ? "Input sequences and filter sequences sets are of different size!"
                                            : "Input sequences and filter sequences are different!");
    std::vector<uint32_t> rows, cols;
    std::vector<double> vals;
    uint32_t row = 0;
    while (getline_cr(line)) {
        if (line.size() <= 2) continue;                 // filter.cpp:107-111: row id NOT advanced
        // tokens between commas, the first one is the name; a token counts iff it splits at ':' into exactly two parts
        // under lz-ani's split() (utils.cpp:15-36: an empty LAST part is dropped, so "3:" has one part and is skipped,
        // ":0.9" has two); parsed in place -- one allocation-free pass instead of a vector of strings per line
        const char *b = line.c_str(), *end = b + line.size();
        const char *tok = (const char *)memchr(b, ',', (size_t)(end - b));
        while (tok) {
            const char *t0 = tok + 1;
            const char *t1 = (const char *)memchr(t0, ',', (size_t)(end - t0));
            const char *te = t1 ? t1 : end;
            const char *c = (const char *)memchr(t0, ':', (size_t)(te - t0));
            if (c && c + 1 < te && !memchr(c + 1, ':', (size_t)(te - c - 1))) {
                const double v = strtod(c + 1, nullptr);        // stops at the ',' / end of line
                if (v >= thr) {
                    rows.push_back(row);
                    cols.push_back((uint32_t)(atoi(t0) - 1));
                    vals.push_back(v);
                }
            }
            tok = t1;
        }
        ++row;
    }
    vb_pairs *p = vb_pairs_alloc(rows.size(), g->count());
    for (size_t i = 0; i < rows.size(); ++i) {
        if (rows[i] >= g->count() || cols[i] >= g->count()) {
            vb_pairs_free_impl(p);
            throw vb_error(VB_ERR_IO, "Incorrect kmer-db filter file: id out of range");
        }
        p->row[i] = rows[i]; p->col[i] = cols[i]; p->ani[i] = vals[i];
    }
    return p;
}

// ---------------------------------------------------------------------------------------------------------------
// ani.tsv / ids.tsv   (lz_matcher.cpp:280-579)
// ---------------------------------------------------------------------------------------------------------------
void vb_write_ani_impl(const vb_genomes *g, const vb_align_out *res, const char *ani_path, const char *ids_path,
                       const char *const *columns, int n_columns, const double out_filters[5])
{
    enum { C_QUERY, C_REFERENCE, C_QIDX, C_RIDX, C_QLEN, C_RLEN, C_TANI, C_GANI, C_ANI, C_QCOV, C_RCOV, C_LEN_RATIO,
           C_NT_MATCH, C_NT_MISMATCH, C_NUM_ALNS };
    static const char *col_names[] = {"query", "reference", "qidx", "ridx", "qlen", "rlen", "tani", "gani", "ani",
                                      "qcov", "rcov", "len_ratio", "nt_match", "nt_mismatch", "num_alns"};
    std::vector<int> cols;
    for (int i = 0; i < n_columns; ++i) {
        int id = -1;
        for (int j = 0; j < 15; ++j) if (!strcmp(columns[i], col_names[j])) id = j;
        if (id < 0) throw vb_error(VB_ERR_ARG, std::string("Unknown output-format component: ") + columns[i]);
        cols.push_back(id);
    }
    const uint32_t n = res->n_genomes;
    if (n != g->count()) throw vb_error(VB_ERR_ARG, "vb_write_ani: result and genome set differ in size");
    bool any_filter = false;
    double flt[5] = {0, 0, 0, 0, 0};
    if (out_filters) for (int i = 0; i < 5; ++i) { flt[i] = out_filters[i]; any_filter |= out_filters[i] != 0; }

    {   // ids file, in LZ-ANI order
        OutFile f(ids_path, "Cannot open output file: ");
        std::string ids = "id\tseq_len\tno_parts\n";
        char num[32];
        for (uint32_t i = 0; i < n; ++i) {
            ids += g->names[res->order[i]];
            ids += '\t';
            ids.append(num, put_u64(g->length(res->order[i]), num));
            ids += "\t1\n";
            if (ids.size() > (1 << 20)) { f.write(ids); ids.clear(); }
        }
        f.write(ids);
        f.close();
    }
    OutFile f(ani_path, "Cannot open output file: ");
    std::string out;
    out.reserve(4 << 20);
    for (size_t i = 0; i < cols.size(); ++i) { if (i) out += '\t'; out += col_names[cols[i]]; }
    out += '\n';

    // CSR over ref rows (res is sorted by (ref, qry))
    std::vector<uint64_t> start(n + 1, 0);
    for (uint64_t i = 0; i < res->n; ++i) start[res->ref[i] + 1]++;
    for (uint32_t i = 0; i < n; ++i) start[i + 1] += start[i];
    auto find_first = [&](uint32_t r, uint32_t q) -> int64_t {   // lower_bound on qry within row r
        uint64_t lo = start[r], hi = start[r + 1];
        while (lo < hi) { uint64_t mid = (lo + hi) / 2; if (res->qry[mid] < q) lo = mid + 1; else hi = mid; }
        return (lo < start[r + 1] && res->qry[lo] == q) ? (int64_t)lo : -1;
    };
    // rows [a_lo, a_hi) formatted into `dst`; big results are formatted by several host threads on disjoint row ranges
    // (equal numbers of directed pairs) and written in row order -- number formatting is the whole cost of this call
    auto format_rows = [&](uint32_t a_lo, uint32_t a_hi, std::string &dst) {
    char num[64];
    for (uint32_t a = a_lo; a < a_hi; ++a) {
        for (uint64_t qi = start[a]; qi < start[a + 1]; ++qi) {
            uint32_t b = res->qry[qi];
            if (a >= b) continue;
            int64_t pi = find_first(b, a);
            if (pi < 0) continue;       // cannot happen for symmetric pair lists
            uint32_t ids[2] = {a, b};
            uint64_t len[2] = {g->length(res->order[b]), g->length(res->order[a])};
            int32_t mat[2] = {res->sym_in_matches[qi], res->sym_in_matches[pi]};
            int32_t lit[2] = {res->sym_in_literals[qi], res->sym_in_literals[pi]};
            int32_t reg[2] = {res->no_components[qi], res->no_components[pi]};
            double tani = (double)(mat[0] + mat[1]) / (double)(uint32_t)(len[0] + len[1]);
            double gani[2] = {(double)mat[0] / (uint32_t)len[0], (double)mat[1] / (uint32_t)len[1]};
            double ani[2] = {mat[0] + lit[0] != 0 ? (double)mat[0] / (mat[0] + lit[0]) : 0,
                             mat[1] + lit[1] != 0 ? (double)mat[1] / (mat[1] + lit[1]) : 0};
            double cov[2] = {(double)(mat[0] + lit[0]) / (uint32_t)len[0], (double)(mat[1] + lit[1]) / (uint32_t)len[1]};
            for (int i = 0; i < 2; ++i) {
                if (any_filter) {
                    if (gani[i] < flt[1] || ani[i] < flt[2] || tani < flt[0] || cov[i] < flt[3] || cov[!i] < flt[4])
                        continue;
                }
                for (size_t c = 0; c < cols.size(); ++c) {
                    int nn = 0;
                    switch (cols[c]) {
                    case C_RIDX: nn = put_u64(ids[i], num); break;
                    case C_QIDX: nn = put_u64(ids[!i], num); break;
                    case C_REFERENCE: dst += g->names[res->order[ids[i]]]; break;
                    case C_QUERY: dst += g->names[res->order[ids[!i]]]; break;
                    case C_QCOV: nn = vb_fmt_real(cov[i], 6, num); break;
                    case C_RCOV: nn = vb_fmt_real(cov[!i], 6, num); break;
                    case C_GANI: nn = vb_fmt_real(gani[i], 6, num); break;
                    case C_ANI: nn = vb_fmt_real(ani[i], 6, num); break;
                    case C_TANI: nn = vb_fmt_real(tani, 6, num); break;
                    case C_RLEN: nn = put_u64(len[!i], num); break;
                    case C_QLEN: nn = put_u64(len[i], num); break;
                    case C_NUM_ALNS: nn = put_u64((uint64_t)reg[i], num); break;
                    case C_NT_MATCH: nn = put_u64((uint64_t)mat[i], num); break;
                    case C_NT_MISMATCH: nn = put_u64((uint64_t)lit[i], num); break;
                    case C_LEN_RATIO:
                        if (len[0] && len[1]) {
                            double lr = len[i] < len[!i] ? (double)(uint32_t)len[i] / (uint32_t)len[!i]
                                                         : (double)(uint32_t)len[!i] / (uint32_t)len[i];
                            nn = vb_fmt_real(lr, 4, num);
                        } else { num[0] = '0'; nn = 1; }
                        break;
                    }
                    dst.append(num, nn);
                    dst += (c + 1 < cols.size()) ? '\t' : '\n';
                }
                if (cols.empty()) dst += '\n';
            }
        }
    }
    };
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const char *thr_env = getenv("VB_WRITE_THREADS");                  // test hook
    const unsigned n_thr = thr_env ? (unsigned)std::clamp(atoi(thr_env), 1, 64) : (res->n < 50000 ? 1u : std::min(hw, 16u));
    if (n_thr == 1) {
        format_rows(0, n, out);
        f.write(out);
    } else {
        std::vector<uint32_t> cut(n_thr + 1, n);
        cut[0] = 0;
        for (unsigned t = 1; t < n_thr; ++t) {
            const uint64_t want = res->n * (uint64_t)t / n_thr;
            cut[t] = (uint32_t)(std::lower_bound(start.begin(), start.end(), want) - start.begin());
            cut[t] = std::min(std::max(cut[t], cut[t - 1]), n);
        }
        std::vector<std::string> parts(n_thr);
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < n_thr; ++t)
            pool.emplace_back([&, t]() { parts[t].reserve(1 << 20); format_rows(cut[t], cut[t + 1], parts[t]); });
        for (auto &th : pool) th.join();
        f.write(out);                                          // header
        for (auto &part : parts) f.write(part);
    }
    f.close();
}

// lz-ani lz_matcher.cpp:102-169 (store_alignment): one line per region; coordinates 1-based inclusive; a region on the
// reverse-complement half of the reference text is reported on the forward genome with rstart > rend (:158-162).
void vb_write_aln_impl(const vb_genomes *g, const vb_regions *r, const char *path, const double out_filters[5])
{
    OutFile f(path, "Cannot open output file for alignment storage: ");
    std::string out;
    out.reserve(4 << 20);
    out += "query\treference\tpident\talnlen\tqstart\tqend\trstart\trend\tnt_match\tnt_mismatch\n";
    bool any_filter = false;
    double flt[5] = {0, 0, 0, 0, 0};
    if (out_filters) for (int i = 0; i < 5; ++i) { flt[i] = out_filters[i]; any_filter |= out_filters[i] != 0; }
    const uint32_t ng = g->count();
    char num[64];
    for (uint64_t b = 0; b < r->n;) {
        uint64_t e = b;
        while (e < r->n && r->ref[e] == r->ref[b] && r->qry[e] == r->qry[b]) ++e;      // the regions of one directed pair
        const uint32_t ref = r->ref[b], qry = r->qry[b];
        if (ref >= ng || qry >= ng) throw vb_error(VB_ERR_ARG, "vb_write_aln: genome id out of range");
        const int seq1_len = (int)g->length(ref), seq2_len = (int)g->length(qry);
        const int rc_correction = 2 * seq1_len + 2 * r->mrd + 1;
        bool keep = true;
        if (any_filter) {
            int32_t si_mat = 0, si_lit = 0;
            for (uint64_t i = b; i < e; ++i) { si_mat += r->matches[i]; si_lit += r->mismatches[i]; }
            const double global_ani = (double)si_mat / seq2_len;
            const double local_ani = si_mat + si_lit != 0 ? (double)si_mat / (si_mat + si_lit) : 0;
            const double qcov = (double)(si_mat + si_lit) / seq2_len;
            keep = !(global_ani < flt[1] || local_ani < flt[2] || qcov < flt[3]);
        }
        if (keep)
            for (uint64_t i = b; i < e; ++i) {
                const int len = r->q_end[i] - r->q_start[i];
                out += g->names[qry]; out += '\t';
                out += g->names[ref]; out += '\t';
                out.append(num, vb_fmt_real(100.0 * r->matches[i] / len, 6, num)); out += '\t';
                out.append(num, put_u64((uint64_t)len, num)); out += '\t';
                out.append(num, put_u64((uint64_t)(1 + r->q_start[i]), num)); out += '\t';
                out.append(num, put_u64((uint64_t)r->q_end[i], num)); out += '\t';
                if (r->r_start[i] < seq1_len) {
                    out.append(num, put_u64((uint64_t)(1 + r->r_start[i]), num)); out += '\t';
                    out.append(num, put_u64((uint64_t)r->r_end[i], num)); out += '\t';
                } else {
                    out.append(num, put_u64((uint64_t)(rc_correction - (1 + r->r_start[i])), num)); out += '\t';
                    out.append(num, put_u64((uint64_t)(rc_correction - r->r_end[i]), num)); out += '\t';
                }
                out.append(num, put_u64((uint64_t)r->matches[i], num)); out += '\t';
                out.append(num, put_u64((uint64_t)r->mismatches[i], num)); out += '\n';
                if (out.size() > (3u << 20)) { f.write(out); out.clear(); }
            }
        b = e;
    }
    f.write(out);
    f.close();
}
