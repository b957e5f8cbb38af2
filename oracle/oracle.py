"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes front-end of the plain-C restatements (kmer_oracle.c, lz_oracle.c) plus numpy/pure-Python restatements of
the reference's host-side text handling (FASTA reading rules of both tools, filter-file and ani.tsv formatting).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this module;
the product (vclust_b200/) never does.

Reference citations are relative to /root/reference/3rd_party/ (K = kmer-db/src, L = lz-ani/src).
"""
from __future__ import annotations

import ctypes as C
import gzip
import os
import subprocess
from decimal import Decimal
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "liboracle.so"
REF_DIR = HERE / "_ref"
SOURCES = [HERE / "kmer_oracle.c", HERE / "lz_oracle.c"]


def build(force: bool = False) -> Path:
    """gcc the C restatements into oracle/liboracle.so (and, when /root/reference is present, oracle/_ref)."""
    newest = max(s.stat().st_mtime for s in SOURCES)
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < newest:
        cmd = ["gcc", "-O2", "-std=c99", "-shared", "-fPIC", "-o", str(LIB_PATH)] + [str(s) for s in SOURCES] + ["-lm"]
        subprocess.run(cmd, check=True)
    return LIB_PATH


def build_ref() -> bool:
    """Build the unmodified reference binaries into oracle/_ref (no-op when the sources are absent)."""
    subprocess.run(["bash", str(HERE / "build_ref.sh")], check=True)
    if ref_available():
        from . import make_dropin        # the reference's vclust.py + test.py with the GPU patch applied (tests/test_dropin.py)
        make_dropin.make()
    return ref_available()


def ref_available() -> bool:
    return (REF_DIR / "kmer-db").exists() and (REF_DIR / "lz-ani").exists()


class LzParams(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("mal", "msl", "mrd", "mqd", "reg", "aw", "am", "ar")]

    @classmethod
    def default(cls, **kw):
        d = dict(mal=11, msl=7, mrd=40, mqd=40, reg=35, aw=15, am=7, ar=3)
        d.update(kw)
        return cls(**d)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(LIB_PATH))
        L.kmo_extract.restype = C.c_size_t
        L.kmo_extract.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_double, C.c_void_p]
        L.kmo_sort_unique.restype = C.c_size_t
        L.kmo_sort_unique.argtypes = [C.c_void_p, C.c_size_t]
        L.kmo_minhash.restype = C.c_uint64
        L.kmo_minhash.argtypes = [C.c_uint64, C.c_int]
        L.kmo_common.restype = C.c_int64
        L.kmo_common.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.POINTER(C.POINTER(C.c_uint32))] * 3
        L.kmo_free.argtypes = [C.c_void_p]
        L.kmo_ani_shorter.restype = C.c_double
        L.kmo_ani_shorter.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
        L.lzo_run_pairs.restype = None
        L.lzo_run_pairs.argtypes = [C.POINTER(LzParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                    C.c_void_p]
        _lib = L
    return _lib


# ----------------------------------------------------------------------------------------------------------------
# FASTA reading rules
# ----------------------------------------------------------------------------------------------------------------
def _read_bytes(path) -> bytes:
    raw = Path(path).read_bytes()
    if raw[:2] == b"\x1f\x8b":
        raw = gzip.decompress(raw)
    return raw


def read_records_kmerdb(path):
    """K/genome_input_file.h:287-338: split at every '>', name = header up to first space (and before \\r/\\n),
    sequence = all bytes up to the next '>' with \\n and \\r removed."""
    data = _read_bytes(path)
    recs = []
    pos = data.find(b">")
    while pos >= 0:
        eol = data.find(b"\n", pos)
        if eol < 0:
            eol = len(data)
        header = data[pos + 1:eol].rstrip(b"\r")
        sp = header.find(b" ")
        if sp >= 0:
            header = header[:sp]
        nxt = data.find(b">", eol)
        body = data[eol + 1:nxt if nxt >= 0 else len(data)]
        recs.append((header.decode(), body.replace(b"\n", b"").replace(b"\r", b"")))
        pos = nxt
    return recs


def read_records_lzani(path, multifasta: bool = True):
    """L/seq_reservoir.cpp:156-210 (multi-FASTA) -- line based; a '>' only counts at the start of a line; each line
    loses one trailing \\r; empty lines skipped; an unterminated last line is dropped (getline returns <0 at EOF,
    file_wrapper.h:917-950 + seq_reservoir.cpp:177)."""
    data = _read_bytes(path)
    lines = data.split(b"\n")
    last_unterminated = lines.pop()          # text after the final \n ('' when the file ends with \n)
    recs, name, seq = [], None, []
    for ln in lines:
        if ln.endswith(b"\r"):
            ln = ln[:-1]
        if not ln:
            continue
        if ln[:1] == b">":
            if name:
                recs.append((name, b"".join(seq)))
            name = ln[1:]
            seq = []
        else:
            seq.append(ln)
    if not multifasta and last_unterminated:
        ln = last_unterminated[:-1] if last_unterminated.endswith(b"\r") else last_unterminated
        if ln and ln[:1] != b">":
            seq.append(ln)
    if name:
        recs.append((name, b"".join(seq)))
    out = []
    for nm, s in recs:
        sp = nm.find(b" ")
        out.append(((nm[:sp] if sp >= 0 else nm).decode(), s))
    return out


_LZ_CODE = np.full(256, 5, dtype=np.uint8)          # L/seq_reservoir.h:243-247
for _i, _ch in enumerate("ACGT"):
    _LZ_CODE[ord(_ch)] = _i
    _LZ_CODE[ord(_ch.lower())] = _i


def lz_codes(seq: bytes) -> np.ndarray:
    return _LZ_CODE[np.frombuffer(seq, dtype=np.uint8)]


def load_genomes_lzani(paths, multifasta: bool, mrd: int = 40):
    """Names + code arrays in INPUT order.  Directory mode (one genome per file, L/seq_reservoir.cpp:90-153): contigs are
    joined by mrd 'N' codes and the genome is named after the file (with extension)."""
    names, codes = [], []
    if multifasta:
        for p in paths:
            for nm, s in read_records_lzani(p, True):
                names.append(nm)
                codes.append(lz_codes(s))
    else:
        sep = np.full(mrd, 5, dtype=np.uint8)
        for p in paths:
            parts = [lz_codes(s) for _, s in read_records_lzani(p, False)]
            joined = []
            for i, c in enumerate(parts):
                if i and sum(x.size for x in joined):
                    joined.append(sep)
                joined.append(c)
            names.append(Path(p).name)
            codes.append(np.concatenate(joined) if joined else np.zeros(0, np.uint8))
    return names, codes


def lz_order(names, lens):
    """L/seq_reservoir.cpp:215-251: stable sort by (len - 2*no_parts) desc then name asc; no_parts is always 1."""
    idx = sorted(range(len(names)), key=lambda i: (-(lens[i] - 2), names[i].encode()))
    return idx


# ----------------------------------------------------------------------------------------------------------------
# prefilter (kmer-db) restatement
# ----------------------------------------------------------------------------------------------------------------
def kmer_sets(samples, k: int, fraction: float):
    """samples: list of lists of byte strings (the records pooled into one sample).  Returns list of sorted
    unique uint64 arrays."""
    L = lib()
    out = []
    for recs in samples:
        total = sum(len(r) for r in recs)
        buf = np.empty(max(total, 1), dtype=np.uint64)
        n = 0
        for r in recs:
            n += L.kmo_extract(r, len(r), k, float(fraction), buf[n:].ctypes.data)
        n = L.kmo_sort_unique(buf.ctypes.data, n)
        out.append(buf[:n].copy())
    return out


def common_matrix(sets):
    L = lib()
    off = np.zeros(len(sets) + 1, dtype=np.int64)
    off[1:] = np.cumsum([s.size for s in sets])
    allk = np.concatenate(sets) if sets else np.zeros(0, np.uint64)
    allk = np.ascontiguousarray(allk, dtype=np.uint64)
    pr, pc, pv = (C.POINTER(C.c_uint32)() for _ in range(3))
    n = L.kmo_common(allk.ctypes.data, off.ctypes.data, len(sets), C.byref(pr), C.byref(pc), C.byref(pv))
    rows = np.ctypeslib.as_array(pr, shape=(max(n, 1),))[:n].copy()
    cols = np.ctypeslib.as_array(pc, shape=(max(n, 1),))[:n].copy()
    vals = np.ctypeslib.as_array(pv, shape=(max(n, 1),))[:n].copy()
    for p in (pr, pc, pv):
        L.kmo_free(p)
    return rows, cols, vals


def fixed6(v: float) -> str:
    """K/conversion.h:167-219 Double2PChar(val, 6)."""
    neg = v < 0
    if neg:
        v = -v
    x = int(v * 1e6 + 0.5)
    s = "%d.%06d" % (x // 1_000_000, x % 1_000_000)
    return ("-" if neg else "") + s


def fraction_str(f: float) -> str:
    return "%g" % f            # ostream << double (K/console_distance.cpp:40)


def prefilter_pairs(sets, k: int, min_kmers: int, min_ident: float, max_seqs: int = 0):
    """(row, col, common, ani_shorter) for pairs passing K/sparse_filters.h:49-61 with vclust's two -min filters.
    max_seqs > 0: then `-sample-rows ani-shorter:N` (K/sampler.h:45-121, K/array.h:450-543): each passing pair is
    offered to both of its rows, a row keeps its N best items (the heap evicts the lowest score, ties the largest
    item id), and the list holds entries on both sides of the diagonal, sorted by (row, item)."""
    L = lib()
    rows, cols, vals = common_matrix(sets)
    tot = [s.size for s in sets]
    out = []
    for r, c, v in zip(rows.tolist(), cols.tolist(), vals.tolist()):
        if v < min_kmers:
            continue
        a = L.kmo_ani_shorter(v, tot[r], tot[c], k)
        if a >= min_ident:
            out.append((r, c, v, a))
    if max_seqs > 0:
        rows = {}
        for r, c, v, a in out:
            rows.setdefault(r, []).append((c, v, a))
            rows.setdefault(c, []).append((r, v, a))
        out = []
        for r in sorted(rows):
            best = sorted(rows[r], key=lambda e: (-e[2], e[0]))[:max_seqs]
            out += [(r, c, v, a) for c, v, a in sorted(best)]
    return out


def filter_text(names, sets, k: int, fraction: float, min_kmers: int, min_ident: float, max_seqs: int = 0) -> str:
    """The `kmer-db distance ani-shorter -sparse` text (K/console_distance.cpp:37-42,183-204)."""
    pairs = prefilter_pairs(sets, k, min_kmers, min_ident, max_seqs)
    by_row = {}
    for r, c, v, a in pairs:
        by_row.setdefault(r, []).append((c, a))
    lines = ["kmer-length: %d fraction: %s ,%s," % (k, fraction_str(fraction), ",".join(names))]
    for r, nm in enumerate(names):
        ent = "".join("%d:%s," % (c + 1, fixed6(a)) for c, a in sorted(by_row.get(r, [])))
        lines.append(nm + "," + ent)
    return "\n".join(lines) + "\n"


def prefilter_text_from_fasta(paths, multifasta: bool, k=25, fraction=1.0, min_kmers=20, min_ident=0.7, max_seqs=0) -> str:
    names, samples = [], []
    if multifasta:
        for p in paths:
            for nm, s in read_records_kmerdb(p):
                names.append(nm)
                samples.append([s])
    else:
        for p in paths:
            names.append(Path(p).name)
            samples.append([s for _, s in read_records_kmerdb(p)])
    return filter_text(names, kmer_sets(samples, k, fraction), k, fraction, min_kmers, min_ident, max_seqs)


# ----------------------------------------------------------------------------------------------------------------
# align (lz-ani) restatement
# ----------------------------------------------------------------------------------------------------------------
def read_filter(path, thr: float, names):
    """L/filter.cpp:20-298: symmetric adjacency (input ids) from a kmer-db distance file."""
    with open(path, "rb") as fh:
        lines = fh.read().decode().split("\n")
    hdr = lines[0].rstrip("\r").split(",")
    hdr = [h for h in hdr][1:]
    while hdr and hdr[-1] == "":
        hdr.pop()
    if hdr != list(names):
        raise ValueError("Input sequences and filter sequences are different!")
    adj = [[] for _ in names]
    row = 0
    for ln in lines[1:]:
        ln = ln.rstrip("\r")
        if len(ln) <= 2:
            continue
        parts = ln.split(",")
        for p in parts[1:]:
            e = p.split(":")
            if len(e) == 2 and float(e[1]) >= thr:
                j = int(e[0]) - 1
                adj[row].append(j)
        row += 1
    first = [list(a) for a in adj]
    for i, a in enumerate(first):
        for j in a:
            adj[j].append(i)
    return adj


def run_pairs(codes, pair_ref, pair_qry, params: LzParams | None = None, return_oob: bool = False):
    """stats[k] = (sym_in_matches, sym_in_literals, no_components) for query pair_qry[k] parsed against pair_ref[k].
    return_oob: also return a bool array, True where the parse compared against positions outside the reference text --
    the reference reads uninitialised memory there (only possible with mqd > mrd), so its result is undefined."""
    L = lib()
    params = params or LzParams.default()
    off = np.zeros(len(codes) + 1, dtype=np.int64)
    off[1:] = np.cumsum([c.size for c in codes])
    flat = np.ascontiguousarray(np.concatenate(codes) if codes else np.zeros(0, np.uint8), dtype=np.uint8)
    pr = np.ascontiguousarray(pair_ref, dtype=np.int32)
    pq = np.ascontiguousarray(pair_qry, dtype=np.int32)
    order = np.argsort(pr, kind="stable")
    st = np.zeros((pr.size, 3), dtype=np.int32)
    pr_s, pq_s = np.ascontiguousarray(pr[order]), np.ascontiguousarray(pq[order])
    st_s = np.zeros((pr.size, 3), dtype=np.int32)
    oob_s = np.zeros(pr.size, dtype=np.uint8)
    L.lzo_run_pairs_oob.restype = None
    L.lzo_run_pairs_oob.argtypes = [C.POINTER(LzParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                    C.c_void_p]
    L.lzo_run_pairs_oob(C.byref(params), flat.ctypes.data, off.ctypes.data, pr_s.ctypes.data, pq_s.ctypes.data, pr.size,
                        st_s.ctypes.data, oob_s.ctypes.data)
    st[order] = st_s
    if return_oob:
        oob = np.zeros(pr.size, dtype=bool)
        oob[order] = oob_s.astype(bool)
        return st, oob
    return st


class LzRegion(C.Structure):                     # lzo_region in lz_oracle.c (region_t, lz-ani defs.h:67-142)
    _fields_ = [(k, C.c_int) for k in ("ref_start", "ref_end", "seq_start", "seq_end", "num_matches", "num_mismatches")]


def run_pairs_regions(codes, pair_ref, pair_qry, params: LzParams | None = None):
    """Like run_pairs, plus the regions of every pair (calc_regions order): list of (n_i, 6) int arrays
    [ref_start, ref_end, seq_start, seq_end, matches, mismatches]."""
    L = lib()
    params = params or LzParams.default()
    L.lzo_create.restype = C.c_void_p
    L.lzo_create.argtypes = [C.POINTER(LzParams)]
    L.lzo_destroy.argtypes = [C.c_void_p]
    L.lzo_set_reference.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.lzo_parse_query.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.lzo_last_regions.argtypes = [C.c_void_p, C.POINTER(LzRegion), C.c_int]
    L.lzo_last_regions.restype = C.c_int
    ctx = L.lzo_create(C.byref(params))
    stats = np.zeros((len(pair_ref), 3), dtype=np.int32)
    regions = []
    cap = 1 << 16
    buf = (LzRegion * cap)()
    st = (C.c_int * 3)()
    cur = None
    try:
        for k in np.argsort(np.asarray(pair_ref), kind="stable"):
            r, q = int(pair_ref[k]), int(pair_qry[k])
            if r != cur:
                ref = np.ascontiguousarray(codes[r], dtype=np.uint8)
                L.lzo_set_reference(ctx, ref.ctypes.data, ref.size)
                cur = r
            qc = np.ascontiguousarray(codes[q], dtype=np.uint8)
            L.lzo_parse_query(ctx, qc.ctypes.data, qc.size, st)
            stats[k] = list(st)
            n = L.lzo_last_regions(ctx, buf, cap)
            regions.append((int(k), np.array([[b.ref_start, b.ref_end, b.seq_start, b.seq_end, b.num_matches, b.num_mismatches]
                                              for b in buf[:n]], dtype=np.int64).reshape(n, 6)))
    finally:
        L.lzo_destroy(ctx)
    regions.sort(key=lambda t: t[0])
    return stats, [r for _, r in regions]


def aln_lines(names, codes, pair_ref, pair_qry, params: LzParams | None = None, out_filters=None):
    """lz-ani --out-alignment rows (without the header) for directed pairs in INPUT ids, as a list of strings:
    CLZMatcher::store_alignment, L/lz_matcher.cpp:102-169.  The reference writes them in thread-completion order, so
    callers compare sorted lists."""
    params = params or LzParams.default()
    flt = {k: float(v) for k, v in (out_filters or {}).items() if v}
    _, regs = run_pairs_regions(codes, pair_ref, pair_qry, params)
    out = []
    for k, rg in enumerate(regs):
        r, q = int(pair_ref[k]), int(pair_qry[k])
        len1, len2 = int(codes[r].size), int(codes[q].size)
        rc_corr = 2 * len1 + 2 * params.mrd + 1
        if flt:
            m, l = int(rg[:, 4].sum()), int(rg[:, 5].sum())
            if m / len2 < flt.get("gani", 0) or (m / (m + l) if m + l else 0.0) < flt.get("ani", 0) or (m + l) / len2 < flt.get("qcov", 0):
                continue
        for rs, re_, ss, se, nm, nmm in rg.tolist():
            ln = se - ss
            if rs < len1:
                a, b = 1 + rs, re_
            else:
                a, b = rc_corr - (1 + rs), rc_corr - re_
            out.append("\t".join([names[q], names[r], real_str(100.0 * nm / ln, 6), str(ln), str(1 + ss), str(se), str(a), str(b),
                                   str(nm), str(nmm)]))
    return out


def real_str(val: float, prec: int) -> str:
    """refresh::real_to_pchar (lz-ani/libs/refresh/conversions/lib/numeric_conversions.h:229-299,342-390):
    shortest round-trip decimal, then round-half-up to `prec` significant digits, no zero padding."""
    if val == 0:
        return "0"
    sign, digits, exponent = Decimal(repr(float(val))).as_tuple()
    sig = int("".join(map(str, digits)))
    while sig % 10 == 0:
        sig //= 10
        exponent += 1
    n_dig = len(str(sig))
    if n_dig > prec:
        p10 = 10 ** (n_dig - prec)
        sig = (sig + p10 // 2) // p10
        exponent += n_dig - prec
        n_dig = prec
        if sig >= 10 ** prec:
            sig //= 10
            exponent += 1
    s = str(sig)
    pre = "-" if sign else ""
    if exponent == 0:
        return pre + s
    if exponent > 0 or -exponent >= n_dig + 4:
        if n_dig == 1:
            body = s
        else:
            body = s[0] + "." + s[1:]
            exponent += n_dig - 1
        e = abs(exponent)
        return pre + body + "e" + ("-" if exponent < 0 else "+") + ("%02d" % e)
    if -exponent < n_dig:
        return pre + s[:n_dig + exponent] + "." + s[n_dig + exponent:]
    return pre + "0." + "0" * (-exponent - n_dig) + s


OUTFMT = {
    "lite": "qidx,ridx,tani,gani,ani,qcov,rcov,num_alns,len_ratio".split(","),
    "standard": "qidx,ridx,query,reference,tani,gani,ani,qcov,rcov,num_alns,len_ratio".split(","),
    "complete": "qidx,ridx,query,reference,tani,gani,ani,qcov,rcov,num_alns,len_ratio,qlen,rlen,nt_match,nt_mismatch".split(","),
}


def align_text(names, codes, adj=None, params: LzParams | None = None, columns=None, out_filters=None):
    """Full lz-ani all2all restatement: returns (ani_tsv_text, ids_tsv_text, stats dict keyed by (ref, qry) in
    re-ordered ids).  adj = symmetric adjacency in input ids (None = all-vs-all).  L/lz_matcher.cpp:172-579."""
    columns = columns or OUTFMT["standard"]
    out_filters = out_filters or {}
    n = len(names)
    lens = [int(c.size) for c in codes]
    order = lz_order(names, lens)
    rank = {g: i for i, g in enumerate(order)}
    rn = [names[g] for g in order]
    rc = [codes[g] for g in order]
    rl = [lens[g] for g in order]
    pr, pq = [], []
    for r in range(n):
        if adj is None:
            qs = [q for q in range(n) if q != r]
        else:
            qs = [rank[x] for x in adj[order[r]]]
        for q in qs:
            pr.append(r)
            pq.append(q)
    st = run_pairs(rc, pr, pq, params)
    res = [dict() for _ in range(n)]
    rows_sorted = [[] for _ in range(n)]
    for k, (r, q) in enumerate(zip(pr, pq)):
        rows_sorted[r].append((q, k))
    for r in range(n):
        rows_sorted[r].sort()
    for r in range(n):
        for q, k in rows_sorted[r]:
            res[r].setdefault(q, tuple(int(v) for v in st[k]))
    ids_txt = "id\tseq_len\tno_parts\n" + "".join("%s\t%d\t1\n" % (rn[i], rl[i]) for i in range(n))
    out = ["\t".join(columns) + "\n"]
    flt = {k: float(v) for k, v in out_filters.items() if v}
    for a in range(n):
        for b, _k in rows_sorted[a]:
            if a >= b:
                continue
            ids = (a, b)
            ln = (rl[b], rl[a])
            m = (res[a][b][0], res[b][a][0])
            l = (res[a][b][1], res[b][a][1])
            nr = (res[a][b][2], res[b][a][2])
            tani = (m[0] + m[1]) / (ln[0] + ln[1])
            gani = (m[0] / ln[0], m[1] / ln[1])
            ani = tuple(m[i] / (m[i] + l[i]) if m[i] + l[i] else 0.0 for i in range(2))
            cov = ((m[0] + l[0]) / ln[0], (m[1] + l[1]) / ln[1])
            for i in range(2):
                j = 1 - i
                if flt:
                    if gani[i] < flt.get("gani", 0) or ani[i] < flt.get("ani", 0) or tani < flt.get("tani", 0) \
                            or cov[i] < flt.get("qcov", 0) or cov[j] < flt.get("rcov", 0):
                        continue
                f = []
                for col in columns:
                    if col == "ridx": f.append(str(ids[i]))
                    elif col == "qidx": f.append(str(ids[j]))
                    elif col == "reference": f.append(rn[ids[i]])
                    elif col == "query": f.append(rn[ids[j]])
                    elif col == "qcov": f.append(real_str(cov[i], 6))
                    elif col == "rcov": f.append(real_str(cov[j], 6))
                    elif col == "gani": f.append(real_str(gani[i], 6))
                    elif col == "ani": f.append(real_str(ani[i], 6))
                    elif col == "tani": f.append(real_str(tani, 6))
                    elif col == "rlen": f.append(str(ln[j]))
                    elif col == "qlen": f.append(str(ln[i]))
                    elif col == "num_alns": f.append(str(nr[i]))
                    elif col == "nt_match": f.append(str(m[i]))
                    elif col == "nt_mismatch": f.append(str(l[i]))
                    elif col == "len_ratio":
                        if ln[0] and ln[1]:
                            f.append(real_str(min(ln) / max(ln), 4))
                        else:
                            f.append("0")
                out.append("\t".join(f) + "\n")
    stats = {(r, q): res[r][q] for r in range(n) for q in res[r]}
    return "".join(out), ids_txt, stats


# ----------------------------------------------------------------------------------------------------------------
# the unmodified reference binaries (oracle/_ref), for validating the restatement and as CPU baseline
# ----------------------------------------------------------------------------------------------------------------
def ref_prefilter(fasta_paths, out_path, workdir, multifasta=True, k=25, fraction=1.0, min_kmers=20, min_ident=0.7,
                  threads=None, max_seqs=0, timings=None):
    """Run `kmer-db build | all2all-sp | distance` exactly as vclust.py:915-1055 builds the commands."""
    import time
    threads = threads or os.cpu_count()
    wd = Path(workdir)
    wd.mkdir(parents=True, exist_ok=True)
    kdb = str(REF_DIR / "kmer-db")
    (wd / "whole.txt").write_text("".join("%s\n" % p for p in fasta_paths))
    cmds = [
        [kdb, "build"] + (["-multisample-fasta"] if multifasta else []) +
        ["-k", str(k), "-f", str(fraction), "-t", str(threads), str(wd / "whole.txt"), str(wd / "whole.kdb")],
        [kdb, "all2all-sp", "-sparse", "-min", "num-kmers:%d" % min_kmers] +
        (["-sample-rows", "ani-shorter:%d" % max_seqs] if max_seqs > 0 else []) +
        ["-min", "ani-shorter:%s" % min_ident, "-t", str(threads), str(wd / "whole.kdb"), str(wd / "all2all.txt")],
        [kdb, "distance", "ani-shorter", "-sparse", "-min", str(min_ident), "-t", str(threads),
         str(wd / "all2all.txt"), str(out_path)],
    ]
    for name, cmd in zip(("build", "all2all", "distance"), cmds):
        t0 = time.perf_counter()
        subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        if timings is not None:
            timings[name] = time.perf_counter() - t0
    return out_path


def ref_align(fasta_paths, out_path, workdir, multifasta=True, filter_path=None, filter_thr=0.0, threads=None,
              columns=None, params: LzParams | None = None, out_aln=None, timings=None):
    """Run `lz-ani all2all` exactly as vclust.py:1058-1181 builds the command."""
    import re
    import time
    threads = threads or os.cpu_count()
    params = params or LzParams.default()
    columns = columns or OUTFMT["standard"]
    wd = Path(workdir)
    wd.mkdir(parents=True, exist_ok=True)
    (wd / "ids.txt").write_text("".join("%s\n" % p for p in fasta_paths))
    cmd = [str(REF_DIR / "lz-ani"), "all2all", "--in-txt", str(wd / "ids.txt"), "-o", str(out_path), "-t", str(threads)]
    for nm in ("mal", "msl", "mrd", "mqd", "reg", "aw", "am", "ar"):
        cmd += ["--" + nm, str(getattr(params, nm))]
    cmd += ["--multisample-fasta", "true" if multifasta else "false", "--out-type", "tsv", "--out-format", ",".join(columns)]
    if filter_path:
        cmd += ["--flt-kmerdb", str(filter_path), str(filter_thr)]
    if out_aln:
        cmd += ["--out-alignment", str(out_aln)]
    cmd += ["--verbose", "2"]
    t0 = time.perf_counter()
    p = subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
    if timings is not None:
        timings["wall"] = time.perf_counter() - t0
        m = re.search(r"LZ matching : ([0-9.eE+-]+)s", p.stderr)
        if m:
            timings["lz_matching"] = float(m.group(1))
    return out_path
