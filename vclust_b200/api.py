"""Host-side mirror of the reference's interface for the hot path.

The reference's "operator API" for this path is the pair of command builders in vclust.py -- ``cmd_kmerdb_build`` /
``cmd_kmerdb_all2all`` / ``cmd_kmerdb_distance`` (vclust.py:915-1055) and ``cmd_lzani`` (vclust.py:1058-1181) -- plus
``run()``.  ``prefilter()`` and ``align()`` below take the same arguments with the same meaning and produce the same
files; instead of spawning kmer-db / lz-ani they call libvclust_b200.so through ctypes.  Errors raise ``VbError``
(the CLI maps that to ``sys.exit(1)`` exactly like vclust.py:797-805).
"""
from __future__ import annotations

import ctypes as C
import time
from pathlib import Path
from typing import Iterable, Sequence

import numpy as np

from . import _lib
from ._lib import AlignOut, AlignParams, Pairs, PrefilterParams, VbError, check

# vclust.py:38-47
ALIGN_OUTFMT = {
    "lite": ["qidx", "ridx", "tani", "gani", "ani", "qcov", "rcov", "num_alns", "len_ratio"],
    "standard": ["qidx", "ridx", "query", "reference", "tani", "gani", "ani", "qcov", "rcov", "num_alns", "len_ratio"],
    "complete": ["qidx", "ridx", "query", "reference", "tani", "gani", "ani", "qcov", "rcov", "num_alns", "len_ratio",
                 "qlen", "rlen", "nt_match", "nt_mismatch"],
}

FASTA_KMERDB, FASTA_LZANI = 0, 1


def version() -> str:
    buf = C.create_string_buffer(256)
    check(_lib.load().vb_version(buf, 256))
    return buf.value.decode()


def device_count() -> int:
    return int(_lib.load().vb_device_count())


class Context:
    """One CUDA device (vb_ctx).  Raises VbError when no device is usable -- there is no CPU fallback."""

    def __init__(self, device: int = 0, stream: int | None = None):
        """stream: a cudaStream_t (e.g. ``torch.cuda.Stream().cuda_stream``) the library enqueues all its work on, so that
        the caller's collectives on the same stream are ordered with it; None = the context's own stream."""
        self._L = _lib.load()
        self._h = C.c_void_p()
        if stream:
            check(self._L.vb_ctx_create_on_stream(int(device), C.c_void_p(int(stream)), C.byref(self._h)))
        else:
            check(self._L.vb_ctx_create(int(device), C.byref(self._h)))

    def close(self):
        if self._h:
            self._L.vb_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def timing(self, key: str) -> float:
        v = C.c_double()
        check(self._L.vb_ctx_timing(self._h, key.encode(), C.byref(v)))
        return v.value

    def timings(self, prefix: str) -> dict:
        keys = {
            "prefilter": ["total_ms", "upload_pack_ms", "extract_ms", "sort_ms", "segment_ms", "exchange_ms", "emit_ms", "host_post_ms",
                          "passes", "tuples", "survivors", "grouped", "bubbles", "table_slots", "candidates", "peer_bytes"],
            "align": ["total_ms", "upload_pack_ms", "list_ms", "index_ms", "parse_ms", "gather_ms", "api_prep_ms", "host_prep_ms",
                      "host_post_ms", "batches", "pairs"],
        }[prefix]
        out = {}
        for k in keys:
            try:
                out[k] = self.timing(prefix + "." + k)
            except VbError:
                pass
        return out

    @property
    def launches(self) -> int:
        return int(self._L.vb_ctx_launches(self._h))

    def mark(self, slot: int) -> None:
        check(self._L.vb_ctx_mark(self._h, slot))

    def elapsed_ms(self, a: int, b: int) -> float:
        v = C.c_double()
        check(self._L.vb_ctx_elapsed_ms(self._h, a, b, C.byref(v)))
        return v.value

    def make_resident(self, genomes: "Genomes", rule: int, mrd: int = 40) -> None:
        check(self._L.vb_genomes_make_resident(self._h, genomes._h, rule, mrd))

    def evict(self, genomes: "Genomes | None" = None) -> None:
        check(self._L.vb_genomes_evict(self._h, genomes._h if genomes is not None else None))


class Genomes:
    """A genome set on the host (vb_genomes)."""

    def __init__(self, handle, keepalive=None):
        self._L = _lib.load()
        self._h = handle
        self._keep = keepalive

    @classmethod
    def load(cls, paths: Sequence, multisample: bool, flavor: int, sep_len: int = 40) -> "Genomes":
        L = _lib.load()
        arr = (C.c_char_p * len(paths))(*[str(p).encode() for p in paths])
        h = C.c_void_p()
        check(L.vb_genomes_load(arr, len(paths), 1 if multisample else 0, flavor, sep_len, C.byref(h)))
        return cls(h)

    @classmethod
    def from_memory(cls, names: Sequence[str], seqs: Sequence) -> "Genomes":
        """seqs: bytes objects or uint8 numpy arrays of ASCII bases."""
        L = _lib.load()
        n = len(names)
        arrs = [np.ascontiguousarray(np.frombuffer(s, dtype=np.uint8) if isinstance(s, (bytes, bytearray)) else s,
                                     dtype=np.uint8) for s in seqs]
        c_names = (C.c_char_p * n)(*[x.encode() for x in names])
        c_seqs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
        c_lens = (C.c_uint64 * n)(*[a.size for a in arrs])
        h = C.c_void_p()
        check(L.vb_genomes_from_memory(c_names, c_seqs, c_lens, n, C.byref(h)))
        return cls(h)

    @classmethod
    def skeleton(cls, names: Sequence[str], lengths: Sequence[int]) -> "Genomes":
        """Names and lengths only (what a rank of a multi-GPU run knows about the genomes it does not hold)."""
        L = _lib.load()
        n = len(names)
        c_names = (C.c_char_p * n)(*[x.encode() for x in names])
        c_lens = (C.c_uint64 * n)(*[int(x) for x in lengths])
        h = C.c_void_p()
        check(L.vb_genomes_skeleton(c_names, c_lens, n, C.byref(h)))
        return cls(h)

    def __len__(self):
        return int(self._L.vb_genomes_count(self._h))

    def name(self, i: int) -> str:
        return self._L.vb_genomes_name(self._h, i).decode()

    def names(self):
        return [self.name(i) for i in range(len(self))]

    def length(self, i: int) -> int:
        return int(self._L.vb_genomes_length(self._h, i))

    def sequence(self, i: int) -> bytes:
        n = self.length(i)
        p = self._L.vb_genomes_sequence(self._h, i)
        return C.string_at(p, n) if n else b""

    @property
    def total_bases(self) -> int:
        return int(self._L.vb_genomes_total_bases(self._h))

    def close(self):
        if self._h:
            self._L.vb_genomes_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PairList:
    """vb_pairs: sparse lower-triangular prefilter result (or a filter file read back)."""

    def __init__(self, ptr):
        self._L = _lib.load()
        self._p = ptr

    def _arr(self, field, n, dtype):
        if n == 0:
            return np.zeros(0, dtype=dtype)
        return np.ctypeslib.as_array(getattr(self._p.contents, field), shape=(n,)).copy()

    @property
    def n_pairs(self):
        return int(self._p.contents.n_pairs)

    @property
    def rows(self):
        return self._arr("row", self.n_pairs, np.uint32)

    @property
    def cols(self):
        return self._arr("col", self.n_pairs, np.uint32)

    @property
    def common(self):
        return self._arr("common", self.n_pairs, np.uint32)

    @property
    def ani(self):
        return self._arr("ani", self.n_pairs, np.float64)

    @property
    def total_kmers(self):
        return self._arr("total_kmers", int(self._p.contents.n_genomes), np.uint32)

    def close(self):
        if self._p:
            self._L.vb_pairs_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class AlignResult:
    def __init__(self, ptr):
        self._L = _lib.load()
        self._p = ptr

    def _arr(self, field, n, dtype):
        if n == 0:
            return np.zeros(0, dtype=dtype)
        return np.ctypeslib.as_array(getattr(self._p.contents, field), shape=(n,)).copy()

    @property
    def n(self):
        return int(self._p.contents.n)

    @property
    def ref(self):
        return self._arr("ref", self.n, np.uint32)

    @property
    def qry(self):
        return self._arr("qry", self.n, np.uint32)

    @property
    def stats(self):
        return np.stack([self._arr("sym_in_matches", self.n, np.int32), self._arr("sym_in_literals", self.n, np.int32),
                         self._arr("no_components", self.n, np.int32)], axis=1) if self.n else np.zeros((0, 3), np.int32)

    @property
    def order(self):
        return self._arr("order", int(self._p.contents.n_genomes), np.uint32)

    def close(self):
        if self._p:
            self._L.vb_align_out_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ----------------------------------------------------------------------------------------------------------------
# operator-level calls
# ----------------------------------------------------------------------------------------------------------------
def prefilter_genomes(ctx: Context, genomes: Genomes, k: int = 25, min_kmers: int = 20, min_ident: float = 0.7,
                      kmers_fraction: float = 1.0, max_seqs: int = 0, batch_size: int = 0) -> PairList:
    p = PrefilterParams(k, min_kmers, min_ident, kmers_fraction, max_seqs, batch_size)
    out = C.POINTER(Pairs)()
    check(ctx._L.vb_prefilter(ctx._h, genomes._h, C.byref(p), C.byref(out)))
    return PairList(out)


def prefilter_partial(ctx: Context, genomes: Genomes, shard_index: int, shard_count: int, k: int = 25,
                      kmers_fraction: float = 1.0) -> PairList:
    """Partial counts of one k-mer hash shard (multi-GPU building block): no thresholds applied."""
    p = PrefilterParams(k, 0, 0.0, kmers_fraction, 0, 0)
    out = C.POINTER(Pairs)()
    check(ctx._L.vb_prefilter_partial(ctx._h, genomes._h, C.byref(p), shard_index, shard_count, C.byref(out)))
    return PairList(out)


def merge_pairs(rows, cols, common, total_kmers, k: int = 25, min_kmers: int = 20, min_ident: float = 0.7,
                kmers_fraction: float = 1.0, max_seqs: int = 0) -> PairList:
    """Sum partial (row, col, common) triples, apply the two -min filters exactly (and --max-seqs); host only."""
    r = np.ascontiguousarray(rows, dtype=np.uint32)
    c = np.ascontiguousarray(cols, dtype=np.uint32)
    v = np.ascontiguousarray(common, dtype=np.uint32)
    t = np.ascontiguousarray(total_kmers, dtype=np.uint32)
    p = PrefilterParams(k, min_kmers, min_ident, kmers_fraction, max_seqs, 0)
    out = C.POINTER(Pairs)()
    check(_lib.load().vb_pairs_merge(r.ctypes.data, c.ctypes.data, v.ctypes.data, r.size, t.ctypes.data, t.size,
                                     C.byref(p), C.byref(out)))
    return PairList(out)


def align_result_from_pairs(genomes: Genomes, ref, qry, stats) -> AlignResult:
    r = np.ascontiguousarray(ref, dtype=np.uint32)
    q = np.ascontiguousarray(qry, dtype=np.uint32)
    st = np.ascontiguousarray(stats, dtype=np.int32)
    out = C.POINTER(AlignOut)()
    check(_lib.load().vb_align_out_from_pairs(genomes._h, r.ctypes.data, q.ctypes.data, st.ctypes.data, r.size,
                                              C.byref(out)))
    return AlignResult(out)


def write_filter(genomes: Genomes, pairs: PairList, path) -> None:
    check(_lib.load().vb_write_filter(genomes._h, pairs._p, str(path).encode()))


def read_filter(path, thr: float, genomes: Genomes) -> PairList:
    out = C.POINTER(Pairs)()
    check(_lib.load().vb_read_filter(str(path).encode(), float(thr), genomes._h, C.byref(out)))
    return PairList(out)


def align_params(mal=11, msl=7, mrd=40, mqd=40, reg=35, aw=15, am=7, ar=3) -> AlignParams:
    return AlignParams(mal, msl, mrd, mqd, reg, aw, am, ar)


def align_genomes(ctx: Context, genomes: Genomes, pairs: PairList | None = None, params: AlignParams | None = None
                  ) -> AlignResult:
    params = params or align_params()
    out = C.POINTER(AlignOut)()
    check(ctx._L.vb_align(ctx._h, genomes._h, pairs._p if pairs is not None else None, C.byref(params), C.byref(out)))
    return AlignResult(out)


def align_pairs(ctx: Context, genomes: Genomes, ref: Iterable[int], qry: Iterable[int],
                params: AlignParams | None = None) -> np.ndarray:
    """Directed pairs in input-order ids -> (n, 3) int32 array of (sym_in_matches, sym_in_literals, no_components)."""
    params = params or align_params()
    r = np.ascontiguousarray(ref, dtype=np.uint32)
    q = np.ascontiguousarray(qry, dtype=np.uint32)
    st = np.zeros((r.size, 3), dtype=np.int32)
    check(ctx._L.vb_align_pairs(ctx._h, genomes._h, r.ctypes.data, q.ctypes.data, r.size, C.byref(params), st.ctypes.data))
    return st


class RegionList:
    """Alignment regions (vb_regions): one row per local alignment, grouped by directed pair."""

    def __init__(self, ptr):
        self._p = ptr
        self._L = _lib.load()

    @property
    def n(self) -> int:
        return int(self._p.contents.n)

    def _arr(self, name, dtype):
        return np.ctypeslib.as_array(getattr(self._p.contents, name), shape=(self.n,)).astype(dtype) if self.n else np.zeros(0, dtype)

    def table(self) -> np.ndarray:
        """(n, 8) int64: ref, qry (input-order ids), q_start, q_end, r_start, r_end, matches, mismatches."""
        cols = [self._arr("ref", np.int64), self._arr("qry", np.int64)] + \
               [self._arr(k, np.int64) for k in ("q_start", "q_end", "r_start", "r_end", "matches", "mismatches")]
        return np.stack(cols, axis=1) if self.n else np.zeros((0, 8), np.int64)

    def close(self):
        if self._p:
            self._L.vb_regions_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def align_genomes_regions(ctx: Context, genomes: Genomes, pairs: PairList | None = None, params: AlignParams | None = None):
    """vb_align_regions: the statistics of vb_align plus the alignment regions of every directed pair."""
    params = params or align_params()
    out = C.POINTER(AlignOut)()
    reg = C.POINTER(_lib.Regions)()
    check(ctx._L.vb_align_regions(ctx._h, genomes._h, pairs._p if pairs is not None else None, C.byref(params),
                                  C.byref(out), C.byref(reg)))
    return AlignResult(out), RegionList(reg)


def align_pairs_regions(ctx: Context, genomes: Genomes, ref: Iterable[int], qry: Iterable[int],
                        params: AlignParams | None = None):
    params = params or align_params()
    r = np.ascontiguousarray(ref, dtype=np.uint32)
    q = np.ascontiguousarray(qry, dtype=np.uint32)
    st = np.zeros((r.size, 3), dtype=np.int32)
    reg = C.POINTER(_lib.Regions)()
    check(ctx._L.vb_align_pairs_regions(ctx._h, genomes._h, r.ctypes.data, q.ctypes.data, r.size, C.byref(params),
                                        st.ctypes.data, C.byref(reg)))
    return st, RegionList(reg)


def write_aln(genomes: Genomes, regions: RegionList, path, out_filters: dict | None = None) -> None:
    of = out_filters or {}
    flt = (C.c_double * 5)(*[float(of.get(k, 0) or 0) for k in ("tani", "gani", "ani", "qcov", "rcov")])
    check(_lib.load().vb_write_aln(genomes._h, regions._p, str(path).encode(), flt))


def write_ani(genomes: Genomes, res: AlignResult, ani_path, ids_path=None, columns: Sequence[str] | None = None,
              out_filters: dict | None = None) -> None:
    columns = list(columns or ALIGN_OUTFMT["standard"])
    ani_path = Path(ani_path)
    if ids_path is None:                      # lz_matcher.cpp:295-302
        s = str(ani_path)
        dot = s.rfind(".")
        ids_path = s + ".ids" if dot < 0 else s[:dot] + ".ids" + s[dot:]
    cols = (C.c_char_p * len(columns))(*[c.encode() for c in columns])
    of = out_filters or {}
    flt = (C.c_double * 5)(*[float(of.get(k, 0) or 0) for k in ("tani", "gani", "ani", "qcov", "rcov")])
    check(_lib.load().vb_write_ani(genomes._h, res._p, str(ani_path).encode(), str(ids_path).encode(), cols,
                                   len(columns), flt))


def prefilter(input_paths: Sequence, output_path, is_multisample_fasta: bool, kmer_size: int = 25,
              kmers_fraction: float = 1.0, min_kmers: int = 20, min_ident: float = 0.7, max_seqs: int = 0,
              batch_size: int = 0, device: int = 0) -> dict:
    """`vclust prefilter` body: cmd_kmerdb_build + cmd_kmerdb_all2all + cmd_kmerdb_distance (vclust.py:915-1055)."""
    t = [time.perf_counter()]
    lap = lambda: t.append(time.perf_counter())
    with Context(device) as ctx:
        lap()
        g = Genomes.load(input_paths, is_multisample_fasta, FASTA_KMERDB)
        lap()
        pairs = prefilter_genomes(ctx, g, kmer_size, min_kmers, min_ident, kmers_fraction, max_seqs, batch_size)
        lap()
        write_filter(g, pairs, output_path)
        lap()
        info = ctx.timings("prefilter")
        info["pairs"] = pairs.n_pairs
        pairs.close()
        g.close()
    lap()
    info["wall_s"] = dict(zip(("context", "read_fasta", "prefilter", "write_filter", "close"), np.diff(t).round(4).tolist()))
    return info


def align(input_paths: Sequence, output_path, is_multisample_fasta: bool, out_format: Sequence[str] | None = None,
          filter_file=None, filter_threshold: float = 0.0, out_filters: dict | None = None, mal=11, msl=7, mrd=40,
          mqd=40, reg=35, aw=15, am=7, ar=3, device: int = 0, out_aln=None) -> dict:
    """`vclust align` body: cmd_lzani (vclust.py:1058-1181); out_aln = --out-aln (lz-ani --out-alignment)."""
    t = [time.perf_counter()]
    lap = lambda: t.append(time.perf_counter())
    with Context(device) as ctx:
        lap()
        g = Genomes.load(input_paths, is_multisample_fasta, FASTA_LZANI, sep_len=mrd)
        lap()
        pairs = read_filter(filter_file, filter_threshold, g) if filter_file else None
        lap()
        if out_aln:
            res, regions = align_genomes_regions(ctx, g, pairs, align_params(mal, msl, mrd, mqd, reg, aw, am, ar))
            write_aln(g, regions, out_aln, out_filters)
            regions.close()
        else:
            res = align_genomes(ctx, g, pairs, align_params(mal, msl, mrd, mqd, reg, aw, am, ar))
        lap()
        write_ani(g, res, output_path, None, out_format or ALIGN_OUTFMT["standard"], out_filters)
        lap()
        info = ctx.timings("align")
        res.close()
        if pairs is not None:
            pairs.close()
        g.close()
    lap()
    info["wall_s"] = dict(zip(("context", "read_fasta", "read_filter", "align", "write_ani", "close"), np.diff(t).round(4).tolist()))
    return info
