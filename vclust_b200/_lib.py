"""ctypes binding of libvclust_b200.so (include/vclust_b200.h).  There is no CPU fallback: if the shared library is
missing the import fails loudly, and if no CUDA device is present ``Context()`` raises."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "libvclust_b200.so"

EXPORTS = [
    "vb_version", "vb_device_count", "vb_last_error", "vb_ctx_create", "vb_ctx_destroy", "vb_ctx_timing",
    "vb_ctx_launches", "vb_ctx_mark", "vb_ctx_elapsed_ms", "vb_genomes_make_resident", "vb_genomes_evict", "vb_genomes_load", "vb_genomes_from_memory", "vb_genomes_count", "vb_genomes_name",
    "vb_genomes_length", "vb_genomes_total_bases", "vb_genomes_sequence", "vb_genomes_free", "vb_prefilter", "vb_write_filter",
    "vb_read_filter", "vb_pairs_free", "vb_prefilter_partial", "vb_pairs_merge", "vb_align_out_from_pairs", "vb_align", "vb_align_pairs", "vb_write_ani", "vb_align_out_free",
    "vb_align_regions", "vb_align_pairs_regions", "vb_write_aln", "vb_regions_free",
    "vb_ctx_create_on_stream", "vb_genomes_skeleton", "vb_shard_create", "vb_shard_prefilter", "vb_shard_align",
    "vb_shard_destroy", "vb_comm_selftest",
]


class VbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("libvclust_b200 error %d: %s" % (code, msg))
        self.code = code


class PrefilterParams(C.Structure):
    _fields_ = [("k", C.c_int32), ("min_kmers", C.c_int32), ("min_ident", C.c_double), ("kmers_fraction", C.c_double),
                ("max_seqs", C.c_int32), ("batch_size", C.c_int32)]


class Pairs(C.Structure):
    _fields_ = [("n_pairs", C.c_uint64), ("row", C.POINTER(C.c_uint32)), ("col", C.POINTER(C.c_uint32)),
                ("common", C.POINTER(C.c_uint32)), ("ani", C.POINTER(C.c_double)), ("n_genomes", C.c_uint32),
                ("total_kmers", C.POINTER(C.c_uint32)), ("k", C.c_int32), ("kmers_fraction", C.c_double)]


class AlignParams(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("mal", "msl", "mrd", "mqd", "reg", "aw", "am", "ar")]


class AlignOut(C.Structure):
    _fields_ = [("n", C.c_uint64), ("ref", C.POINTER(C.c_uint32)), ("qry", C.POINTER(C.c_uint32)),
                ("sym_in_matches", C.POINTER(C.c_int32)), ("sym_in_literals", C.POINTER(C.c_int32)),
                ("no_components", C.POINTER(C.c_int32)), ("n_genomes", C.c_uint32), ("order", C.POINTER(C.c_uint32))]


class Regions(C.Structure):
    _fields_ = [("n", C.c_uint64), ("ref", C.POINTER(C.c_uint32)), ("qry", C.POINTER(C.c_uint32))] + \
               [(k, C.POINTER(C.c_int32)) for k in ("q_start", "q_end", "r_start", "r_end", "matches", "mismatches")] + \
               [("mrd", C.c_int32)]


# vb_comm: the collectives a multi-GPU run needs, supplied by the host as C callbacks on device memory
A2A_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64), C.c_void_p, C.POINTER(C.c_uint64), C.c_uint32)
GATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64)
REDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_uint64)


class Comm(C.Structure):
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("user", C.c_void_p), ("all_to_all", A2A_FN),
                ("all_gather", GATHER_FN), ("all_reduce_sum_u32", REDUCE_FN)]


_lib = None


def load():
    """Load the shared library (built in-tree by vclust_b200/build.py); raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError("%s is missing: run `python -m vclust_b200.build` (needs nvcc); there is no CPU fallback"
                          % LIB_PATH)
    L = C.CDLL(str(LIB_PATH))
    vp, cp, i32, u32, u64, dbl = C.c_void_p, C.c_char_p, C.c_int, C.c_uint32, C.c_uint64, C.c_double
    sig = {
        "vb_version": (i32, [C.c_char_p, C.c_size_t]),
        "vb_device_count": (i32, []),
        "vb_last_error": (cp, []),
        "vb_ctx_create": (i32, [i32, C.POINTER(vp)]),
        "vb_ctx_destroy": (None, [vp]),
        "vb_ctx_timing": (i32, [vp, cp, C.POINTER(dbl)]),
        "vb_ctx_launches": (u64, [vp]),
        "vb_ctx_mark": (i32, [vp, i32]),
        "vb_ctx_elapsed_ms": (i32, [vp, i32, i32, C.POINTER(dbl)]),
        "vb_genomes_make_resident": (i32, [vp, vp, i32, i32]),
        "vb_genomes_evict": (i32, [vp, vp]),
        "vb_genomes_load": (i32, [C.POINTER(cp), i32, i32, i32, i32, C.POINTER(vp)]),
        "vb_genomes_from_memory": (i32, [C.POINTER(cp), C.POINTER(vp), C.POINTER(u64), u32, C.POINTER(vp)]),
        "vb_genomes_count": (u32, [vp]),
        "vb_genomes_name": (cp, [vp, u32]),
        "vb_genomes_length": (u64, [vp, u32]),
        "vb_genomes_total_bases": (u64, [vp]),
        "vb_genomes_sequence": (vp, [vp, u32]),
        "vb_genomes_free": (None, [vp]),
        "vb_prefilter": (i32, [vp, vp, C.POINTER(PrefilterParams), C.POINTER(C.POINTER(Pairs))]),
        "vb_prefilter_partial": (i32, [vp, vp, C.POINTER(PrefilterParams), u32, u32, C.POINTER(C.POINTER(Pairs))]),
        "vb_pairs_merge": (i32, [vp, vp, vp, u64, vp, u32, C.POINTER(PrefilterParams), C.POINTER(C.POINTER(Pairs))]),
        "vb_align_out_from_pairs": (i32, [vp, vp, vp, vp, u64, C.POINTER(C.POINTER(AlignOut))]),
        "vb_write_filter": (i32, [vp, C.POINTER(Pairs), cp]),
        "vb_read_filter": (i32, [cp, dbl, vp, C.POINTER(C.POINTER(Pairs))]),
        "vb_pairs_free": (None, [C.POINTER(Pairs)]),
        "vb_align": (i32, [vp, vp, C.POINTER(Pairs), C.POINTER(AlignParams), C.POINTER(C.POINTER(AlignOut))]),
        "vb_align_pairs": (i32, [vp, vp, vp, vp, u64, C.POINTER(AlignParams), vp]),
        "vb_write_ani": (i32, [vp, C.POINTER(AlignOut), cp, cp, C.POINTER(cp), i32, C.POINTER(dbl)]),
        "vb_align_out_free": (None, [C.POINTER(AlignOut)]),
        "vb_align_regions": (i32, [vp, vp, C.POINTER(Pairs), C.POINTER(AlignParams), C.POINTER(C.POINTER(AlignOut)),
                                   C.POINTER(C.POINTER(Regions))]),
        "vb_align_pairs_regions": (i32, [vp, vp, vp, vp, u64, C.POINTER(AlignParams), vp, C.POINTER(C.POINTER(Regions))]),
        "vb_write_aln": (i32, [vp, C.POINTER(Regions), cp, C.POINTER(dbl)]),
        "vb_regions_free": (None, [C.POINTER(Regions)]),
        "vb_ctx_create_on_stream": (i32, [i32, vp, C.POINTER(vp)]),
        "vb_genomes_skeleton": (i32, [C.POINTER(cp), C.POINTER(u64), u32, C.POINTER(vp)]),
        "vb_shard_create": (i32, [vp, C.POINTER(Comm), vp, vp, u32, i32, C.POINTER(vp)]),
        "vb_shard_prefilter": (i32, [vp, C.POINTER(PrefilterParams), C.POINTER(C.POINTER(Pairs))]),
        "vb_shard_align": (i32, [vp, C.POINTER(AlignParams), C.POINTER(C.POINTER(AlignOut))]),
        "vb_shard_destroy": (None, [vp]),
        "vb_comm_selftest": (i32, [vp, C.POINTER(Comm)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def check(rc: int):
    if rc != 0:
        raise VbError(rc, load().vb_last_error().decode(errors="replace"))
