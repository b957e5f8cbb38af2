"""The patch that puts libvclust_b200.so under the reference's own entry point, `vclust.py`.

``apply(source)`` takes the text of the reference's vclust.py and returns it with (1) a ctypes binding of the C ABI
(stdlib only, like vclust.py itself) next to the BIN_* constants and (2) one line at the top of ``handle_prefilter``
(vclust.py:1380) and of ``handle_align`` (vclust.py:1474) that hands the work to the library when it and a GPU are
present, and otherwise falls through to the unchanged subprocess path (kmer-db / lz-ani).  INTEGRATION.md shows the same
code; ``oracle/make_dropin.py`` applies it to a copy of the reference (build output under oracle/_ref/, never committed)
and tests/test_dropin.py runs the reference's own test.py against the result.

    python -m vclust_b200.integration /path/to/vclust.py > vclust_patched.py
"""
from __future__ import annotations

BINDING = r'''
# ---- libvclust_b200: B200-native `prefilter` and `align` (C ABI: include/vclust_b200.h) --------------------------------
import ctypes
LIB_B200 = pathlib.Path(os.environ.get('VCLUST_B200_LIB', str(BIN_DIR / 'libvclust_b200.so')))


class _B200PrefilterParams(ctypes.Structure):
    _fields_ = [('k', ctypes.c_int32), ('min_kmers', ctypes.c_int32), ('min_ident', ctypes.c_double),
                ('kmers_fraction', ctypes.c_double), ('max_seqs', ctypes.c_int32), ('batch_size', ctypes.c_int32)]


class _B200AlignParams(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ('mal', 'msl', 'mrd', 'mqd', 'reg', 'aw', 'am', 'ar')]


def _b200():
    """The loaded library, or None when it (or a GPU) is absent -> the subprocess path is used."""
    if os.environ.get('VCLUST_B200_DISABLE') or not LIB_B200.exists():
        return None
    try:
        lib = ctypes.CDLL(str(LIB_B200))
    except OSError:
        return None
    lib.vb_last_error.restype = ctypes.c_char_p
    return lib if lib.vb_device_count() > 0 else None


def _b200_check(lib, rc, what, logger):
    if rc != 0:                                   # same contract as run(): log + exit 1
        logger.error(f'{what} failed with message: {lib.vb_last_error().decode(errors="replace")}')
        sys.exit(1)


def _b200_paths(paths):
    return (ctypes.c_char_p * len(paths))(*[str(p).encode() for p in paths])


def _b200_prefilter(args, logger) -> bool:
    lib = _b200()
    if lib is None:
        return False
    what = f'libvclust_b200 prefilter -k {args.k} --min-kmers {args.min_kmers} --min-ident {args.min_ident}'
    logger.info(f'Running: {what}')
    ctx, g, pairs = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
    _b200_check(lib, lib.vb_ctx_create(0, ctypes.byref(ctx)), what, logger)
    _b200_check(lib, lib.vb_genomes_load(_b200_paths(args.fasta_paths), len(args.fasta_paths), int(args.is_multifasta), 0, 0,
                                         ctypes.byref(g)), what, logger)
    p = _B200PrefilterParams(args.k, args.min_kmers, args.min_ident, args.kmers_fraction, args.max_seqs or 0,
                             args.batch_size or 0)
    _b200_check(lib, lib.vb_prefilter(ctx, g, ctypes.byref(p), ctypes.byref(pairs)), what, logger)
    _b200_check(lib, lib.vb_write_filter(g, pairs, str(args.output_path).encode()), what, logger)
    lib.vb_pairs_free(pairs); lib.vb_genomes_free(g); lib.vb_ctx_destroy(ctx)
    logger.info('Completed')
    return True


def _b200_align(args, logger) -> bool:
    lib = _b200()
    if lib is None:
        return False
    what = f'libvclust_b200 align --mal {args.mal} --msl {args.msl} --mrd {args.mrd} --mqd {args.mqd}'
    logger.info(f'Running: {what}')
    lib.vb_read_filter.argtypes = [ctypes.c_char_p, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p]
    ctx, g, flt, res, reg = (ctypes.c_void_p() for _ in range(5))
    _b200_check(lib, lib.vb_ctx_create(0, ctypes.byref(ctx)), what, logger)
    _b200_check(lib, lib.vb_genomes_load(_b200_paths(args.fasta_paths), len(args.fasta_paths), int(args.is_multifasta), 1,
                                         int(args.mrd), ctypes.byref(g)), what, logger)
    if args.filter_path:
        _b200_check(lib, lib.vb_read_filter(str(args.filter_path).encode(), float(args.filter_threshold), g,
                                            ctypes.byref(flt)), what, logger)
    p = _B200AlignParams(args.mal, args.msl, args.mrd, args.mqd, args.reg, args.aw, args.am, args.ar)
    out_filters = (ctypes.c_double * 5)(*[float(v or 0) for v in (args.tani, args.gani, args.ani, args.qcov, args.rcov)])
    if args.aln_path:
        _b200_check(lib, lib.vb_align_regions(ctx, g, flt, ctypes.byref(p), ctypes.byref(res), ctypes.byref(reg)), what, logger)
        _b200_check(lib, lib.vb_write_aln(g, reg, str(args.aln_path).encode(), out_filters), what, logger)
        lib.vb_regions_free(reg)
    else:
        _b200_check(lib, lib.vb_align(ctx, g, flt, ctypes.byref(p), ctypes.byref(res)), what, logger)
    cols = ALIGN_OUTFMT[args.outfmt]
    c_cols = (ctypes.c_char_p * len(cols))(*[c.encode() for c in cols])
    out = str(args.output_path)
    dot = out.rfind('.')                          # lz-ani names the ids file <stem>.ids<ext>
    ids = out + '.ids' if dot < 0 or '/' in out[dot:] else out[:dot] + '.ids' + out[dot:]
    _b200_check(lib, lib.vb_write_ani(g, res, out.encode(), ids.encode(), c_cols, len(cols), out_filters), what, logger)
    lib.vb_align_out_free(res)
    if flt:
        lib.vb_pairs_free(flt)
    lib.vb_genomes_free(g); lib.vb_ctx_destroy(ctx)
    logger.info('Completed')
    return True

'''

ANCHOR_BINDING = "# LZ-ANI output columns\n"
ANCHOR_PREFILTER = ("    validate_binary(BIN_KMERDB)\n    args = validate_args_prefilter(args, parser)\n"
                    "    args = validate_args_fasta_input(args, parser)\n")
ANCHOR_ALIGN = "    validate_binary(BIN_LZANI)\n    args = validate_args_fasta_input(args, parser)\n"


def apply(source: str) -> str:
    """vclust.py text -> vclust.py text with the GPU branch.  Raises if an anchor is missing (a different vclust.py)."""
    if "_b200_prefilter" in source:
        raise ValueError("this vclust.py already carries the libvclust_b200 patch")
    for anchor in (ANCHOR_BINDING, ANCHOR_PREFILTER, ANCHOR_ALIGN):
        if source.count(anchor) != 1:
            raise ValueError("vclust.py does not look like the version this patch was written for: %r" % anchor.strip().splitlines()[0])
    out = source.replace(ANCHOR_BINDING, BINDING + "\n" + ANCHOR_BINDING)
    out = out.replace(ANCHOR_PREFILTER, ANCHOR_PREFILTER + "    if _b200_prefilter(args, logger):\n        return\n")
    out = out.replace(ANCHOR_ALIGN, ANCHOR_ALIGN + "    if _b200_align(args, logger):\n        return\n")
    return out


if __name__ == "__main__":
    import sys
    sys.stdout.write(apply(open(sys.argv[1]).read()))
