"""`vclust prefilter` / `vclust align` with the GPU path behind them.

Same flag surface, defaults, validation messages and exit behaviour as the two sub-commands of the reference's
vclust.py (argument definitions vclust.py:172-421, handlers :1380-1521); `cluster`, `deduplicate` and `info` are out of
scope here and stay with the reference's vclust.py (INTEGRATION.md shows the two-line patch that makes vclust.py call
this module from handle_prefilter / handle_align).

    python -m vclust_b200.cli prefilter -i genomes.fna -o fltr.txt
    python -m vclust_b200.cli align -i genomes.fna -o ani.tsv --filter fltr.txt
"""
from __future__ import annotations

import argparse
import logging
import multiprocessing
import pathlib
import sys

from . import api

DEFAULT_THREAD_COUNT = min(multiprocessing.cpu_count(), 64)


def _input_path(value):
    path = pathlib.Path(value)
    if not path.exists():
        raise argparse.ArgumentTypeError(f"input does not exist: {value}")
    return path


def _ranged_float(value):
    f = float(value)
    if f < 0 or f > 1:
        raise argparse.ArgumentTypeError(f"{value} must be between 0 and 1")
    return f


def get_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(prog="vclust-b200", description="B200-native vclust prefilter / align")
    sub = parser.add_subparsers(dest="command")

    def common(p):
        p.add_argument("-t", "--threads", metavar="<int>", dest="num_threads", type=int, default=DEFAULT_THREAD_COUNT,
                       help="Number of host threads [%(default)s] (accepted for compatibility)")
        p.add_argument("-v", metavar="<int>", dest="verbosity_level", type=int, default=1, choices=[0, 1, 2],
                       help="Verbosity level [%(default)s]")
        p.add_argument("--device", metavar="<int>", type=int, default=0, help="CUDA device index [%(default)s]")

    pf = sub.add_parser("prefilter", help="Prefilter genome pairs for alignment")
    pf.add_argument("-i", "--in", metavar="<file>", type=_input_path, dest="input_path", required=True)
    pf.add_argument("-o", "--out", metavar="<file>", type=pathlib.Path, dest="output_path", required=True)
    pf.add_argument("-k", "--k", metavar="<int>", type=int, default=25, choices=range(15, 31))
    pf.add_argument("--min-kmers", metavar="<int>", type=int, default=20)
    pf.add_argument("--min-ident", metavar="<float>", type=_ranged_float, default=0.7)
    pf.add_argument("--batch-size", metavar="<int>", type=int, default=0)
    pf.add_argument("--kmers-fraction", metavar="<float>", type=_ranged_float, default=1.0)
    pf.add_argument("--max-seqs", metavar="<int>", type=int, default=0)
    common(pf)

    al = sub.add_parser("align", help="Align genome sequence pairs and calculate ANI measures")
    al.add_argument("-i", "--in", metavar="<file>", type=_input_path, dest="input_path", required=True)
    al.add_argument("-o", "--out", metavar="<file>", type=pathlib.Path, dest="output_path", required=True)
    al.add_argument("--filter", metavar="<file>", type=_input_path, dest="filter_path")
    al.add_argument("--filter-threshold", metavar="<float>", dest="filter_threshold", type=_ranged_float, default=0)
    al.add_argument("--outfmt", metavar="<str>", choices=api.ALIGN_OUTFMT.keys(), dest="outfmt", default="standard")
    al.add_argument("--out-aln", metavar="<file>", type=pathlib.Path, dest="aln_path")
    for name in ("ani", "tani", "gani", "qcov", "rcov"):
        al.add_argument("--out-" + name, metavar="<float>", dest=name, type=_ranged_float, default=0)
    for name, default in (("mal", 11), ("msl", 7), ("mrd", 40), ("mqd", 40), ("reg", 35), ("aw", 15), ("am", 7), ("ar", 3)):
        al.add_argument("--" + name, metavar="<int>", type=int, default=default)
    common(al)
    return parser


def _fasta_inputs(args, parser):
    """vclust.py:685-702 validate_args_fasta_input"""
    args.is_multifasta = True
    args.fasta_paths = [args.input_path]
    if args.input_path.is_dir():
        args.is_multifasta = False
        args.fasta_paths = sorted(f for f in args.input_path.iterdir() if f.is_file())
    if not args.is_multifasta and len(args.fasta_paths) < 2:
        parser.error(f"Too few fasta files found in {args.input_path}. Expected at least 2, found {len(args.fasta_paths)}.")
    return args


def handle_prefilter(args, parser, logger) -> None:
    if args.batch_size and args.input_path.is_dir():        # vclust.py:731-736
        parser.error("--batch-size only handles a multi-fasta file, not a directory.")
    args = _fasta_inputs(args, parser)
    try:
        info = api.prefilter(args.fasta_paths, args.output_path, args.is_multifasta, kmer_size=args.k,
                             kmers_fraction=args.kmers_fraction, min_kmers=args.min_kmers, min_ident=args.min_ident,
                             max_seqs=args.max_seqs, batch_size=args.batch_size, device=args.device)
    except (api.VbError, ImportError) as e:
        logger.error(f"prefilter failed with message: {e}")
        sys.exit(1)
    logger.info("Completed: %d pairs, %.2f ms on the GPU" % (info.get("pairs", 0), info.get("total_ms", 0.0)))


def handle_align(args, parser, logger) -> None:
    args = _fasta_inputs(args, parser)
    try:
        info = api.align(args.fasta_paths, args.output_path, args.is_multifasta,
                         out_format=api.ALIGN_OUTFMT[args.outfmt], filter_file=args.filter_path,
                         filter_threshold=args.filter_threshold,
                         out_filters=dict(tani=args.tani, gani=args.gani, ani=args.ani, qcov=args.qcov, rcov=args.rcov),
                         mal=args.mal, msl=args.msl, mrd=args.mrd, mqd=args.mqd, reg=args.reg, aw=args.aw, am=args.am,
                         ar=args.ar, device=args.device, out_aln=args.aln_path)
    except (api.VbError, ImportError) as e:
        logger.error(f"align failed with message: {e}")
        sys.exit(1)
    logger.info("Completed: %d directed pairs, %.2f ms on the GPU" % (info.get("pairs", 0), info.get("total_ms", 0.0)))


def main(argv=None) -> None:
    parser = get_parser()
    args = parser.parse_args(argv)
    if not args.command:
        parser.print_help()
        return
    logger = logging.getLogger("vclust-b200")
    logger.handlers.clear()
    if args.verbosity_level:
        h = logging.StreamHandler(sys.stderr)
        h.setFormatter(logging.Formatter("%(asctime)s [%(levelname)-7s] %(message)s", "%Y-%m-%d %H:%M:%S"))
        logger.addHandler(h)
        logger.setLevel(logging.INFO)
    else:                                     # -v 0: nothing on stderr (reference test.py:359-360)
        logger.addHandler(logging.NullHandler())
        logger.setLevel(logging.CRITICAL + 1)
    {"prefilter": handle_prefilter, "align": handle_align}[args.command](args, parser, logger)


if __name__ == "__main__":
    main()
