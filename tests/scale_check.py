#!/usr/bin/env python
"""Full-size parity + timing check of one BASELINE configuration on a GPU box (not collected by pytest: minutes of CPU).

    python tests/scale_check.py c3            # 10 000 x 40 kb, families of 20 (95 000 candidate pairs)
    python tests/scale_check.py c3_s200 --no-ref
    python tests/scale_check.py n20k          # 20 000 x 10 kb: the hashed pair table (more than 11 585 genomes)
    python tests/scale_check.py n30k          # 30 000 x 40 kb: 1.2 x 10^9 k-mers, two prefilter passes

Runs `prefilter` + `align` through the file-level API (FASTA in, filter / ani.tsv / ids.tsv out), then the unmodified
reference tools from oracle/_ref on the same FASTA with all host cores, and compares the three output files byte for
byte.  Prints one JSON line with the sizes and the wall-clock times (ours: whole calls, including FASTA parsing, upload
and text output)."""
import argparse
import json
import os
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config")
    ap.add_argument("--no-ref", action="store_true")
    ap.add_argument("--n", type=int, default=0, help="only the first N genomes (whole families)")
    args = ap.parse_args()
    from oracle import oracle
    from vclust_b200 import api, synth
    kw = dict(synth.CONFIGS[args.config])
    if args.n:
        kw["n"] = args.n
    t0 = time.perf_counter()
    names, seqs = synth.make_genomes(**kw)
    out = {"config": args.config, "genomes": len(names), "bases": int(sum(s.size for s in seqs)),
           "generate_s": round(time.perf_counter() - t0, 2)}
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        fa = td / "in.fna"
        synth.write_fasta(fa, names, seqs)
        del seqs
        for rep in range(2):                       # second pass = warm (context, pool and arena sized)
            t0 = time.perf_counter()
            info_p = api.prefilter([fa], td / "fltr.txt", True)
            t1 = time.perf_counter()
            info_a = api.align([fa], td / "ani.tsv", True, filter_file=td / "fltr.txt", out_format=api.ALIGN_OUTFMT["complete"])
            t2 = time.perf_counter()
        flt = (td / "fltr.txt").read_text().splitlines()[1:]
        pairs = sum(ln.count(":") for ln in flt)
        out.update(candidate_pairs=pairs, ours_prefilter_s=round(t1 - t0, 3), ours_align_s=round(t2 - t1, 3),
                   ours_pairs_per_s=round(pairs / (t2 - t0), 1),
                   ours_gpu_ms={"prefilter": {k: round(v, 3) for k, v in info_p.items()},
                                "align": {k: round(v, 3) for k, v in info_a.items()}})
        if not args.no_ref and oracle.ref_available():
            thr = os.cpu_count() or 1
            t0 = time.perf_counter()
            oracle.ref_prefilter([fa], td / "ref_fltr.txt", td / "p", threads=thr)
            t1 = time.perf_counter()
            oracle.ref_align([fa], td / "ref_ani.tsv", td / "a", filter_path=td / "ref_fltr.txt", threads=thr,
                             columns=api.ALIGN_OUTFMT["complete"])
            t2 = time.perf_counter()
            same = {
                "filter": (td / "fltr.txt").read_bytes() == (td / "ref_fltr.txt").read_bytes(),
                "ani": (td / "ani.tsv").read_bytes() == (td / "ref_ani.tsv").read_bytes(),
                "ids": (td / "ani.ids.tsv").read_bytes() == (td / "ref_ani.ids.tsv").read_bytes(),
            }
            out.update(ref_threads=thr, ref_prefilter_s=round(t1 - t0, 2), ref_align_s=round(t2 - t1, 2),
                       ref_pairs_per_s=round(pairs / (t2 - t0), 1), byte_identical=same)
    print(json.dumps(out))
    if "byte_identical" in out and not all(out["byte_identical"].values()):
        sys.exit(1)


if __name__ == "__main__":
    main()
