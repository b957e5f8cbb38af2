#!/usr/bin/env python
"""Regenerates tests/golden/ from the read-only reference checkout (run in the build container only).

Two kinds of fixtures:
 1. the reference's own golden vectors / test inputs, copied as DATA (no source code):
      example/multifasta.fna.gz + example/output/{fltr.txt,ani.tsv,ani.ids.tsv,ani.aln.tsv}
      3rd_party/lz-ani/test/vir61/*.fna (concatenated, sorted by file name) + vir61.ani.tsv / .ids.tsv  (LZ-ANI CI gate)
      3rd_party/kmer-db/test/synth/{synth.fa,a2a-sparse}                                            (k=21 KAT)
 2. outputs of the UNMODIFIED reference binaries (oracle/_ref, built by oracle/build_ref.sh) on small seeded
    synthetic genome sets from vclust_b200.synth -- these pin flag combinations the reference goldens do not
    cover (k=15 / k=30, --kmers-fraction 0.2, --max-seqs, N runs + lower case, non-default LZ parameters).
"""
import gzip
import shutil
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
REF = Path("/root/reference")

from oracle import oracle  # noqa: E402
from vclust_b200 import synth  # noqa: E402


def gz_write(path: Path, data: bytes):
    with open(path, "wb") as raw, gzip.GzipFile(filename="", mode="wb", fileobj=raw, mtime=0) as fh:
        fh.write(data)


def main():
    """No arguments: everything.  Arguments: only the named ref_synth cases (kmer-db build at k=30 needs ~60 GB of RAM,
    so s40_k30 can only be regenerated on a large host)."""
    assert REF.exists(), "needs the reference checkout"
    oracle.build_ref()
    only = set(sys.argv[1:])
    if not only:
        copy_reference_goldens()
    make_ref_synth(only)


def copy_reference_goldens():
    ex = HERE / "example"
    ex.mkdir(exist_ok=True)
    shutil.copy(REF / "example/multifasta.fna.gz", ex / "multifasta.fna.gz")
    for f in ("fltr.txt", "ani.tsv", "ani.ids.tsv"):
        shutil.copy(REF / "example/output" / f, ex / f)
    gz_write(ex / "ani.aln.tsv.gz", (REF / "example/output/ani.aln.tsv").read_bytes())

    v = HERE / "vir61"
    v.mkdir(exist_ok=True)
    files = sorted((REF / "3rd_party/lz-ani/test/vir61").iterdir())
    gz_write(v / "vir61.fna.gz", b"".join(p.read_bytes() for p in files))
    shutil.copy(REF / "3rd_party/lz-ani/test/vir61.ani.tsv", v / "vir61.ani.tsv")
    shutil.copy(REF / "3rd_party/lz-ani/test/vir61.ani.ids.tsv", v / "vir61.ani.ids.tsv")

    ks = HERE / "kmerdb_synth"
    ks.mkdir(exist_ok=True)
    for f in ("synth.fa", "a2a-sparse"):
        shutil.copy(REF / "3rd_party/kmer-db/test/synth" / f, ks / f)



def make_ref_synth(only):
    # ---- reference-binary outputs on seeded synthetic sets
    rs = HERE / "ref_synth"
    rs.mkdir(exist_ok=True)
    cases = {
        # name: (generator kwargs, prefilter kwargs, lz params)
        "s60": (dict(n=60, length=8000, family=6, seed=synth.BASE_SEED + 100, n_frac=0.2, lower_frac=0.2),
                dict(k=25, fraction=1.0, min_kmers=20, min_ident=0.7), {}),
        "s60_k15": (dict(n=60, length=8000, family=6, seed=synth.BASE_SEED + 100, n_frac=0.2, lower_frac=0.2),
                    dict(k=15, fraction=1.0, min_kmers=10, min_ident=0.5), dict(mal=9, msl=6, mrd=30, mqd=25, reg=30, aw=12, am=5, ar=2)),
        "s60_f02": (dict(n=60, length=8000, family=6, seed=synth.BASE_SEED + 100, n_frac=0.2, lower_frac=0.2),
                    dict(k=25, fraction=0.2, min_kmers=4, min_ident=0.7), {}),
        "s40_k30": (dict(n=40, length=(2000, 30000), family=5, seed=synth.BASE_SEED + 101, max_div=0.2),
                    dict(k=30, fraction=1.0, min_kmers=1, min_ident=0.3), dict(mal=13, msl=8, mrd=60, mqd=50, reg=40, aw=20, am=9, ar=4)),
        # --max-seqs 3: families of 6 => every row is cut from 5 to 3 items, the filter gets entries on both sides of
        # the diagonal and lz-ani reports a pair once per row that kept it
        "s60_ms3": (dict(n=60, length=8000, family=6, seed=synth.BASE_SEED + 100, n_frac=0.2, lower_frac=0.2),
                    dict(k=25, fraction=1.0, min_kmers=20, min_ident=0.7, max_seqs=3), {}),
    }
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        for name, (gk, pk, lk) in cases.items():
            if only and name not in only:
                continue
            names, seqs = synth.make_genomes(**gk)
            fa = td / (name + ".fna")
            synth.write_fasta(fa, names, seqs)
            flt = rs / (name + ".fltr.txt")
            oracle.ref_prefilter([fa], flt, td / (name + "_p"), **pk)
            lp = oracle.LzParams.default(**lk)
            oracle.ref_align([fa], rs / (name + ".ani.tsv"), td / (name + "_a"), filter_path=flt, params=lp,
                             columns=oracle.OUTFMT["complete"])
            (rs / (name + ".ani.ids.tsv")).unlink()
        # one unfiltered all-vs-all (many unrelated pairs) with default parameters
        if not only or "s30_all" in only:
            names, seqs = synth.make_genomes(n=30, length=(3000, 20000), family=3, seed=synth.BASE_SEED + 102, n_frac=0.3)
            fa = td / "s30_all.fna"
            synth.write_fasta(fa, names, seqs)
            oracle.ref_align([fa], rs / "s30_all.ani.tsv", td / "s30_a", columns=oracle.OUTFMT["complete"])
            (rs / "s30_all.ani.ids.tsv").unlink()
    (rs / "CASES.txt").write_text("\n".join("%s\t%r\t%r\t%r" % (k, *v) for k, v in cases.items()) +
                                  "\ns30_all\t%r\t-\t{}\n" % dict(n=30, length=(3000, 20000), family=3,
                                                                 seed=synth.BASE_SEED + 102, n_frac=0.3))


if __name__ == "__main__":
    main()
