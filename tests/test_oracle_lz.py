"""Pins oracle/lz_oracle.c (+ TSV formatting in oracle/oracle.py) against LZ-ANI's golden outputs."""
import gzip

import pytest

from oracle import oracle
from vclust_b200 import synth


def test_real_str():
    f = oracle.real_str
    assert f(0.0137651234, 6) == "0.0137651"
    assert f(0.016848, 6) == "0.016848"
    assert f(0.01684796, 6) == "0.0168480"
    assert f(1.0, 6) == "1"
    assert f(0.5, 6) == "0.5"
    assert f(100.0, 6) == "1e+02"
    assert f(89.28928929, 6) == "89.2893"
    assert f(0.00001, 6) == "1e-05"
    assert f(0.0001, 6) == "0.0001"
    assert f(0.99275, 4) == "0.9928"
    assert f(0.9999996, 6) == "1.00000"      # rounding carry keeps 6 digits (numeric_conversions.h:241-253)
    assert f(1234567.0, 6) == "1.23457e+06"


def test_example_all_vs_all_byte_exact(golden):
    names, codes = oracle.load_genomes_lzani([golden / "example" / "multifasta.fna.gz"], True)
    ani, ids, _ = oracle.align_text(names, codes)
    assert ids.encode() == (golden / "example" / "ani.ids.tsv").read_bytes()
    assert ani.encode() == (golden / "example" / "ani.tsv").read_bytes()


def test_vir61_ci_gate_byte_exact(golden):
    names, codes = oracle.load_genomes_lzani([golden / "vir61" / "vir61.fna.gz"], True)
    cols = "qidx,ridx,query,reference,tani,gani,ani,qcov,num_alns,len_ratio".split(",")   # lz-ani's own "standard"
    ani, ids, _ = oracle.align_text(names, codes, columns=cols)
    assert ids.encode() == (golden / "vir61" / "vir61.ani.ids.tsv").read_bytes()
    assert ani.encode() == (golden / "vir61" / "vir61.ani.tsv").read_bytes()


GEN = {
    "s60": dict(n=60, length=8000, family=6, seed=synth.BASE_SEED + 100, n_frac=0.2, lower_frac=0.2),
    "s40_k30": dict(n=40, length=(2000, 30000), family=5, seed=synth.BASE_SEED + 101, max_div=0.2),
    "s30_all": dict(n=30, length=(3000, 20000), family=3, seed=synth.BASE_SEED + 102, n_frac=0.3),
}
LZP = {
    "s60": {}, "s60_f02": {}, "s60_ms3": {},
    "s60_k15": dict(mal=9, msl=6, mrd=30, mqd=25, reg=30, aw=12, am=5, ar=2),
    "s40_k30": dict(mal=13, msl=8, mrd=60, mqd=50, reg=40, aw=20, am=9, ar=4),
    "s30_all": {},
}


@pytest.mark.parametrize("case", list(LZP))
def test_against_reference_binary_outputs(golden, tmp_path, case):
    gk = GEN[case] if case in GEN else GEN["s60"]
    names, seqs = synth.make_genomes(**gk)
    fa = tmp_path / "in.fna"
    synth.write_fasta(fa, names, seqs)
    n2, codes = oracle.load_genomes_lzani([fa], True)
    assert n2 == names
    adj = None
    if case != "s30_all":
        adj = oracle.read_filter(golden / "ref_synth" / (case + ".fltr.txt"), 0.0, names)
    ani, _, _ = oracle.align_text(names, codes, adj=adj, params=oracle.LzParams.default(**LZP[case]),
                                  columns=oracle.OUTFMT["complete"])
    assert ani.encode() == (golden / "ref_synth" / (case + ".ani.tsv")).read_bytes()


def test_example_alignment_regions_vs_golden(golden):
    """calc_regions / store_alignment restatement against example/output/ani.aln.tsv (row order in the reference depends
    on thread timing, so the rows are compared as a sorted list)."""
    names, codes = oracle.load_genomes_lzani([golden / "example" / "multifasta.fna.gz"], True)
    n = len(names)
    pr = [r for r in range(n) for q in range(n) if q != r]
    pq = [q for r in range(n) for q in range(n) if q != r]
    got = sorted(oracle.aln_lines(names, codes, pr, pq))
    want = gzip.open(golden / "example" / "ani.aln.tsv.gz", "rt").read().splitlines()
    assert want[0] == "query\treference\tpident\talnlen\tqstart\tqend\trstart\trend\tnt_match\tnt_mismatch"
    assert got == sorted(want[1:])
