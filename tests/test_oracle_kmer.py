"""Pins oracle/kmer_oracle.c (+ the text formatting in oracle/oracle.py) against the reference's golden vectors."""
import gzip

import numpy as np
import pytest

from oracle import oracle
from vclust_b200 import synth

# total-kmers rows of the reference's all2all.txt for example/multifasta.fna (SURVEY.md 8(c) KATs)
KAT_TOTALS = {
    (25, 1.0): [38557, 38607, 38908, 39682, 57222, 57392, 58459, 42629, 39537, 61292, 45598, 45598],
    (25, 0.2): [7828, 7829, 7892, 8027, 11362, 11465, 11589, 8477, 7840, 12263, 9085, 9085],
    (15, 1.0): [38433, 38486, 38676, 39553, 57131, 57299, 57913, 42613, 39387, 61046, 45565, 45565],
}
KAT_COMMON = {
    (25, 1.0): {(1, 0): 35785, (2, 0): 29648, (2, 1): 27579, (11, 10): 45550},
    (25, 0.2): {(1, 0): 7280, (2, 0): 6042, (2, 1): 5636},
    (15, 1.0): {(1, 0): 36784, (2, 0): 32815, (2, 1): 31446},
}


@pytest.fixture(scope="module")
def example_records(golden, tmp_path_factory):
    return oracle.read_records_kmerdb(golden / "example" / "multifasta.fna.gz")


@pytest.mark.parametrize("k,f", list(KAT_TOTALS))
def test_example_totals_and_commons(example_records, k, f):
    sets = oracle.kmer_sets([[s] for _, s in example_records], k, f)
    assert [int(s.size) for s in sets] == KAT_TOTALS[(k, f)]
    rows, cols, vals = oracle.common_matrix(sets)
    got = {(int(r), int(c)): int(v) for r, c, v in zip(rows, cols, vals)}
    for key, v in KAT_COMMON[(k, f)].items():
        assert got[key] == v


def test_example_filter_file_byte_exact(golden, example_records):
    names = [n for n, _ in example_records]
    sets = oracle.kmer_sets([[s] for _, s in example_records], 25, 1.0)
    txt = oracle.filter_text(names, sets, 25, 1.0, 20, 0.7)
    assert txt.encode() == (golden / "example" / "fltr.txt").read_bytes()


def test_kmerdb_synth_kat(golden):
    recs = oracle.read_records_kmerdb(golden / "kmerdb_synth" / "synth.fa")
    assert [n for n, _ in recs] == list("ABCDE")
    sets = oracle.kmer_sets([[s] for _, s in recs], 21, 1.0)
    assert [int(s.size) for s in sets] == [80, 80, 39, 31, 80]
    rows, cols, vals = oracle.common_matrix(sets)
    got = {(int(r), int(c)): int(v) for r, c, v in zip(rows, cols, vals)}
    assert got == {(1, 0): 36, (2, 0): 13, (2, 1): 34, (4, 0): 80, (4, 1): 36, (4, 2): 13}


def test_minhash_spec():
    # independent pure-Python restatement of K/filter.h:96-115
    M = (1 << 64) - 1

    def fmix(k):
        k ^= k >> 33; k = k * 0xff51afd7ed558ccd & M
        k ^= k >> 33; k = k * 0xc4ceb9fe1a85ec53 & M
        k ^= k >> 33
        return k

    def py_hash(kmer, k):
        c = -(-k // 4)
        h = kmer * 0x87c37b91114253d5 & M
        h = ((h << 31) | (h >> 33)) & M
        h = h * 0x4cf5ad432745937f & M
        h1 = 42 ^ h ^ c
        h2 = 42 ^ c
        h1 = (h1 + h2) & M; h2 = (h2 + h1) & M
        h1 = fmix(h1); h2 = fmix(h2)
        h1 = (h1 + h2) & M; h2 = (h2 + h1) & M
        return h1 ^ h2

    rng = np.random.default_rng(1)
    L = oracle.lib()
    for k in (15, 21, 25, 30):
        for x in rng.integers(0, 1 << 50, size=50):
            assert L.kmo_minhash(int(x), k) == py_hash(int(x), k)


@pytest.mark.parametrize("case,kw", [
    ("s60", dict(k=25, fraction=1.0, min_kmers=20, min_ident=0.7)),
    ("s60_k15", dict(k=15, fraction=1.0, min_kmers=10, min_ident=0.5)),
    ("s60_f02", dict(k=25, fraction=0.2, min_kmers=4, min_ident=0.7)),
    ("s40_k30", dict(k=30, fraction=1.0, min_kmers=1, min_ident=0.3)),
    ("s60_ms3", dict(k=25, fraction=1.0, min_kmers=20, min_ident=0.7, max_seqs=3)),
])
def test_against_reference_binary_outputs(golden, tmp_path, case, kw):
    gen = {
        "s60": dict(n=60, length=8000, family=6, seed=synth.BASE_SEED + 100, n_frac=0.2, lower_frac=0.2),
        "s40_k30": dict(n=40, length=(2000, 30000), family=5, seed=synth.BASE_SEED + 101, max_div=0.2),
    }
    names, seqs = synth.make_genomes(**gen["s40_k30" if case == "s40_k30" else "s60"])
    fa = tmp_path / "in.fna"
    synth.write_fasta(fa, names, seqs)
    txt = oracle.prefilter_text_from_fasta([fa], True, **kw)
    assert txt.encode() == (golden / "ref_synth" / (case + ".fltr.txt")).read_bytes()


def test_fixed6():
    assert oracle.fixed6(0.99847951) == "0.998480"
    assert oracle.fixed6(1.0) == "1.000000"
    assert oracle.fixed6(0.0000004) == "0.000000"
    assert oracle.fixed6(0.7) == "0.700000"
