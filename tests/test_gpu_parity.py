"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI, against the committed
golden vectors and against the CPU oracle on seeded inputs.  Integer/byte results must be bit-exact."""
import gzip

import numpy as np
import pytest

from oracle import oracle
from vclust_b200 import api, synth

pytestmark = pytest.mark.gpu

GEN = {
    "s60": dict(n=60, length=8000, family=6, seed=synth.BASE_SEED + 100, n_frac=0.2, lower_frac=0.2),
    "s40_k30": dict(n=40, length=(2000, 30000), family=5, seed=synth.BASE_SEED + 101, max_div=0.2),
    "s30_all": dict(n=30, length=(3000, 20000), family=3, seed=synth.BASE_SEED + 102, n_frac=0.3),
}
PRE = {
    "s60": dict(k=25, kmers_fraction=1.0, min_kmers=20, min_ident=0.7),
    "s60_k15": dict(k=15, kmers_fraction=1.0, min_kmers=10, min_ident=0.5),
    "s60_f02": dict(k=25, kmers_fraction=0.2, min_kmers=4, min_ident=0.7),
    "s40_k30": dict(k=30, kmers_fraction=1.0, min_kmers=1, min_ident=0.3),
    "s60_ms3": dict(k=25, kmers_fraction=1.0, min_kmers=20, min_ident=0.7, max_seqs=3),
}
LZP = {
    "s60": {}, "s60_f02": {}, "s60_ms3": {},
    "s60_k15": dict(mal=9, msl=6, mrd=30, mqd=25, reg=30, aw=12, am=5, ar=2),
    "s40_k30": dict(mal=13, msl=8, mrd=60, mqd=50, reg=40, aw=20, am=9, ar=4),
    "s30_all": {},
}


@pytest.fixture(scope="module")
def ctx():
    with api.Context(0) as c:
        yield c


def _synth_fasta(tmp_path, case):
    names, seqs = synth.make_genomes(**GEN["s40_k30" if case == "s40_k30" else ("s30_all" if case == "s30_all" else "s60")])
    fa = tmp_path / (case + ".fna")
    synth.write_fasta(fa, names, seqs)
    return fa, names, seqs


# ---------------------------------------------------------------- prefilter
def test_prefilter_example_byte_exact(ctx, golden, tmp_path):
    g = api.Genomes.load([golden / "example" / "multifasta.fna.gz"], True, api.FASTA_KMERDB)
    pairs = api.prefilter_genomes(ctx, g)
    assert pairs.total_kmers.tolist() == [38557, 38607, 38908, 39682, 57222, 57392, 58459, 42629, 39537, 61292, 45598, 45598]
    got = {(int(r), int(c)): int(v) for r, c, v in zip(pairs.rows, pairs.cols, pairs.common)}
    assert got[(1, 0)] == 35785 and got[(11, 10)] == 45550
    out = tmp_path / "fltr.txt"
    api.write_filter(g, pairs, out)
    assert out.read_bytes() == (golden / "example" / "fltr.txt").read_bytes()


@pytest.mark.parametrize("case", list(PRE))
def test_prefilter_vs_reference_binary_outputs(ctx, golden, tmp_path, case):
    fa, names, seqs = _synth_fasta(tmp_path, case)
    out = tmp_path / "fltr.txt"
    kw = PRE[case]
    api.prefilter([fa], out, True, kmer_size=kw["k"], kmers_fraction=kw["kmers_fraction"], min_kmers=kw["min_kmers"],
                  min_ident=kw["min_ident"], max_seqs=kw.get("max_seqs", 0))
    assert out.read_bytes() == (golden / "ref_synth" / (case + ".fltr.txt")).read_bytes()


@pytest.mark.parametrize("k,f", [(25, 1.0), (25, 0.2), (15, 1.0), (21, 0.5), (31, 1.0), (18, 1.0)])
def test_prefilter_counts_vs_oracle(ctx, golden, k, f):
    recs = oracle.read_records_kmerdb(golden / "example" / "multifasta.fna.gz")
    sets = oracle.kmer_sets([[s] for _, s in recs], k, f)
    rows, cols, vals = oracle.common_matrix(sets)
    g = api.Genomes.load([golden / "example" / "multifasta.fna.gz"], True, api.FASTA_KMERDB)
    pairs = api.prefilter_genomes(ctx, g, k=k, min_kmers=1, min_ident=0.0, kmers_fraction=f)
    assert pairs.total_kmers.tolist() == [int(s.size) for s in sets]
    want = {(int(r), int(c)): int(v) for r, c, v in zip(rows, cols, vals)}
    got = {(int(r), int(c)): int(v) for r, c, v in zip(pairs.rows, pairs.cols, pairs.common)}
    assert got == want


@pytest.mark.parametrize("env", [
    {"VB_PREFILTER_CHUNK": "4096"},                                  # slot axis walked in many small chunks (> 2^32-slot inputs)
    {"VB_PREFILTER_CHUNK": "6144", "VB_PREFILTER_EXACT": "1"},       # + survivors counted before the list is allocated
    {"VB_PREFILTER_SEEN": "0"},                                      # singleton screen off (large inputs)
    {"VB_PREFILTER_SEEN": "0", "VB_PREFILTER_EXACT": "1", "VB_PREFILTER_CHUNK": "2048"},
    {"VB_PREFILTER_SEEN": "12"},                                     # tiny seen table: heavy slot collisions must be harmless
    {"VB_PREFILTER_HASH": "1"},                                      # hashed pair table (N(N-1)/2 > 2^26)
    {"VB_PREFILTER_HASH": "1", "VB_PREFILTER_TABLE": "1024"},        # hashed pair table far too small: overflow -> the call is redone larger
    {"VB_PREFILTER_HASH": "1", "VB_PREFILTER_TABLE": "4096", "VB_PREFILTER_PASSES": "4"},   # ... grown between passes while it fills up
    {"VB_PREFILTER_PASSES": "3"},                                    # > 10^9 k-mers: several passes over k-mer hash shards
    {"VB_PREFILTER_PASSES": "2", "VB_PREFILTER_HASH": "1", "VB_PREFILTER_SEEN": "0"},
    {"VB_PREFILTER_BUCKET": "flat"},                                 # grouping kernel for large families (picked from the size hint)
    {"VB_PREFILTER_BUCKET": "flat", "VB_PREFILTER_HASH": "1", "VB_PREFILTER_SEEN": "0"},
    {"VB_PREFILTER_BUCKET": "chain", "VB_PREFILTER_HASH": "1", "VB_PREFILTER_SEEN": "0"},
    {"VB_PREFILTER_L1BITS": "2"},                                    # lopsided fan-out: few parents, many sub-buckets
    {"VB_PREFILTER_L1BITS": "9", "VB_PREFILTER_SEEN": "0"},         # ... and the reverse
])
def test_prefilter_large_input_paths_vs_oracle(ctx, golden, monkeypatch, env):
    """The code paths that only very large inputs select by themselves, forced on a small input: same integers."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    names, seqs = synth.make_genomes(n=36, length=(3000, 9000), family=6, seed=31, n_frac=0.2, lower_frac=0.2)
    raw = [s.tobytes() for s in seqs] + [b"", b"ACGTAC", b"N" * 500]
    names = names + ["e0", "e1", "e2"]
    for k, f in ((25, 1.0), (17, 0.5)):
        sets = oracle.kmer_sets([[s] for s in raw], k, f)
        rows, cols, vals = oracle.common_matrix(sets)
        g = api.Genomes.from_memory(names, raw)
        pairs = api.prefilter_genomes(ctx, g, k=k, min_kmers=1, min_ident=0.0, kmers_fraction=f)
        assert pairs.total_kmers.tolist() == [int(s.size) for s in sets]
        assert {(int(r), int(c)): int(v) for r, c, v in zip(pairs.rows, pairs.cols, pairs.common)} == \
               {(int(r), int(c)): int(v) for r, c, v in zip(rows, cols, vals)}


@pytest.mark.parametrize("env", [{}, {"VB_PREFILTER_BUCKET": "flat"}, {"VB_PREFILTER_SEEN": "0"},
                                 {"VB_PREFILTER_SEEN": "0", "VB_PREFILTER_BUCKET": "flat"}, {"VB_PREFILTER_HASH": "1", "VB_PREFILTER_SEEN": "0"}])
def test_prefilter_kmers_repeated_inside_genomes(ctx, monkeypatch, env):
    """A k-mer counts once per genome however often it occurs there (kmer-db sorts and de-duplicates a sample's k-mers):
    families whose members carry a block several times, forward and reverse-complemented, with some copies mutated --
    the grouping kernels must tell repeats of one genome from members of the group."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    rng = np.random.default_rng(77)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    comp = np.zeros(256, dtype=np.uint8)
    comp[list(b"ACGT")] = list(b"TGCA")
    seqs = []
    for fam in range(12):
        block = acgt[rng.integers(0, 4, size=400)]
        backbone = acgt[rng.integers(0, 4, size=3000)]
        for member in range(9):
            parts = [backbone[:1000].copy()]
            for rep in range(int(rng.integers(1, 5))):                   # 1 to 4 copies of the block, some reversed, some mutated
                b = block.copy()
                if rng.random() < 0.5:
                    b[rng.integers(0, b.size, size=3)] = acgt[rng.integers(0, 4, size=3)]
                parts.append(comp[b[::-1]] if rng.random() < 0.4 else b)
                parts.append(acgt[rng.integers(0, 4, size=int(rng.integers(0, 60)))])
            parts.append(backbone[1000:].copy())
            s = np.concatenate(parts)
            s[rng.integers(0, s.size, size=member * 6)] = acgt[rng.integers(0, 4, size=member * 6)]
            seqs.append(s.tobytes())
    names = ["r%d" % i for i in range(len(seqs))]
    for k in (25, 15):
        sets = oracle.kmer_sets([[s] for s in seqs], k, 1.0)
        rows, cols, vals = oracle.common_matrix(sets)
        g = api.Genomes.from_memory(names, seqs)
        pairs = api.prefilter_genomes(ctx, g, k=k, min_kmers=1, min_ident=0.0)
        assert pairs.total_kmers.tolist() == [int(s.size) for s in sets]
        assert {(int(r), int(c)): int(v) for r, c, v in zip(pairs.rows, pairs.cols, pairs.common)} == \
               {(int(r), int(c)): int(v) for r, c, v in zip(rows, cols, vals)}
        pairs.close(); g.close()


@pytest.mark.parametrize("env", [{}, {"VB_PREFILTER_SEEN": "0"}, {"VB_PREFILTER_SEEN": "0", "VB_PREFILTER_NO_EARLY": "1"}])
def test_chunked_pinned_upload_gives_the_same_result(ctx, monkeypatch, env):
    """From the second upload of a host set on, the genomes travel in several chunks on a copy stream and the screen
    pass of each chunk -- without a screen (large inputs): the whole extraction of each chunk -- is enqueued while the
    next chunk is in flight (e2e path of bench.py): same pairs, same counts, and the align stage finds the store the
    prefilter uploaded."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    names, seqs = synth.make_genomes(n=320, length=40_000, family=8, seed=5, n_frac=0.05, lower_frac=0.05)
    g = api.Genomes.from_memory(names, seqs)
    runs = []
    for _ in range(3):                      # 1st: pageable, one chunk; 2nd, 3rd: pinned, chunked
        ctx.evict()
        pairs = api.prefilter_genomes(ctx, g)
        res = api.align_genomes(ctx, g, pairs)
        runs.append((pairs.rows.tolist(), pairs.cols.tolist(), pairs.common.tolist(), pairs.total_kmers.tolist(),
                     res.ref.tolist(), res.qry.tolist(), res.stats.tolist()))
        pairs.close(); res.close()
    ctx.evict()
    assert runs[0] == runs[1] == runs[2]
    assert len(runs[0][0]) == 320 * 7 // 2
    # spot-check against the oracle on two families
    sub = list(range(16))
    sets = oracle.kmer_sets([[seqs[i].tobytes()] for i in sub], 25, 1.0)
    want = {(r, c): v for r, c, v, _ in oracle.prefilter_pairs(sets, 25, 20, 0.7)}
    got = {(r, c): v for r, c, v in zip(*runs[0][:3]) if r < 16 and c < 16}
    assert got == want


@pytest.mark.parametrize("env", [{}, {"VB_PREFILTER_HASH": "1"}])
def test_prefilter_kmers_shared_by_thousands_of_genomes(ctx, monkeypatch, env):
    """A k-mer present in more genomes than a shared-memory bucket holds (kmer-db's "bubble" case, c5) takes the generic
    global-memory path of the grouping kernel: 2 200 copies of one short sequence + unrelated genomes."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    rng = np.random.default_rng(17)
    core = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=140)].tobytes()
    others = [np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=300)].tobytes() for _ in range(40)]
    seqs = [core] * 2200 + others
    names = ["g%d" % i for i in range(len(seqs))]
    g = api.Genomes.from_memory(names, seqs)
    pairs = api.prefilter_genomes(ctx, g, k=21, min_kmers=1, min_ident=0.0)
    sets = oracle.kmer_sets([[core]] + [[s] for s in others], 21, 1.0)
    n_core = int(sets[0].size)
    assert pairs.total_kmers.tolist() == [n_core] * 2200 + [int(s.size) for s in sets[1:]]
    rows, cols, common = pairs.rows.astype(np.int64), pairs.cols.astype(np.int64), pairs.common
    among = (rows < 2200) & (cols < 2200)
    assert int(among.sum()) == 2200 * 2199 // 2 and bool((common[among] == n_core).all())
    # everything that involves an unrelated genome: as the oracle says for (core, others)
    r2, c2, v2 = oracle.common_matrix(sets)
    want = {}
    for r, c, v in zip(r2.tolist(), c2.tolist(), v2.tolist()):
        for rr in ([r + 2199] if r > 0 else range(2200)):
            for cc in ([c + 2199] if c > 0 else range(2200)):
                if rr != cc:
                    want[(max(rr, cc), min(rr, cc))] = v
    got = {(int(r), int(c)): int(v) for r, c, v in zip(rows[~among], cols[~among], common[~among])}
    assert got == {k: v for k, v in want.items() if not (k[0] < 2200 and k[1] < 2200)}


def test_prefilter_edge_cases(ctx):
    # empty genome, genome shorter than k, all-N genome, U handled as T, lower case, duplicate genomes
    seqs = [b"", b"ACGTACGT", b"N" * 100, b"ACGU" * 30, b"acgt" * 30, b"ACGT" * 30, b"ACGTTGCAAGGCTA" * 10]
    names = ["g%d" % i for i in range(len(seqs))]
    g = api.Genomes.from_memory(names, seqs)
    pairs = api.prefilter_genomes(ctx, g, k=15, min_kmers=1, min_ident=0.0)
    sets = oracle.kmer_sets([[s] for s in seqs], 15, 1.0)
    assert pairs.total_kmers.tolist() == [int(s.size) for s in sets]
    rows, cols, vals = oracle.common_matrix(sets)
    assert {(int(r), int(c)): int(v) for r, c, v in zip(pairs.rows, pairs.cols, pairs.common)} == \
           {(int(r), int(c)): int(v) for r, c, v in zip(rows, cols, vals)}


# ---------------------------------------------------------------- align
def test_align_example_all_vs_all_byte_exact(ctx, golden, tmp_path):
    out = tmp_path / "ani.tsv"
    api.align([golden / "example" / "multifasta.fna.gz"], out, True)
    assert (tmp_path / "ani.ids.tsv").read_bytes() == (golden / "example" / "ani.ids.tsv").read_bytes()
    assert out.read_bytes() == (golden / "example" / "ani.tsv").read_bytes()


def test_align_vir61_ci_gate_byte_exact(ctx, golden, tmp_path):
    out = tmp_path / "vir61.ani.tsv"
    cols = "qidx,ridx,query,reference,tani,gani,ani,qcov,num_alns,len_ratio".split(",")
    api.align([golden / "vir61" / "vir61.fna.gz"], out, True, out_format=cols)
    assert (tmp_path / "vir61.ani.ids.tsv").read_bytes() == (golden / "vir61" / "vir61.ani.ids.tsv").read_bytes()
    assert out.read_bytes() == (golden / "vir61" / "vir61.ani.tsv").read_bytes()


@pytest.mark.parametrize("case", list(LZP))
def test_align_vs_reference_binary_outputs(ctx, golden, tmp_path, case):
    fa, names, seqs = _synth_fasta(tmp_path, case)
    out = tmp_path / "ani.tsv"
    flt = None if case == "s30_all" else golden / "ref_synth" / (case + ".fltr.txt")
    api.align([fa], out, True, out_format=api.ALIGN_OUTFMT["complete"], filter_file=flt, **LZP[case])
    assert out.read_bytes() == (golden / "ref_synth" / (case + ".ani.tsv")).read_bytes()


def test_align_reference_batches_vs_single_batch(ctx, monkeypatch, golden, tmp_path):
    """When the reference texts + anchor tables of a call exceed the memory budget they are built batch by batch
    (c4 on one GPU); forced here with a 1 MB budget: same statistics, same files, regions included."""
    names, seqs = synth.make_genomes(n=40, length=(3000, 9000), family=5, seed=8, n_frac=0.2)
    raw = [s.tobytes() for s in seqs]
    rng = np.random.default_rng(3)
    ref = rng.integers(0, 40, size=400)
    qry = (ref // 5) * 5 + rng.integers(0, 5, size=400)
    qry[:50] = rng.integers(0, 40, size=50)
    g = api.Genomes.from_memory(names, raw)
    want = api.align_pairs(ctx, g, ref, qry)
    monkeypatch.setenv("VB_ALIGN_BUDGET_MB", "1")
    got = api.align_pairs(ctx, g, ref, qry)
    assert ctx.timing("align.batches") > 3
    assert np.array_equal(got, want)
    out = tmp_path / "ani.tsv"
    aln = tmp_path / "ani.aln.tsv"
    monkeypatch.setenv("VB_ALIGN_BUDGET_MB", "4")           # 64 kb genomes: ~1.3 MB per reference
    api.align([golden / "example" / "multifasta.fna.gz"], out, True, out_aln=aln)
    assert out.read_bytes() == (golden / "example" / "ani.tsv").read_bytes()
    want_aln = gzip.open(golden / "example" / "ani.aln.tsv.gz", "rt").read().splitlines()
    got_aln = aln.read_text().splitlines()
    assert got_aln[0] == want_aln[0] and sorted(got_aln[1:]) == sorted(want_aln[1:])


def test_align_pairs_vs_oracle_random(ctx):
    names, seqs = synth.make_genomes(n=48, length=(1500, 12000), family=4, seed=99, max_div=0.25, n_frac=0.3, lower_frac=0.3)
    # add degenerate genomes: empty, tiny, all N, a homopolymer and an exact duplicate
    extra = [b"", b"ACG", b"N" * 300, b"A" * 2000, seqs[0].tobytes()]
    seqs = [s.tobytes() for s in seqs] + extra
    names = names + ["x%d" % i for i in range(len(extra))]
    rng = np.random.default_rng(5)
    n = len(names)
    ref = rng.integers(0, n, size=1500)
    qry = rng.integers(0, n, size=1500)
    # make sure related pairs and self pairs are present
    ref[:200] = (np.arange(200) % 48)
    qry[:200] = (ref[:200] // 4) * 4 + rng.integers(0, 4, size=200)
    g = api.Genomes.from_memory(names, seqs)
    got = api.align_pairs(ctx, g, ref, qry)
    want = oracle.run_pairs([oracle.lz_codes(s) for s in seqs], ref, qry)
    bad = np.nonzero((got != want).any(axis=1))[0]
    assert bad.size == 0, "first mismatches: %s" % [(int(ref[i]), int(qry[i]), got[i].tolist(), want[i].tolist()) for i in bad[:5]]


@pytest.mark.parametrize("params", [
    dict(mal=9, msl=5, mrd=20, mqd=60, reg=20, aw=8, am=3, ar=1),
    dict(mal=16, msl=12, mrd=100, mqd=10, reg=60, aw=32, am=20, ar=6),
    dict(mal=11, msl=7, mrd=40, mqd=40, reg=35, aw=15, am=15, ar=3),      # window rule can never fire
])
def test_align_pairs_vs_oracle_params(ctx, params):
    names, seqs = synth.make_genomes(n=24, length=(2000, 9000), family=4, seed=123, max_div=0.2, n_frac=0.2)
    seqs = [s.tobytes() for s in seqs]
    n = len(names)
    ref, qry = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    ref, qry = ref.ravel(), qry.ravel()
    g = api.Genomes.from_memory(names, seqs)
    got = api.align_pairs(ctx, g, ref, qry, api.align_params(**params))
    want, oob = oracle.run_pairs([oracle.lz_codes(s) for s in seqs], ref, qry, oracle.LzParams.default(**params), return_oob=True)
    # With mqd > mrd the reference's tail comparison can leave the reference text (L/parser.cpp:713 -> :210-248, no bounds
    # check) and reads whatever the vector's capacity holds: those pairs are undefined in the reference and are excluded;
    # every pair whose parse stays inside the text must agree.
    assert params["mqd"] > params["mrd"] or not oob.any()
    assert oob.mean() < 0.5
    bad = np.nonzero((got != want).any(axis=1) & ~oob)[0]
    assert bad.size == 0, "first mismatches: %s" % [(int(ref[i]), int(qry[i]), got[i].tolist(), want[i].tolist()) for i in bad[:5]]


# ---------------------------------------------------------------- --out-aln (alignment regions)
def test_out_aln_example_vs_golden(ctx, golden, tmp_path):
    import gzip
    api.align([golden / "example" / "multifasta.fna.gz"], tmp_path / "ani.tsv", True, out_aln=tmp_path / "ani.aln.tsv")
    assert (tmp_path / "ani.tsv").read_bytes() == (golden / "example" / "ani.tsv").read_bytes()
    got = (tmp_path / "ani.aln.tsv").read_text().splitlines()
    want = gzip.open(golden / "example" / "ani.aln.tsv.gz", "rt").read().splitlines()
    assert got[0] == want[0]
    assert sorted(got[1:]) == sorted(want[1:])


@pytest.mark.parametrize("params", [
    {},
    dict(mal=9, msl=5, mrd=20, mqd=60, reg=20, aw=8, am=3, ar=1),
    dict(mal=13, msl=8, mrd=60, mqd=50, reg=40, aw=20, am=9, ar=4),
])
def test_regions_vs_oracle(ctx, params):
    names, seqs = synth.make_genomes(n=32, length=(1500, 12000), family=4, seed=321, max_div=0.25, n_frac=0.3, lower_frac=0.2)
    seqs = [s.tobytes() for s in seqs] + [b"", b"ACGTTGCA" * 40, b"N" * 200]
    names = names + ["e0", "e1", "e2"]
    n = len(names)
    rng = np.random.default_rng(17)
    ref = rng.integers(0, n, size=700)
    qry = rng.integers(0, n, size=700)
    ref[:300] = np.arange(300) % 32
    qry[:300] = (ref[:300] // 4) * 4 + rng.integers(0, 4, size=300)
    g = api.Genomes.from_memory(names, seqs)
    st, regions = api.align_pairs_regions(ctx, g, ref, qry, api.align_params(**params))
    codes = [oracle.lz_codes(s) for s in seqs]
    want_st, want_regs = oracle.run_pairs_regions(codes, ref, qry, oracle.LzParams.default(**params))
    assert np.array_equal(st, want_st)
    assert np.array_equal(st, api.align_pairs(ctx, g, ref, qry, api.align_params(**params)))    # both kernel variants agree
    tab = regions.table()
    want_rows = []
    for k, rg in enumerate(want_regs):                       # oracle columns: ref_start, ref_end, seq_start, seq_end, m, mm
        for rs, re_, ss, se, m, mm in rg.tolist():
            want_rows.append((int(ref[k]), int(qry[k]), ss, se, rs, re_, m, mm))
    # duplicates of a (ref, qry) pair in the input list produce duplicate regions on both sides; compare as multisets
    assert sorted(map(tuple, tab.tolist())) == sorted(want_rows)


# ---------------------------------------------------------------- multi-GPU building blocks (on one GPU)
@pytest.mark.parametrize("world", [2, 3])
def test_prefilter_partial_shards_sum_to_full(ctx, golden, world):
    g = api.Genomes.load([golden / "example" / "multifasta.fna.gz"], True, api.FASTA_KMERDB)
    full = api.prefilter_genomes(ctx, g, k=25, min_kmers=20, min_ident=0.7)
    rows, cols, vals, totals = [], [], [], np.zeros(len(g), dtype=np.int64)
    for r in range(world):
        part = api.prefilter_partial(ctx, g, r, world, k=25)
        rows.append(part.rows); cols.append(part.cols); vals.append(part.common)
        totals += part.total_kmers
    assert totals.tolist() == full.total_kmers.tolist()
    m = api.merge_pairs(np.concatenate(rows), np.concatenate(cols), np.concatenate(vals), totals.astype(np.uint32),
                        k=25, min_kmers=20, min_ident=0.7)
    assert list(zip(m.rows.tolist(), m.cols.tolist(), m.common.tolist())) == \
           list(zip(full.rows.tolist(), full.cols.tolist(), full.common.tolist()))
    assert np.array_equal(m.ani, full.ani)


def test_align_result_assembly_matches_vb_align(ctx, golden, tmp_path):
    p = golden / "example" / "multifasta.fna.gz"
    g = api.Genomes.load([p], True, api.FASTA_LZANI)
    pairs = api.read_filter(golden / "example" / "fltr.txt", 0.0, g)
    ref = np.concatenate([pairs.rows, pairs.cols]); qry = np.concatenate([pairs.cols, pairs.rows])
    st = api.align_pairs(ctx, g, ref, qry)
    res = api.align_result_from_pairs(g, ref, qry, st)
    api.write_ani(g, res, tmp_path / "a.tsv")
    api.align([p], tmp_path / "b.tsv", True, filter_file=golden / "example" / "fltr.txt")
    assert (tmp_path / "a.tsv").read_bytes() == (tmp_path / "b.tsv").read_bytes()
    assert (tmp_path / "a.ids.tsv").read_bytes() == (tmp_path / "b.ids.tsv").read_bytes()


# ---------------------------------------------------------------- full-size configuration (BASELINE configs[1] = c2)
def test_c2_full_size_vs_reference_binaries(ctx, tmp_path):
    """1000 x 40 kb genomes: filter file and ani.tsv must be byte-identical to the unmodified reference tools
    (oracle/_ref, run here on the host cores).  Skipped when the binaries did not travel."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref binaries not present")
    names, seqs = synth.make_genomes(**synth.CONFIGS["c2"])
    fa = tmp_path / "c2.fna"
    synth.write_fasta(fa, names, seqs)
    api.prefilter([fa], tmp_path / "fltr.txt", True)
    oracle.ref_prefilter([fa], tmp_path / "ref_fltr.txt", tmp_path / "p")
    assert (tmp_path / "fltr.txt").read_bytes() == (tmp_path / "ref_fltr.txt").read_bytes()
    api.align([fa], tmp_path / "ani.tsv", True, filter_file=tmp_path / "fltr.txt", out_format=api.ALIGN_OUTFMT["complete"])
    oracle.ref_align([fa], tmp_path / "ref_ani.tsv", tmp_path / "a", filter_path=tmp_path / "ref_fltr.txt",
                     columns=api.ALIGN_OUTFMT["complete"])
    assert (tmp_path / "ani.tsv").read_bytes() == (tmp_path / "ref_ani.tsv").read_bytes()
    assert (tmp_path / "ani.ids.tsv").read_bytes() == (tmp_path / "ref_ani.ids.tsv").read_bytes()


def test_c2_unrelated_pairs_sample_vs_oracle(ctx):
    """All-vs-all slice of c2 (mostly unrelated pairs: the lost-mode / spurious-anchor regime) against the C oracle."""
    names, seqs = synth.make_genomes(**synth.CONFIGS["c2"])
    raw = [s.tobytes() for s in seqs]
    rng = np.random.default_rng(11)
    ref = rng.integers(0, len(raw), size=600)
    qry = rng.integers(0, len(raw), size=600)
    g = api.Genomes.from_memory(names, raw)
    got = api.align_pairs(ctx, g, ref, qry)
    sub = sorted(set(ref.tolist()) | set(qry.tolist()))
    remap = {x: i for i, x in enumerate(sub)}
    want = oracle.run_pairs([oracle.lz_codes(raw[x]) for x in sub], [remap[x] for x in ref], [remap[x] for x in qry])
    assert np.array_equal(got, want)


# ---------------------------------------------------------------- the command line (vclust.py's two sub-commands)
def test_cli_reproduces_the_reference_example(golden, tmp_path, capfd):
    """`vclust prefilter` + `vclust align --filter ... --out-aln ...` through vclust_b200.cli, multi-FASTA and directory
    input, against example/output/* (the reference's own test.py:300-530 checks the same files); -v 0 prints nothing."""
    from vclust_b200 import cli
    fa = golden / "example" / "multifasta.fna.gz"
    flt, ani, aln = tmp_path / "fltr.txt", tmp_path / "ani.tsv", tmp_path / "ani.aln.tsv"
    cli.main(["prefilter", "-i", str(fa), "-o", str(flt), "--min-ident", "0.7", "-v", "0", "--batch-size", "4"])
    assert flt.read_bytes() == (golden / "example" / "fltr.txt").read_bytes()
    cli.main(["align", "-i", str(fa), "-o", str(ani), "--out-aln", str(aln), "-v", "0"])
    assert ani.read_bytes() == (golden / "example" / "ani.tsv").read_bytes()
    assert (tmp_path / "ani.ids.tsv").read_bytes() == (golden / "example" / "ani.ids.tsv").read_bytes()
    want_aln = gzip.open(golden / "example" / "ani.aln.tsv.gz", "rt").read().splitlines()
    got_aln = aln.read_text().splitlines()
    assert got_aln[0] == want_aln[0] and sorted(got_aln[1:]) == sorted(want_aln[1:])
    out, err = capfd.readouterr()
    assert out == "" and err == ""
    # filtered align, lite format, output filter: the rows of the golden that pass --out-ani 0.95
    ani2 = tmp_path / "ani2.tsv"
    cli.main(["align", "-i", str(fa), "-o", str(ani2), "--filter", str(flt), "--outfmt", "lite", "--out-ani", "0.95", "-v", "0"])
    rows = [ln.split("\t") for ln in ani2.read_text().splitlines()]
    assert rows[0] == api.ALIGN_OUTFMT["lite"]
    assert len(rows) > 1 and all(float(r[4]) >= 0.95 for r in rows[1:])
    # directory mode: one genome per file, named by the file name (vclust.py:685-702)
    d = tmp_path / "fna"
    d.mkdir()
    recs = oracle.read_records_kmerdb(fa)
    for name, seq in recs[:4]:
        (d / (name + ".fna")).write_bytes(b">" + name.encode() + b"\n" + seq + b"\n")
    flt_d = tmp_path / "fltr_dir.txt"
    cli.main(["prefilter", "-i", str(d), "-o", str(flt_d), "-v", "0"])
    want = oracle.prefilter_text_from_fasta(sorted(d.iterdir()), False)
    assert flt_d.read_text() == want
    with pytest.raises(SystemExit):
        cli.main(["prefilter", "-i", str(d), "-o", str(flt_d), "--batch-size", "2"])     # vclust.py:731-736


# ---------------------------------------------------------------- randomised differential runs (tests/fuzz_check.py)
def test_fuzz_rounds_vs_oracle(ctx):
    """A few rounds of the randomised run (random LZ / prefilter parameters and genome shapes); `python tests/fuzz_check.py
    120 1` is the long version (profiles/r01_fuzz.txt)."""
    import importlib.util
    from pathlib import Path
    spec = importlib.util.spec_from_file_location("fuzz_check", Path(__file__).resolve().parent / "fuzz_check.py")
    fz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fz)
    rng = np.random.default_rng(20261017)
    assert fz.fuzz_align(ctx, 6, rng) > 0
    assert fz.fuzz_prefilter(ctx, 8, rng) > 0


def test_prefilter_passes_same_filter_file(ctx, golden, tmp_path, monkeypatch):
    """Several passes over k-mer hash shards (inputs beyond 10^9 k-mers; --batch-size in spirit): byte-identical filter,
    thresholds and --max-seqs applied after the partial counts are summed."""
    monkeypatch.setenv("VB_PREFILTER_PASSES", "4")
    out = tmp_path / "fltr.txt"
    api.prefilter([golden / "example" / "multifasta.fna.gz"], out, True)
    assert out.read_bytes() == (golden / "example" / "fltr.txt").read_bytes()
    names, seqs = synth.make_genomes(**GEN["s60"])
    fa = tmp_path / "s60.fna"
    synth.write_fasta(fa, names, seqs)
    api.prefilter([fa], out, True, max_seqs=3)
    assert out.read_bytes() == (golden / "ref_synth" / "s60_ms3.fltr.txt").read_bytes()
    api.prefilter([fa], out, True, kmers_fraction=0.2, min_kmers=4)
    assert out.read_bytes() == (golden / "ref_synth" / "s60_f02.fltr.txt").read_bytes()


# ---------------------------------------------------------------- bubbles, scale, the device-built pair list, sharding
def _bubble_set():
    """8 500 short genomes that all carry one 140-base core (its k-mers are shared by > 8 000 genomes: kmer-db's bubble
    case, bubble_helper.h:79-151), 5 000 of them a second core, plus a 400-base body that is shared inside families of 5."""
    rng = np.random.default_rng(99)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    core1 = acgt[rng.integers(0, 4, size=140)]
    core2 = acgt[rng.integers(0, 4, size=100)]
    seqs, root = [], None
    for i in range(8500):
        if i % 5 == 0:
            root = acgt[rng.integers(0, 4, size=400)]
        body = root.copy()
        pos = rng.integers(0, 400, size=6)
        body[pos] = acgt[rng.integers(0, 4, size=6)]
        parts = [core1, np.frombuffer(b"N", dtype=np.uint8), body]
        if i % 17 < 10:
            parts += [np.frombuffer(b"N", dtype=np.uint8), core2]
        seqs.append(np.concatenate(parts))
    return ["b%05d" % i for i in range(len(seqs))], seqs


@pytest.mark.parametrize("env,min_kmers,min_ident", [
    ({}, 230, 0.5),                                                     # dense counters: bubbles expanded at once
    ({"VB_PREFILTER_HASH": "1"}, 230, 0.5),                             # hashed table: bubbles deferred, added when the thresholds are applied
    ({"VB_PREFILTER_HASH": "1", "VB_PREFILTER_PASSES": "3"}, 100, 0.97),   # ... and pairs of "heavy" genomes (bubble weight >= min-kmers) expanded
    ({"VB_PREFILTER_NO_COLLAPSE": "1", "VB_PREFILTER_PASSES": "2"}, 230, 0.5),
])
def test_prefilter_bubbles_vs_kmerdb(ctx, tmp_path, monkeypatch, env, min_kmers, min_ident):
    """k-mers shared by more than 8 000 genomes against the unmodified kmer-db (which defers them as "bubbles"): the
    filter file must be byte-identical.  --min-kmers above the size of the cores (or a high --min-ident) keeps the file
    small while every kept pair still needs the bubble counts to be right."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref binaries not present")
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    names, seqs = _bubble_set()
    fa = tmp_path / "bubbles.fna"
    synth.write_fasta(fa, names, seqs)
    api.prefilter([fa], tmp_path / "fltr.txt", True, kmer_size=21, min_kmers=min_kmers, min_ident=min_ident)
    oracle.ref_prefilter([fa], tmp_path / "ref.txt", tmp_path / "p", k=21, min_kmers=min_kmers, min_ident=min_ident)
    got = (tmp_path / "fltr.txt").read_bytes()
    assert got == (tmp_path / "ref.txt").read_bytes()
    assert got.count(b":") > 8000            # families of 5 -> ~17 000 pairs


def test_align_device_built_list_equals_host_built_list(ctx, golden, tmp_path, monkeypatch):
    """vb_align builds the directed pair list, its order and the schedule on the GPU (from the candidate list the prefilter
    left on the device, or from an upload); VB_ALIGN_HOST_LIST=1 selects the host-built list: same result object."""
    names, seqs = synth.make_genomes(**GEN["s60"])
    g = api.Genomes.from_memory(names, seqs)
    runs = []
    for mode in ("device-cached", "device-upload", "host"):
        pairs = api.prefilter_genomes(ctx, g)
        if mode == "device-upload":
            ctx.evict()                       # drops the packed genomes, not the pair list; a fresh list object has no device copy:
            flt = tmp_path / "f.txt"
            api.write_filter(g, pairs, flt)
            pairs.close()
            pairs = api.read_filter(flt, 0.0, g)
        if mode == "host":
            monkeypatch.setenv("VB_ALIGN_HOST_LIST", "1")
        res = api.align_genomes(ctx, g, pairs)
        runs.append((res.ref.tolist(), res.qry.tolist(), res.stats.tolist(), res.order.tolist()))
        pairs.close(); res.close()
    assert runs[0] == runs[1] == runs[2]
    assert len(runs[0][0]) > 100
    # all-vs-all: generated on the device
    monkeypatch.delenv("VB_ALIGN_HOST_LIST")
    sub = api.Genomes.from_memory(names[:9], seqs[:9])
    a = api.align_genomes(ctx, sub)
    monkeypatch.setenv("VB_ALIGN_HOST_LIST", "1")
    b = api.align_genomes(ctx, sub)
    assert (a.ref.tolist(), a.qry.tolist(), a.stats.tolist()) == (b.ref.tolist(), b.qry.tolist(), b.stats.tolist())
    assert a.n == 72


def test_anchor_table_of_a_repeat_genome(ctx):
    """A genome that is one long tandem repeat puts thousands of entries on a handful of probe chains of the anchor
    table (entries are carried from one shared-memory partition of the build into the next): same statistics as the
    oracle."""
    rng = np.random.default_rng(3)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    unit = acgt[rng.integers(0, 4, size=13)]
    rep = np.tile(unit, 4000)                                  # 52 kb, 13 distinct 11-mers per strand
    polya = np.full(40000, ord("A"), dtype=np.uint8)
    mixed = np.concatenate([acgt[rng.integers(0, 4, size=20000)], np.tile(unit, 1500), acgt[rng.integers(0, 4, size=5000)]])
    mut = mixed.copy()
    mut[rng.integers(0, mut.size, size=900)] = acgt[rng.integers(0, 4, size=900)]
    raw = [x.tobytes() for x in (rep, polya, mixed, mut)]
    g = api.Genomes.from_memory(["rep", "polya", "mixed", "mut"], raw)
    ref = [0, 2, 3, 2, 1, 0, 3]
    qry = [2, 0, 2, 3, 3, 3, 1]
    got = api.align_pairs(ctx, g, ref, qry)
    want = oracle.run_pairs([oracle.lz_codes(s) for s in raw], ref, qry)
    assert np.array_equal(got, want)


def _shard_worker(rank, world, port, out_dir, cfg, backend):
    import os
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    import torch
    import torch.distributed as dist

    from vclust_b200 import distributed, synth as sy
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(0)
    dist.init_process_group(backend, rank=rank, world_size=world)
    names, seqs = sy.make_genomes(**cfg)
    lengths = [int(s.size) for s in seqs]
    first, count = distributed.block_partition(lengths, world)[rank]
    run = distributed.ShardedRun(dist, 0, names, lengths, seqs[first:first + count])
    pairs = run.prefilter()
    res = run.align()
    if rank == 0:
        np.savez(Path(out_dir) / "shard.npz", prow=pairs.rows, pcol=pairs.cols, pcommon=pairs.common, pani=pairs.ani,
                 totals=pairs.total_kmers, ref=res.ref, qry=res.qry, stats=res.stats, order=res.order)
    else:
        assert pairs.n_pairs == 0 and res.n == 0
    pairs.close(); res.close()
    run.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [1, 2, 3])
def test_sharded_pipeline_equals_single_gpu(ctx, tmp_path, world):
    """The multi-GPU pipeline (block-partitioned genomes, tuple all-to-all, owner merge, owner-local parses, gather) with
    `world` processes that share this GPU -- the collectives are torch.distributed callbacks (gloo, staged through the
    host here; NCCL on a multi-GPU box, tests/mgpu_nccl_check.py) -- must return exactly what one GPU returns."""
    import socket

    import torch.multiprocessing as mp
    cfg = dict(n=90, length=(3000, 30000), family=6, seed=synth.BASE_SEED + 300, n_frac=0.1, lower_frac=0.1)
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mp.spawn(_shard_worker, args=(world, port, str(tmp_path), cfg, "gloo"), nprocs=world, join=True)
    got = np.load(tmp_path / "shard.npz")
    names, seqs = synth.make_genomes(**cfg)
    g = api.Genomes.from_memory(names, seqs)
    pairs = api.prefilter_genomes(ctx, g)
    res = api.align_genomes(ctx, g, pairs)
    assert got["totals"].tolist() == pairs.total_kmers.tolist()
    assert (got["prow"].tolist(), got["pcol"].tolist(), got["pcommon"].tolist()) == (pairs.rows.tolist(), pairs.cols.tolist(), pairs.common.tolist())
    assert np.array_equal(got["pani"], pairs.ani)
    assert (got["ref"].tolist(), got["qry"].tolist(), got["stats"].tolist(), got["order"].tolist()) == \
           (res.ref.tolist(), res.qry.tolist(), res.stats.tolist(), res.order.tolist())
    assert pairs.n_pairs > 150


def _ref_vs_files(tmp_path, fa, ours_filter, ours_ani):
    oracle.ref_prefilter([fa], tmp_path / "ref_fltr.txt", tmp_path / "p")
    assert ours_filter.read_bytes() == (tmp_path / "ref_fltr.txt").read_bytes()
    oracle.ref_align([fa], tmp_path / "ref_ani.tsv", tmp_path / "a", filter_path=tmp_path / "ref_fltr.txt",
                     columns=api.ALIGN_OUTFMT["complete"])
    assert ours_ani.read_bytes() == (tmp_path / "ref_ani.tsv").read_bytes()
    ids = ours_ani.with_name(ours_ani.stem + ".ids" + ours_ani.suffix)
    assert ids.read_bytes() == (tmp_path / "ref_ani.ids.tsv").read_bytes()


def test_c3_full_size_vs_reference_binaries(tmp_path):
    """BASELINE configs[2]: 10 000 x 40 kb genomes, 95 000 candidate pairs -- filter, ani.tsv and ids.tsv byte-identical to
    the unmodified reference tools."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref binaries not present")
    names, seqs = synth.make_genomes(**synth.CONFIGS["c3"])
    fa = tmp_path / "c3.fna"
    synth.write_fasta(fa, names, seqs)
    del seqs
    api.prefilter([fa], tmp_path / "fltr.txt", True)
    api.align([fa], tmp_path / "ani.tsv", True, filter_file=tmp_path / "fltr.txt", out_format=api.ALIGN_OUTFMT["complete"])
    assert (tmp_path / "fltr.txt").read_bytes().count(b":") == 95000 + 2
    _ref_vs_files(tmp_path, fa, tmp_path / "fltr.txt", tmp_path / "ani.tsv")


def test_c3_s200_slice_vs_reference_binaries(tmp_path):
    """The sum-of-m^2 regime (families of 200): the first 2 000 genomes of c3_s200, 199 000 candidate pairs."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref binaries not present")
    cfg = dict(synth.CONFIGS["c3_s200"], n=2000)
    names, seqs = synth.make_genomes(**cfg)
    fa = tmp_path / "s200.fna"
    synth.write_fasta(fa, names, seqs)
    api.prefilter([fa], tmp_path / "fltr.txt", True)
    api.align([fa], tmp_path / "ani.tsv", True, filter_file=tmp_path / "fltr.txt", out_format=api.ALIGN_OUTFMT["complete"])
    assert (tmp_path / "fltr.txt").read_bytes().count(b":") == 199000 + 2
    _ref_vs_files(tmp_path, fa, tmp_path / "fltr.txt", tmp_path / "ani.tsv")


def _shard_files_worker(rank, world, port, out_dir, cfg):
    import os
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    import torch
    import torch.distributed as dist

    from vclust_b200 import api as ap, distributed, synth as sy
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    names, seqs = sy.make_genomes(**cfg)
    lengths = [int(s.size) for s in seqs]
    first, count = distributed.block_partition(lengths, world)[rank]
    run = distributed.ShardedRun(dist, 0, names, lengths, seqs[first:first + count])
    pairs = run.prefilter()
    res = run.align()
    if rank == 0:
        ap.write_filter(run.meta, pairs, Path(out_dir) / "fltr.txt")
        ap.write_ani(run.meta, res, Path(out_dir) / "ani.tsv", None, ap.ALIGN_OUTFMT["complete"])
    pairs.close(); res.close()
    run.close()
    dist.destroy_process_group()


def test_c4_shaped_sharded_vs_reference_binaries(tmp_path):
    """c4's shape (5-200 kb, N runs, lower case; families of 100) on 5 000 genomes through the SHARDED pipeline (two ranks
    sharing this GPU) against the unmodified reference tools -- not against the single-GPU path."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref binaries not present")
    import socket

    import torch.multiprocessing as mp
    cfg = dict(n=5000, length=(5000, 200000), family=100, seed=synth.BASE_SEED + 4, n_frac=0.01, lower_frac=0.01)
    names, seqs = synth.make_genomes(**cfg)
    fa = tmp_path / "c4s.fna"
    synth.write_fasta(fa, names, seqs)
    del seqs
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mp.spawn(_shard_files_worker, args=(2, port, str(tmp_path), cfg), nprocs=2, join=True)
    _ref_vs_files(tmp_path, fa, tmp_path / "fltr.txt", tmp_path / "ani.tsv")
