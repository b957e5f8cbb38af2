"""World-size-2 gloo test (CPU) of the multi-GPU host logic in vclust_b200/distributed.py: k-mer-sharded partial
counts -> all-reduce of totals -> one all-to-all to the genome owners -> merge + thresholds -> owner-local directed
parses -> gather.  The compute steps are played by the CPU oracle (test infrastructure); the exchange, the merge
(vb_pairs_merge, host code of the C ABI) and the result assembly are the product code under test."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _fmix64(x):
    x = x.astype(np.uint64)
    with np.errstate(over="ignore"):
        x ^= x >> np.uint64(33); x *= np.uint64(0xff51afd7ed558ccd)
        x ^= x >> np.uint64(33); x *= np.uint64(0xc4ceb9fe1a85ec53)
        x ^= x >> np.uint64(33)
    return x


def _worker(rank, world, port, out_dir, max_seqs=0):
    sys.path.insert(0, str(ROOT))
    import torch.distributed as dist

    from oracle import oracle
    from vclust_b200 import api, distributed, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    names, seqs = synth.make_genomes(n=30, length=5000, family=5, seed=4242, n_frac=0.2)
    raw = [s.tobytes() for s in seqs]
    k, min_kmers, min_ident = 21, 10, 0.6
    sets = oracle.kmer_sets([[s] for s in raw], k, 1.0)
    shard = [s[(_fmix64(s) % np.uint64(world)) == np.uint64(rank)] for s in sets]
    rows, cols, vals = oracle.common_matrix(shard)
    partial = (rows, cols, vals, np.array([s.size for s in shard], dtype=np.uint32))
    codes = [oracle.lz_codes(s) for s in raw]

    def merge_fn(r, c, v, totals):
        m = api.merge_pairs(r, c, v, totals, k=k, min_kmers=min_kmers, min_ident=min_ident, max_seqs=max_seqs)
        return m.rows, m.cols, m.common, m.ani

    res = distributed.exchange_and_align(dist, "cpu", partial, merge_fn, lambda r, q: oracle.run_pairs(codes, r, q),
                                         sampled=max_seqs > 0)
    if rank == 0:
        np.savez(Path(out_dir) / "res.npz", totals=res["totals"], prow=res["pairs"][0], pcol=res["pairs"][1],
                 pcommon=res["pairs"][2], pani=res["pairs"][3], ref=res["ref"], qry=res["qry"], stats=res["stats"])
    dist.destroy_process_group()


def test_sharded_exchange_matches_single_process(tmp_path):
    import torch.multiprocessing as mp

    from oracle import oracle
    from vclust_b200 import build, synth
    build.build()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "res.npz")

    names, seqs = synth.make_genomes(n=30, length=5000, family=5, seed=4242, n_frac=0.2)
    raw = [s.tobytes() for s in seqs]
    sets = oracle.kmer_sets([[s] for s in raw], 21, 1.0)
    want = oracle.prefilter_pairs(sets, 21, 10, 0.6)
    assert got["totals"].tolist() == [int(s.size) for s in sets]
    assert list(zip(got["prow"].tolist(), got["pcol"].tolist(), got["pcommon"].tolist())) == [(r, c, v) for r, c, v, _ in want]
    assert np.array_equal(got["pani"], np.array([a for *_, a in want]))
    # every candidate pair parsed in both directions, exactly once, with the single-process statistics
    pairs = sorted(zip(got["ref"].tolist(), got["qry"].tolist()))
    assert pairs == sorted([(r, c) for r, c, *_ in want] + [(c, r) for r, c, *_ in want])
    st = oracle.run_pairs([oracle.lz_codes(s) for s in raw], got["ref"], got["qry"])
    assert np.array_equal(st, got["stats"])


def test_sharded_exchange_with_max_seqs(tmp_path):
    """--max-seqs over two ranks: every owner samples the rows of its own genomes, a second all-to-all returns the kept
    entries to the owners of their items; filter rows and the multiset of directed parses equal the single-process ones."""
    import torch.multiprocessing as mp

    from oracle import oracle
    from vclust_b200 import build, synth
    build.build()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path), 2), nprocs=2, join=True)
    got = np.load(tmp_path / "res.npz")
    names, seqs = synth.make_genomes(n=30, length=5000, family=5, seed=4242, n_frac=0.2)
    raw = [s.tobytes() for s in seqs]
    sets = oracle.kmer_sets([[s] for s in raw], 21, 1.0)
    want = oracle.prefilter_pairs(sets, 21, 10, 0.6, max_seqs=2)
    assert any(r < c for r, c, *_ in want) and len(want) < 2 * len(oracle.prefilter_pairs(sets, 21, 10, 0.6))
    assert list(zip(got["prow"].tolist(), got["pcol"].tolist(), got["pcommon"].tolist())) == [(r, c, v) for r, c, v, _ in want]
    assert np.array_equal(got["pani"], np.array([a for *_, a in want]))
    pairs = sorted(zip(got["ref"].tolist(), got["qry"].tolist()))
    assert pairs == sorted([(r, c) for r, c, *_ in want] + [(c, r) for r, c, *_ in want])       # duplicates included
    st = oracle.run_pairs([oracle.lz_codes(s) for s in raw], got["ref"], got["qry"])
    assert np.array_equal(st, got["stats"])


def test_merge_pairs_sums_duplicates_and_filters():
    from vclust_b200 import api, build
    build.build()
    totals = np.array([100, 80, 50], dtype=np.uint32)
    m = api.merge_pairs([1, 2, 1, 2, 2], [0, 0, 0, 1, 1], [30, 5, 40, 1, 2], totals, k=21, min_kmers=3, min_ident=0.0)
    assert list(zip(m.rows.tolist(), m.cols.tolist(), m.common.tolist())) == [(1, 0, 70), (2, 0, 5), (2, 1, 3)]
    m2 = api.merge_pairs([1, 2, 1, 2, 2], [0, 0, 0, 1, 1], [30, 5, 40, 1, 2], totals, k=21, min_kmers=3, min_ident=0.95)
    assert list(zip(m2.rows.tolist(), m2.cols.tolist())) == [(1, 0)]


def test_merge_pairs_max_seqs_matches_oracle_sampler():
    """--max-seqs (kmer-db -sample-rows ani-shorter:N) through the host code of the C ABI vs the oracle restatement,
    which tests/test_oracle_kmer.py pins against the reference binaries (golden s60_ms3)."""
    from oracle import oracle
    from vclust_b200 import api, build, synth
    build.build()
    names, seqs = synth.make_genomes(n=40, length=3000, family=8, seed=77)
    sets = oracle.kmer_sets([[s.tobytes()] for s in seqs], 21, 1.0)
    rows, cols, vals = oracle.common_matrix(sets)
    totals = np.array([s.size for s in sets], dtype=np.uint32)
    for ms in (1, 2, 5, 100):
        want = oracle.prefilter_pairs(sets, 21, 5, 0.5, max_seqs=ms)
        m = api.merge_pairs(rows, cols, vals, totals, k=21, min_kmers=5, min_ident=0.5, max_seqs=ms)
        assert list(zip(m.rows.tolist(), m.cols.tolist(), m.common.tolist())) == [(r, c, v) for r, c, v, _ in want]
        assert np.array_equal(m.ani, np.array([a for *_, a in want]))
        assert any(r < c for r, c, *_ in want)          # entries on both sides of the diagonal


def test_max_seqs_tie_break_is_smallest_item_first():
    """kmer-db's heap (K/sampler.h:45-66) evicts, among equal scores, the LARGEST item id: with all scores tied a row keeps
    its N smallest partners (confirmed with the reference binary on 7 identical genomes: rows `2:…,3:…` / `1:…,3:…` /
    `1:…,2:…`) -- checked on the host code of the C ABI and on the oracle restatement."""
    from oracle import oracle
    from vclust_b200 import api, build
    build.build()
    n = 7
    totals = np.full(n, 100, dtype=np.uint32)
    rows, cols = zip(*[(r, c) for r in range(n) for c in range(r)])
    m = api.merge_pairs(rows, cols, [50] * len(rows), totals, k=21, min_kmers=1, min_ident=0.0, max_seqs=2)
    got = {}
    for r, c in zip(m.rows.tolist(), m.cols.tolist()):
        got.setdefault(r, []).append(c)
    assert got == {r: sorted(x for x in range(n) if x != r)[:2] for r in range(n)}
    assert len(set(m.ani.tolist())) == 1
    # the same rule in the oracle (fed with k-mer sets that give identical counts: identical genomes)
    sets = [np.arange(100, dtype=np.uint64) for _ in range(n)]
    want = oracle.prefilter_pairs(sets, 21, 1, 0.0, max_seqs=2)
    assert [(r, c) for r, c, *_ in want] == list(zip(m.rows.tolist(), m.cols.tolist()))
