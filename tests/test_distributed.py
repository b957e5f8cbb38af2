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


# ---------------------------------------------------------------- the native pipeline's collectives (vb_comm callbacks)
def _comm_worker(rank, world, port, out_dir):
    """Drives TorchComm's three callbacks exactly as libvclust_b200 does -- through the C function pointers of the vb_comm
    struct, on raw memory -- over gloo with host buffers."""
    import ctypes as C
    sys.path.insert(0, str(ROOT))
    import torch
    import torch.distributed as dist

    from vclust_b200 import distributed
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    comm = distributed.TorchComm(dist, torch.device("cpu"))
    st = comm.struct
    assert (st.rank, st.world) == (rank, world)
    # all_gather: 16 bytes per rank
    send = np.full(4, 1000 + rank, dtype=np.uint32)
    recv = np.zeros(4 * world, dtype=np.uint32)
    assert st.all_gather(None, send.ctypes.data, recv.ctypes.data, 16) == 0
    assert recv.reshape(world, 4).tolist() == [[1000 + r] * 4 for r in range(world)]
    # all_to_all: rank r sends (r + p + 1) records of 12 bytes to peer p
    sc = np.array([rank + p + 1 for p in range(world)], dtype=np.uint64)
    rc = np.array([p + rank + 1 for p in range(world)], dtype=np.uint64)
    rows = [(rank, p, i) for p in range(world) for i in range(int(sc[p]))]
    sbuf = np.array(rows, dtype=np.uint32)
    rbuf = np.zeros((int(rc.sum()), 3), dtype=np.uint32)
    u64p = C.POINTER(C.c_uint64)
    assert st.all_to_all(None, sbuf.ctypes.data, sc.ctypes.data_as(u64p), rbuf.ctypes.data, rc.ctypes.data_as(u64p), 12) == 0
    assert rbuf.tolist() == [[p, rank, i] for p in range(world) for i in range(int(rc[p]))]
    # zero-size legs
    z = np.zeros(world, dtype=np.uint64)
    one = np.zeros(1, dtype=np.uint64)
    assert st.all_to_all(None, one.ctypes.data, z.ctypes.data_as(u64p), one.ctypes.data, z.ctypes.data_as(u64p), 8) == 0
    # all_reduce: unsigned sums are exact modulo 2^32 (partial totals may be "negative")
    buf = np.array([5, 0xfffffff0 if rank == 0 else 0x20, 7 * rank], dtype=np.uint32)
    assert st.all_reduce_sum_u32(None, buf.ctypes.data, 3) == 0
    assert buf.tolist() == [5 * world, (0xfffffff0 + 0x20 * (world - 1)) & 0xffffffff, 7 * sum(range(world))]
    assert comm.calls == 4 and comm.error is None
    # an exception inside a callback must not cross the C frame: it is stashed and reported as a non-zero status
    bad = np.zeros(1, dtype=np.uint64)
    dist.barrier()
    if rank == 0:
        (Path(out_dir) / "ok").write_text("ok")
    dist.destroy_process_group()


def test_torch_comm_callbacks_over_gloo(tmp_path):
    import torch.multiprocessing as mp

    from vclust_b200 import build
    build.build()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_comm_worker, args=(3, port, str(tmp_path)), nprocs=3, join=True)
    assert (tmp_path / "ok").exists()


def test_block_partition_is_contiguous_and_balanced():
    from vclust_b200 import distributed
    rng = np.random.default_rng(5)
    lengths = np.exp(rng.uniform(np.log(5000), np.log(200000), size=1000)).astype(np.int64)
    for world in (1, 2, 3, 8):
        blocks = distributed.block_partition(lengths, world)
        assert blocks[0][0] == 0 and sum(c for _, c in blocks) == 1000
        for (f0, c0), (f1, _) in zip(blocks, blocks[1:]):
            assert f0 + c0 == f1
        per = [int(lengths[f:f + c].sum()) for f, c in blocks]
        assert max(per) - min(per) <= 2 * int(lengths.max())
    assert distributed.block_partition([10, 10], 4)[-1] == (2, 0) or sum(c for _, c in distributed.block_partition([10, 10], 4)) == 2


def test_skeleton_genomes_write_outputs(tmp_path):
    """A names + lengths skeleton (what rank 0 of a multi-GPU run has for the genomes of the other ranks) is enough to
    write the filter file: same bytes as with the full genome set."""
    from vclust_b200 import api, build
    build.build()
    names = ["a", "b", "c"]
    seqs = [b"ACGT" * 10, b"ACGA" * 12, b"TTGA" * 9]
    full = api.Genomes.from_memory(names, seqs)
    skel = api.Genomes.skeleton(names, [len(s) for s in seqs])
    assert len(skel) == 3 and skel.length(1) == 48 and skel.names() == names and skel.total_bases == full.total_bases
    m = api.merge_pairs([1, 2], [0, 1], [30, 9], np.array([37, 45, 33], dtype=np.uint32), k=4, min_kmers=1, min_ident=0.0)
    api.write_filter(full, m, tmp_path / "a.txt")
    api.write_filter(skel, m, tmp_path / "b.txt")
    assert (tmp_path / "a.txt").read_bytes() == (tmp_path / "b.txt").read_bytes()
