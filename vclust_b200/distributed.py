"""One genome set over the GPUs of one node: one process per GPU, ``torch.distributed`` for the plumbing (NCCL on GPUs,
gloo in the CPU tests), ``libvclust_b200.so`` for everything else.

Data path (the library side is csrc/shard.cu + csrc/prefilter.cu + csrc/align.cu):

  setup     the genomes are block-partitioned by cumulative length; every rank packs its block on its GPU and the packed
            align-stage records are ALL-GATHERED, so every GPU can read every genome afterwards
  extract   k-mers of the LOCAL genomes -> (hash, genome) tuples, partitioned by hash range
  exchange  all-to-all #1: every tuple travels to the rank that owns its hash range
  count     grouping + pair counting on the owner's range -> partial common-k-mer counts per genome pair
  reduce    all-reduce of total-kmers; all-to-all #2: partial counts travel to the owners of both genomes
            (owner(g) = g mod world), which sum them and apply the -min filters: each owner then holds the complete
            candidate list of its genomes
  align     every owner parses (reference = own genome, query) for its pairs; results are gathered on rank 0

The library owns all device buffers and runs all kernels on ONE stream that this module creates; the collectives are
handed to it as C callbacks (``TorchComm``) and are issued on that same stream, so kernels and collectives are ordered by
the stream and the host never waits in between except where it needs a count.

``exchange_and_align`` / ``prefilter_align_sharded`` below are the older host-staged variant (k-mer shards of replicated
genomes, numpy exchange); it is kept for ``--max-seqs``, which the native pipeline does not cover.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional, Sequence, Tuple

import numpy as np

from . import _lib


def owner(g, world: int):
    return g % world


def block_partition(lengths: Sequence[int], world: int):
    """Contiguous blocks of genomes with (nearly) equal numbers of bases: [(first, count)] per rank."""
    lengths = np.asarray(lengths, dtype=np.int64)
    cum = np.concatenate([[0], np.cumsum(lengths)])
    total = int(cum[-1])
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(cum, total * r / world, side="left")))
    cuts.append(len(lengths))
    cuts = np.maximum.accumulate(np.minimum(cuts, len(lengths)))
    return [(int(cuts[r]), int(cuts[r + 1] - cuts[r])) for r in range(world)]


def _wrap(ptr: int, nbytes: int, device):
    """A uint8 torch tensor over `nbytes` bytes of memory the library owns (device memory on a GPU, host memory in the
    gloo tests)."""
    import torch
    if nbytes == 0 or not ptr:
        return torch.empty(0, dtype=torch.uint8, device=device)
    if device.type == "cuda":
        class _Mem:
            pass
        m = _Mem()
        m.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 3}
        return torch.as_tensor(m, device=device)
    return torch.frombuffer((C.c_uint8 * int(nbytes)).from_address(int(ptr)), dtype=torch.uint8)


class TorchComm:
    """vb_comm over torch.distributed.  All three callbacks work on memory owned by the caller (the library) and are
    enqueued on the current torch stream; they return 0 or, after stashing the exception in ``self.error``, 1."""

    def __init__(self, dist, device, group=None):
        import torch
        self.dist, self.device, self.group = dist, torch.device(device), group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        # device memory but a CPU-only backend (several ranks sharing one GPU in the tests): stage through the host
        self.staged = self.device.type == "cuda" and dist.get_backend(group) == "gloo"
        self.error: Optional[BaseException] = None
        self.bytes_sent = 0          # payload this rank handed to all-to-all / all-gather (diagnostics)
        self.calls = 0
        self._a2a = _lib.A2A_FN(self._all_to_all)
        self._gather = _lib.GATHER_FN(self._all_gather)
        self._reduce = _lib.REDUCE_FN(self._all_reduce)
        self.struct = _lib.Comm(self.rank, self.world, None, self._a2a, self._gather, self._reduce)

    def _guard(self, fn):
        try:
            fn()
            return 0
        except BaseException as e:      # noqa: BLE001 -- nothing may propagate through the C frames
            self.error = e
            return 1

    def _all_to_all(self, user, send, send_counts, recv, recv_counts, width):
        def run():
            w = self.world
            sc = [int(send_counts[i]) * int(width) for i in range(w)]
            rc = [int(recv_counts[i]) * int(width) for i in range(w)]
            s = _wrap(send, sum(sc), self.device)
            r = _wrap(recv, sum(rc), self.device)
            if self.staged:
                import torch
                self.stream_sync()
                hr = torch.empty(sum(rc), dtype=torch.uint8)
                self.dist.all_to_all_single(hr, s.cpu(), output_split_sizes=rc, input_split_sizes=sc, group=self.group)
                r.copy_(hr)
            else:
                self.dist.all_to_all_single(r, s, output_split_sizes=rc, input_split_sizes=sc, group=self.group)
            self.bytes_sent += sum(sc) - sc[self.rank]
            self.calls += 1
        return self._guard(run)

    def _all_gather(self, user, send, recv, nbytes):
        def run():
            s = _wrap(send, int(nbytes), self.device)
            r = _wrap(recv, int(nbytes) * self.world, self.device)
            if self.staged:
                import torch
                self.stream_sync()
                hr = torch.empty(int(nbytes) * self.world, dtype=torch.uint8)
                self.dist.all_gather(list(hr.chunk(self.world)), s.cpu(), group=self.group)
                r.copy_(hr)
            elif self.device.type == "cuda":
                self.dist.all_gather_into_tensor(r, s, group=self.group)
            else:
                self.dist.all_gather(list(r.chunk(self.world)), s, group=self.group)
            self.bytes_sent += int(nbytes) * (self.world - 1)
            self.calls += 1
        return self._guard(run)

    def _all_reduce(self, user, buf, n):
        def run():
            import torch
            t = _wrap(buf, 4 * int(n), self.device).view(torch.int32)     # two's complement: the sum is exact mod 2^32
            if self.staged:
                self.stream_sync()
                h = t.cpu()
                self.dist.all_reduce(h, op=self.dist.ReduceOp.SUM, group=self.group)
                t.copy_(h)
            else:
                self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
            self.calls += 1
        return self._guard(run)

    def stream_sync(self):
        import torch
        torch.cuda.current_stream(self.device).synchronize()

    def check(self, rc: int):
        if rc != 0:
            err, self.error = self.error, None
            if err is not None:
                raise err
            _lib.check(rc)


class ShardedRun:
    """The multi-GPU pipeline on one rank.

    names / lengths: ALL genomes of the set (every rank knows them); local_seqs: the sequences of this rank's block
    (ASCII, bytes or uint8 arrays).  blocks: [(first, count)] per rank, contiguous and in rank order; default
    ``block_partition(lengths, world)`` (equal numbers of bases)."""

    def __init__(self, dist, device_index: int, names: Sequence[str], lengths: Sequence[int], local_seqs: Sequence, mrd: int = 40,
                 group=None, blocks=None):
        import torch

        from . import api
        self.dist = dist
        self.device = torch.device("cuda", device_index)
        self.stream = torch.cuda.Stream(device=self.device)
        self.comm = TorchComm(dist, self.device, group)
        self.rank, self.world = self.comm.rank, self.comm.world
        self.first, count = (blocks or block_partition(lengths, self.world))[self.rank]
        if len(local_seqs) != count:
            raise ValueError("rank %d holds %d genomes, its block has %d" % (self.rank, len(local_seqs), count))
        self.mrd = int(mrd)
        self.ctx = api.Context(device_index, stream=self.stream.cuda_stream)
        self.meta = api.Genomes.skeleton(names, lengths)
        self.local = api.Genomes.from_memory(list(names[self.first:self.first + count]), local_seqs)
        self._L = _lib.load()
        self._h = C.c_void_p()
        self.load()

    def load(self):
        """Collective: (re-)upload and pack this rank's block, all-gather the packed records (the host-buffer leg of the
        end-to-end measurement; the constructor calls it once)."""
        import torch
        self.unload()
        with torch.cuda.stream(self.stream):
            self.comm.check(self._L.vb_shard_create(self.ctx._h, C.byref(self.comm.struct), self.meta._h, self.local._h,
                                                    self.first, self.mrd, C.byref(self._h)))

    def unload(self):
        if self._h:
            self._L.vb_shard_destroy(self._h)
            self._h = C.c_void_p()

    def prefilter(self, k=25, min_kmers=20, min_ident=0.7, kmers_fraction=1.0):
        """Collective.  Rank 0 gets the complete candidate list (api.PairList), the others an empty one."""
        import torch

        from . import api
        p = _lib.PrefilterParams(k, min_kmers, min_ident, kmers_fraction, 0, 0)
        out = C.POINTER(_lib.Pairs)()
        with torch.cuda.stream(self.stream):
            self.comm.check(self._L.vb_shard_prefilter(self._h, C.byref(p), C.byref(out)))
        return api.PairList(out)

    def align(self, params=None):
        """Collective, after prefilter().  Rank 0 gets the complete api.AlignResult, the others an empty one."""
        import torch

        from . import api
        params = params or api.align_params()
        out = C.POINTER(_lib.AlignOut)()
        with torch.cuda.stream(self.stream):
            self.comm.check(self._L.vb_shard_align(self._h, C.byref(params), C.byref(out)))
        return api.AlignResult(out)

    def close(self):
        self.unload()
        self.local.close()
        self.meta.close()
        self.ctx.close()


# ----------------------------------------------------------------------------------------------------------------
# host-staged variant (kept for --max-seqs)
# ----------------------------------------------------------------------------------------------------------------
def _a2a_variable(dist, send_chunks, device, width: int):
    """all-to-all of int64 rows with per-destination sizes; returns the concatenation of what was received."""
    import torch
    world = dist.get_world_size()
    counts = torch.tensor([c.shape[0] for c in send_chunks], dtype=torch.int64, device=device)
    recv_counts = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_to_all_single(recv_counts, counts)
    rc = recv_counts.cpu().tolist()
    send = torch.from_numpy(np.concatenate(send_chunks, axis=0).astype(np.int64).reshape(-1, width)).to(device)
    recv = torch.empty((int(sum(rc)), width), dtype=torch.int64, device=device)
    dist.all_to_all_single(recv, send, output_split_sizes=[int(x) for x in rc],
                           input_split_sizes=[int(c.shape[0]) for c in send_chunks])
    return recv.cpu().numpy()


def _gather_rows(dist, rows: np.ndarray, device, width: int, dst: int = 0):
    """variable-size gather of int64 rows on rank dst (padded all_gather: works on NCCL and gloo alike)."""
    import torch
    world = dist.get_world_size()
    n = torch.tensor([rows.shape[0]], dtype=torch.int64, device=device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    m = max(max(sizes), 1)
    buf = torch.zeros((m, width), dtype=torch.int64, device=device)
    if rows.shape[0]:
        buf[:rows.shape[0]] = torch.from_numpy(rows.astype(np.int64).reshape(-1, width)).to(device)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    if dist.get_rank() != dst:
        return None
    return np.concatenate([o[:s].cpu().numpy() for o, s in zip(out, sizes)], axis=0)


def exchange_and_align(dist, device, partial: Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray],
                       merge_fn: Callable, align_fn: Callable, sampled: bool = False):
    """Host-staged exchange.  partial = (rows, cols, common, partial_totals) of this rank's k-mer shard.
    merge_fn(rows, cols, common, totals) -> (rows, cols, common, ani) kept pairs (thresholds applied, sorted).
    align_fn(ref, qry) -> (n, 3) int32.
    sampled: merge_fn applies --max-seqs, i.e. returns ROWS of the filter (entries on both sides of the diagonal).  An
    owner then only trusts the rows of its own genomes (it holds every pair they occur in, so their top-N is exact) and
    a second, small all-to-all sends each kept entry (row, item) to the owner of `item`: LZ-ANI symmetrises every filter
    entry (L/filter.cpp:80-81), so genome x is parsed as the reference of y once per entry (x, y) and once per (y, x).
    Returns on rank 0: dict(totals, pairs=(row, col, common, ani), ref, qry, stats); on other ranks None."""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    rows, cols, common, tot = partial
    totals = torch.from_numpy(np.asarray(tot, dtype=np.int64)).to(device)
    dist.all_reduce(totals, op=dist.ReduceOp.SUM)
    totals = totals.cpu().numpy().astype(np.uint32)

    trip = np.stack([np.asarray(rows, np.int64), np.asarray(cols, np.int64), np.asarray(common, np.int64)], axis=1) \
        if len(rows) else np.zeros((0, 3), np.int64)
    o_r, o_c = owner(trip[:, 0], world), owner(trip[:, 1], world)
    chunks = []
    for d in range(world):
        sel = (o_r == d) | (o_c == d)
        chunks.append(trip[sel])
    got = _a2a_variable(dist, chunks, device, 3)

    m_rows, m_cols, m_common, m_ani = merge_fn(got[:, 0].astype(np.uint32), got[:, 1].astype(np.uint32),
                                               got[:, 2].astype(np.uint32), totals)
    m_rows, m_cols = np.asarray(m_rows, np.int64), np.asarray(m_cols, np.int64)
    mine_r = owner(m_rows, world) == rank
    if sampled:
        # rows of foreign genomes were sampled from an incomplete candidate list: drop them, their owners have them
        m_rows, m_cols = m_rows[mine_r], m_cols[mine_r]
        m_common, m_ani = np.asarray(m_common)[mine_r], np.asarray(m_ani)[mine_r]
        mine_r = np.ones(m_rows.size, dtype=bool)
        ent = np.stack([m_rows, m_cols], axis=1) if m_rows.size else np.zeros((0, 2), np.int64)
        o_item = owner(ent[:, 1], world)
        back = _a2a_variable(dist, [ent[o_item == d] for d in range(world)], device, 2)    # entries (row, item), item is mine
        ref = np.concatenate([m_rows, back[:, 1]])
        qry = np.concatenate([m_cols, back[:, 0]])
    else:
        mine_c = owner(m_cols, world) == rank
        ref = np.concatenate([m_rows[mine_r], m_cols[mine_c]])
        qry = np.concatenate([m_cols[mine_r], m_rows[mine_c]])
    stats = np.asarray(align_fn(ref.astype(np.uint32), qry.astype(np.uint32)), dtype=np.int64).reshape(-1, 3)

    res_rows = np.concatenate([ref[:, None], qry[:, None], stats], axis=1) if ref.size else np.zeros((0, 5), np.int64)
    all_res = _gather_rows(dist, res_rows, device, 5)
    # candidate pairs: reported once, by the owner of `row`; ani travels as its IEEE bit pattern
    ani_bits = np.asarray(m_ani, np.float64).view(np.int64)
    pr = np.stack([m_rows, m_cols, np.asarray(m_common, np.int64), ani_bits], axis=1)[mine_r] if m_rows.size \
        else np.zeros((0, 4), np.int64)
    all_pairs = _gather_rows(dist, pr, device, 4)
    if rank != 0:
        return None
    order = np.lexsort((all_pairs[:, 1], all_pairs[:, 0]))
    all_pairs = all_pairs[order]
    return dict(totals=totals,
                pairs=(all_pairs[:, 0].astype(np.uint32), all_pairs[:, 1].astype(np.uint32),
                       all_pairs[:, 2].astype(np.uint32), all_pairs[:, 3].copy().view(np.float64)),
                ref=all_res[:, 0].astype(np.uint32), qry=all_res[:, 1].astype(np.uint32),
                stats=all_res[:, 2:5].astype(np.int32))


def prefilter_align_sharded(ctx, genomes_kmerdb, genomes_lzani, dist, device, k=25, min_kmers=20, min_ident=0.7,
                            kmers_fraction=1.0, lz_params=None, max_seqs=0, passes=0) -> Optional[dict]:
    """Host-staged variant on GPUs: vb_prefilter_partial -> exchange -> vb_pairs_merge -> vb_align_pairs.  Every rank
    holds all genomes (the same input loaded with the two FASTA flavours; they may be the same object when the input has
    no corner cases) and counts the k-mers of hash shard `rank`."""
    from . import api
    rank, world = dist.get_rank(), dist.get_world_size()
    passes = max(1, passes)
    rows, cols, common, totals = [], [], [], None
    for s in range(passes):
        part = api.prefilter_partial(ctx, genomes_kmerdb, rank * passes + s, world * passes, k=k, kmers_fraction=kmers_fraction)
        rows.append(part.rows); cols.append(part.cols); common.append(part.common)
        totals = part.total_kmers.astype(np.int64) if totals is None else totals + part.total_kmers
        part.close()
    partial = (np.concatenate(rows), np.concatenate(cols), np.concatenate(common), totals)

    def merge_fn(r, c, v, totals):
        m = api.merge_pairs(r, c, v, totals, k=k, min_kmers=min_kmers, min_ident=min_ident, kmers_fraction=kmers_fraction,
                            max_seqs=max_seqs)
        out = (m.rows, m.cols, m.common, m.ani)
        m.close()
        return out

    def align_fn(ref, qry):
        return api.align_pairs(ctx, genomes_lzani, ref, qry, lz_params)

    return exchange_and_align(dist, device, partial, merge_fn, align_fn, sampled=max_seqs > 0)
