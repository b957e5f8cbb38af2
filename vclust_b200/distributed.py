"""One genome set over N GPUs of one node (one process per GPU, torch.distributed; NCCL on GPUs, gloo in CPU tests).

Plan, no tuple exchange and exactly one all-to-all on the data path:
  1. every rank holds all genomes; rank r counts only the k-mers of hash shard r (vb_prefilter_partial)
     -> partial (row, col, common) triples + partial total-kmers;
  2. total-kmers: all-reduce(SUM);
  3. ONE all-to-all: every partial triple goes to the owners of both of its genomes (owner(g) = g % N);
  4. every owner sums the partial counts, applies the two -min filters exactly (vb_pairs_merge) and now knows, for each
     of its genomes, the complete candidate list -> the directed parses (ref = own genome, query) are local and each
     reference index is built on exactly one GPU (vb_align_pairs);
  5. results (ref, qry, 3 ints) and the candidate pairs (reported by the owner of `row`) are gathered on rank 0.

The collectives live in ``exchange_and_align``; the compute steps are passed in as callables so that the host logic can
be tested on CPU with gloo (tests/test_distributed.py feeds it oracle-computed partial counts).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np


def owner(g, world: int):
    return g % world


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
    """Steps 2-5.  partial = (rows, cols, common, partial_totals) of this rank's k-mer shard.
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
    """The GPU instantiation: vb_prefilter_partial -> exchange -> vb_pairs_merge -> vb_align_pairs.
    genomes_kmerdb / genomes_lzani: the same input loaded with the two FASTA flavours (they may be the same object
    when the input has no corner cases, e.g. synthetic data)."""
    from . import api
    rank, world = dist.get_rank(), dist.get_world_size()
    # one pass holds ~10^9 k-mer tuples: a rank whose hash shard is larger splits it further and runs the sub-shards back to
    # back (c5: 3 x 10^10 k-mers over 8 GPUs = 4 passes per rank); their partial counts simply join the rank's list
    if passes <= 0:
        passes = max(1, int(np.ceil(genomes_kmerdb.total_bases * min(1.0, kmers_fraction) / world / 1.0e9)))
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
