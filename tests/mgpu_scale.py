"""The 8-GPU configurations of BASELINE.json through the multi-GPU pipeline (run under torchrun; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tests/mgpu_scale.py c4 [--n N] [--runs R]
    ... tests/mgpu_scale.py c5 --n 200000 --no-align

c4: 100 000 contigs of 5-200 kb in families of 200 (N runs, lower case), prefilter + align; c5: 10^6 genomes of 30 kb with
50 core 25-mers planted in 20 000 genomes each, prefilter only.  Every rank GENERATES only its own block (whole families,
synth.make_family_block), so no process ever holds the whole set.  Prints one JSON line on rank 0; with --check N the
first N genomes are also run on rank 0 alone and compared (a family-closed subset: its pairs and statistics must match)."""
import argparse
import json
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import torch.distributed as dist

from vclust_b200 import api, distributed, synth

ap = argparse.ArgumentParser()
ap.add_argument("config", choices=["c4", "c5"])
ap.add_argument("--n", type=int, default=0)
ap.add_argument("--runs", type=int, default=2)
ap.add_argument("--no-align", action="store_true")
ap.add_argument("--check", type=int, default=0)
args = ap.parse_args()
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
cfg = dict(synth.BLOCK_CONFIGS[args.config])
if args.n:
    cfg["n"] = args.n
n = cfg["n"]
blocks = synth.family_blocks(n, cfg["family"], world)
t0 = time.perf_counter()
names_l, seqs_l = synth.make_family_block(blocks[rank][0], blocks[rank][1], **cfg)
gen_s = time.perf_counter() - t0
lens_l = [int(s.size) for s in seqs_l]
all_lens = [None] * world
dist.all_gather_object(all_lens, lens_l)
lengths = [x for part in all_lens for x in part]
names = ["g%07d" % i for i in range(n)]
t0 = time.perf_counter()
run = distributed.ShardedRun(dist, lr, names, lengths, seqs_l, blocks=blocks)
torch.cuda.synchronize(); dist.barrier()
setup_s = time.perf_counter() - t0
out = {"config": args.config, "genomes": n, "bases": int(sum(lengths)), "gpus": world, "generate_s_rank0": round(gen_s, 1),
       "setup_s": round(setup_s, 2), "runs": []}
for rep in range(args.runs):
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    pairs = run.prefilter()
    torch.cuda.synchronize(); dist.barrier()
    t1 = time.perf_counter()
    pre = run.ctx.timings("prefilter")
    n_dir = 0
    res = None
    if not args.no_align and args.config == "c4":
        res = run.align()
        n_dir = res.n
    torch.cuda.synchronize(); dist.barrier()
    t2 = time.perf_counter()
    aln = run.ctx.timings("align") if res is not None else {}
    tm = torch.tensor([pre.get(k, 0.0) for k in ("extract_ms", "sort_ms", "segment_ms", "exchange_ms", "emit_ms")] +
                      [aln.get(k, 0.0) for k in ("list_ms", "index_ms", "parse_ms", "gather_ms")], dtype=torch.float64, device="cuda")
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    if rank == 0:
        peak = 6542.1
        pre_bytes = 24.25 * sum(lengths) + 12 * pairs.n_pairs + 4 * n
        r = {"candidate_pairs": pairs.n_pairs, "directed_parses": int(n_dir), "prefilter_s": round(t1 - t0, 4), "align_s": round(t2 - t1, 4),
             "pairs_per_s": round(pairs.n_pairs / (t2 - t0), 1),
             "max_rank_ms": dict(zip(["extract", "partition", "group", "exchange", "emit", "list", "index", "parse", "gather"], [round(x, 2) for x in tm.tolist()])),
             "prefilter_hbm_frac_per_gpu": round(pre_bytes / world / (t1 - t0) / 1e9 / peak, 4),
             "passes": pre.get("passes"), "bubbles_rank0": pre.get("bubbles"), "table_slots_rank0": pre.get("table_slots")}
        out["runs"].append(r)
    if rep + 1 < args.runs:
        pairs.close()
        if res is not None:
            res.close()
if rank == 0 and args.check:
    # a family-closed prefix on one GPU: every pair among the first `check` genomes, with its statistics
    m = args.check // cfg["family"] * cfg["family"]
    sub_names, sub_seqs = synth.make_family_block(0, m, **dict(cfg, n=n))
    with api.Context(lr) as ctx:
        g = api.Genomes.from_memory(sub_names, sub_seqs)
        one = api.prefilter_genomes(ctx, g)
        sel = pairs.rows < m          # rows < m imply cols < m
        assert (pairs.rows[sel].tolist(), pairs.cols[sel].tolist(), pairs.common[sel].tolist()) == (one.rows.tolist(), one.cols.tolist(), one.common.tolist()), "pairs differ"
        assert pairs.total_kmers[:m].tolist() == one.total_kmers.tolist()
        if res is not None:
            o1 = api.align_genomes(ctx, g, one)
            # (vectorised: a Python loop over 2 x 10^7 results would keep eight GPUs waiting for minutes)
            order = res.order.astype(np.int64)
            gr, gq = order[res.ref.astype(np.int64)], order[res.qry.astype(np.int64)]
            sel2 = (gr < m) & (gq < m)
            got_key = gr[sel2] * n + gq[sel2]
            o_order = o1.order.astype(np.int64)
            want_key = o_order[o1.ref.astype(np.int64)] * n + o_order[o1.qry.astype(np.int64)]
            ga, wa = np.argsort(got_key, kind="stable"), np.argsort(want_key, kind="stable")
            assert np.array_equal(got_key[ga], want_key[wa]), "directed pairs differ"
            assert np.array_equal(res.stats[sel2][ga], o1.stats[wa]), "align statistics differ"
        out["check"] = "first %d genomes equal the single-GPU result (%d pairs)" % (m, one.n_pairs)
        g.close()
if rank == 0:
    print(json.dumps(out))
dist.barrier()
run.close()
dist.destroy_process_group()
