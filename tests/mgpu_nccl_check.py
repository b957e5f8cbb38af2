"""NCCL 2-rank check of vclust_b200.distributed (run under torchrun): sharded result == single-GPU result."""
import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch, torch.distributed as dist
from vclust_b200 import api, distributed, synth
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
# default: 200 x 20 kb; "c4mini": variable lengths 5-200 kb, families of 50, N runs and lower-case blocks (c4's shape)
if len(sys.argv) > 1 and sys.argv[1] == "c4mini":
    names, seqs = synth.make_genomes(n=1500, length=(5000, 200000), family=50, seed=synth.BASE_SEED + 4, n_frac=0.01, lower_frac=0.01)
else:
    names, seqs = synth.make_genomes(n=200, length=20000, family=10, seed=77, n_frac=0.05)
ctx = api.Context(lr)
g = api.Genomes.from_memory(names, seqs)
res = distributed.prefilter_align_sharded(ctx, g, g, dist, torch.device("cuda", lr),
                                          passes=3 if "passes3" in sys.argv else 0)
if rank == 0:
    full = api.prefilter_genomes(ctx, g)
    assert list(zip(res["pairs"][0].tolist(), res["pairs"][1].tolist(), res["pairs"][2].tolist())) == \
        list(zip(full.rows.tolist(), full.cols.tolist(), full.common.tolist())), "pairs differ"
    assert np.array_equal(res["pairs"][3], full.ani)
    assert res["totals"].tolist() == full.total_kmers.tolist()
    ref = np.concatenate([full.rows, full.cols]); qry = np.concatenate([full.cols, full.rows])
    st = api.align_pairs(ctx, g, ref, qry)
    want = {(int(r), int(q)): tuple(s) for r, q, s in zip(ref, qry, st.tolist())}
    got = {(int(r), int(q)): tuple(s) for r, q, s in zip(res["ref"], res["qry"], res["stats"].tolist())}
    assert got == want, "align stats differ"
    print("MGPU OK: %d genomes, %d bases, %d pairs, %d directed parses over %d ranks" %
          (len(names), sum(s.size for s in seqs), full.n_pairs, len(got), world))
dist.barrier()
dist.destroy_process_group()
