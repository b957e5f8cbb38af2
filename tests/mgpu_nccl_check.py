"""NCCL check of the multi-GPU pipeline (run under torchrun on a multi-GPU box; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_nccl_check.py [c4mini] [ref]

Every rank loads its block of the genome set; the sharded result (rank 0) must equal the single-GPU result computed by
rank 0 on the whole set, and -- with `ref` -- the output files must be byte-identical to the reference binaries'."""
import os
import sys
import tempfile
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import torch.distributed as dist

from vclust_b200 import api, distributed, synth

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
# default: 400 x 20 kb; "c4mini": variable lengths 5-200 kb, families of 50, N runs and lower-case blocks (c4's shape)
if "c4mini" in sys.argv:
    names, seqs = synth.make_genomes(n=1500, length=(5000, 200000), family=50, seed=synth.BASE_SEED + 4, n_frac=0.01, lower_frac=0.01)
else:
    names, seqs = synth.make_genomes(n=400, length=20000, family=10, seed=77, n_frac=0.05)
lengths = [int(s.size) for s in seqs]
first, count = distributed.block_partition(lengths, world)[rank]
run = distributed.ShardedRun(dist, lr, names, lengths, seqs[first:first + count])
assert not run.comm.staged
pairs = run.prefilter()
res = run.align()
if rank == 0:
    with api.Context(lr) as ctx:
        g = api.Genomes.from_memory(names, seqs)
        full = api.prefilter_genomes(ctx, g)
        assert (pairs.rows.tolist(), pairs.cols.tolist(), pairs.common.tolist()) == (full.rows.tolist(), full.cols.tolist(), full.common.tolist()), "pairs differ"
        assert np.array_equal(pairs.ani, full.ani)
        assert pairs.total_kmers.tolist() == full.total_kmers.tolist()
        one = api.align_genomes(ctx, g, full)
        assert (res.ref.tolist(), res.qry.tolist(), res.stats.tolist()) == (one.ref.tolist(), one.qry.tolist(), one.stats.tolist()), "align results differ"
        msg = "MGPU OK: %d genomes, %d bases, %d pairs, %d directed parses over %d ranks (NCCL), %.1f MB sent by rank 0 in %d collectives" % (
            len(names), sum(lengths), full.n_pairs, res.n, world, run.comm.bytes_sent / 1e6, run.comm.calls)
        if "ref" in sys.argv:
            from oracle import oracle
            with tempfile.TemporaryDirectory() as td:
                td = Path(td)
                synth.write_fasta(td / "in.fna", names, seqs)
                api.write_filter(run.meta, pairs, td / "fltr.txt")
                api.write_ani(run.meta, res, td / "ani.tsv", None, api.ALIGN_OUTFMT["complete"])
                oracle.ref_prefilter([td / "in.fna"], td / "ref_fltr.txt", td / "p")
                oracle.ref_align([td / "in.fna"], td / "ref_ani.tsv", td / "a", filter_path=td / "ref_fltr.txt", columns=api.ALIGN_OUTFMT["complete"])
                assert (td / "fltr.txt").read_bytes() == (td / "ref_fltr.txt").read_bytes(), "filter differs from kmer-db"
                assert (td / "ani.tsv").read_bytes() == (td / "ref_ani.tsv").read_bytes(), "ani.tsv differs from lz-ani"
                assert (td / "ani.ids.tsv").read_bytes() == (td / "ref_ani.ids.tsv").read_bytes()
            msg += "; files byte-identical to the reference binaries"
        print(msg)
        g.close()
pairs.close(); res.close()
dist.barrier()
run.close()
dist.destroy_process_group()
