"""Unfiltered all-vs-all align of c2 on one GPU (SURVEY 8(d) secondary measurement; not collected by pytest)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from vclust_b200 import api, synth
names, seqs = synth.make_genomes(**synth.CONFIGS["c2"])
ctx = api.Context(0)
g = api.Genomes.from_memory(names, seqs)
ctx.make_resident(g, api.FASTA_LZANI, 40)
for rep in range(2):
    t = time.perf_counter()
    res = api.align_genomes(ctx, g, None)
    dt = time.perf_counter() - t
    print("all-vs-all c2: %d directed parses in %.3f s = %.2f M parses/s; device %s" % (res.n, dt, res.n / dt / 1e6, {k: round(v, 2) for k, v in ctx.timings("align").items() if k in ("total_ms", "index_ms", "parse_ms", "host_prep_ms")}))
    res.close()
