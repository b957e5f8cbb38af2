#!/usr/bin/env python
"""Where the file-to-file time goes (not collected by pytest): python tests/file_timing.py [c3]
Times the phases of `vclust prefilter` + `vclust align` through the file-level API on one workload."""
import json
import sys
import tempfile
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from vclust_b200 import api, synth

cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
names, seqs = synth.make_genomes(**synth.CONFIGS[cfg])
out = {}
with tempfile.TemporaryDirectory() as td:
    td = Path(td)
    fa = td / "in.fna"
    synth.write_fasta(fa, names, seqs)
    for rep in range(2):
        t = {}
        t0 = time.perf_counter()
        ctx = api.Context(0); t["ctx_create"] = time.perf_counter() - t0; t0 = time.perf_counter()
        g = api.Genomes.load([fa], True, api.FASTA_KMERDB); t["load_fasta"] = time.perf_counter() - t0; t0 = time.perf_counter()
        pairs = api.prefilter_genomes(ctx, g); t["prefilter"] = time.perf_counter() - t0; t0 = time.perf_counter()
        api.write_filter(g, pairs, td / "fltr.txt"); t["write_filter"] = time.perf_counter() - t0; t0 = time.perf_counter()
        pairs.close(); g.close(); ctx.close(); t["close"] = time.perf_counter() - t0; t0 = time.perf_counter()
        ctx = api.Context(0); t["ctx_create2"] = time.perf_counter() - t0; t0 = time.perf_counter()
        g = api.Genomes.load([fa], True, api.FASTA_LZANI, sep_len=40); t["load_fasta2"] = time.perf_counter() - t0; t0 = time.perf_counter()
        flt = api.read_filter(td / "fltr.txt", 0.0, g); t["read_filter"] = time.perf_counter() - t0; t0 = time.perf_counter()
        res = api.align_genomes(ctx, g, flt); t["align"] = time.perf_counter() - t0; t0 = time.perf_counter()
        t["align_gpu_ms"] = ctx.timings("align")
        api.write_ani(g, res, td / "ani.tsv"); t["write_ani"] = time.perf_counter() - t0; t0 = time.perf_counter()
        res.close(); flt.close(); g.close(); ctx.close(); t["close2"] = time.perf_counter() - t0
        out["pass%d" % rep] = {k: (round(v, 4) if isinstance(v, float) else v) for k, v in t.items()}
print(json.dumps(out))
