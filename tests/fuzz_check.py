#!/usr/bin/env python
"""Randomised differential check of the align kernels against the CPU oracle (GPU box; not collected by pytest).

    python tests/fuzz_check.py [rounds=40] [seed=1]

Every round draws LZ-ANI parameters (mal, msl, mrd, mqd, reg, aw, am, ar) inside the ranges the C ABI accepts, a genome
set (length range, divergence, indel size, N runs, lower case, degenerate genomes) and a pair list (related pairs,
random pairs, self pairs, duplicates), then compares, for both kernel variants, the per-pair statistics and the
alignment regions with oracle.run_pairs_regions.  Exits non-zero at the first difference and prints the round's draw."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    from oracle import oracle
    from vclust_b200 import api, synth
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    ctx = api.Context(0)
    total_pairs = 0
    for it in range(rounds):
        mal = int(rng.integers(6, 17))
        msl = int(rng.integers(3, mal + 1))
        aw = int(rng.integers(2, 33))
        # mqd <= mrd: beyond that the reference itself is undefined -- its tail comparison (parser.cpp:713 -> :210-248) then
        # runs past the end of the reference text and reads whatever an earlier, longer reference left in the buffer
        # (found by this script: --mqd 187 --mrd 118); the kernels define symbols outside a text as mismatches
        mrd = int(rng.integers(1, 200))
        params = dict(mal=mal, msl=msl, mrd=mrd, mqd=int(rng.integers(0, mrd + 1)),
                      reg=int(rng.integers(1, 80)), aw=aw, am=int(rng.integers(0, aw + 1)), ar=int(rng.integers(1, 9)))
        if it % 4 == 0:
            params.update(mal=11, msl=7, aw=15, am=7, ar=3)                  # the specialised fast paths, odd mrd/mqd/reg
        lo = int(rng.integers(200, 3000))
        gen = dict(n=24, length=(lo, lo + int(rng.integers(1, 9000))), family=int(rng.choice([2, 3, 4, 6])),
                   seed=int(rng.integers(1, 1 << 30)), max_div=float(rng.choice([0.02, 0.1, 0.2, 0.35])),
                   indel=int(rng.choice([0, 30, 500, 2000])), n_frac=float(rng.choice([0, 0.3, 1.0])),
                   lower_frac=float(rng.choice([0, 0.5])))
        names, seqs = synth.make_genomes(**gen)
        seqs = [s.tobytes() for s in seqs] + [b"", b"AC", b"N" * 77, b"ACGTTGCA" * 60, b"A" * 500, seqs[0].tobytes()]
        names = names + ["x%d" % i for i in range(6)]
        n = len(names)
        ref = rng.integers(0, n, size=260)
        qry = rng.integers(0, n, size=260)
        fam = gen["family"]
        ref[:160] = np.arange(160) % 24
        qry[:160] = np.minimum((ref[:160] // fam) * fam + rng.integers(0, fam, size=160), 23)
        ref[160:170] = qry[160:170]                                          # self pairs
        ref[170:180], qry[170:180] = ref[:10], qry[:10]                      # duplicates
        g = api.Genomes.from_memory(names, seqs)
        ap = api.align_params(**params)
        st = api.align_pairs(ctx, g, ref, qry, ap)
        st2, regions = api.align_pairs_regions(ctx, g, ref, qry, ap)
        codes = [oracle.lz_codes(s) for s in seqs]
        want_st, want_regs = oracle.run_pairs_regions(codes, ref, qry, oracle.LzParams.default(**params))
        want_rows = sorted((int(ref[k]), int(qry[k]), ss, se, rs, re_, m, mm)
                           for k, rg in enumerate(want_regs) for rs, re_, ss, se, m, mm in rg.tolist())
        ok = np.array_equal(st, want_st) and np.array_equal(st2, want_st) and sorted(map(tuple, regions.table().tolist())) == want_rows
        regions.close(); g.close()
        total_pairs += ref.size
        if not ok:
            bad = np.nonzero((st != want_st).any(axis=1))[0][:5]
            print("MISMATCH in round %d: params %r genomes %r first differing pairs %r" % (it, params, gen, [(int(ref[b]), int(qry[b]), st[b].tolist(), want_st[b].tolist()) for b in bad]))
            sys.exit(1)
    print("fuzz ok: %d rounds, %d directed pairs, stats and regions identical to the oracle" % (rounds, total_pairs))


if __name__ == "__main__":
    main()
