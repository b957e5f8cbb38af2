"""Deterministic synthetic phage-like genome sets.

Family model, so that the prefilter is non-trivial: ``n // family`` families; each family has a uniform random
ACGT root of length L; every member is the root with a per-member substitution rate d ~ U(0, max_div), one
deletion and one tandem duplication of length U(0, 500), and is reverse-complemented with probability 0.5.
Names are ``g%06d``; FASTA lines are 80 columns.

The generator is pure numpy with a seeded ``default_rng`` so the same (n, L, family, seed) gives the same bytes
here, on the GPU box and in the tests.
"""
from __future__ import annotations

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTacgtNn", b"TGCAtgcaNn"):
    _COMP[_a] = _b

BASE_SEED = 20261017


def _member(rng: np.random.Generator, root: np.ndarray, max_div: float, indel: int) -> np.ndarray:
    seq = root.copy()
    d = rng.uniform(0.0, max_div)
    n_sub = rng.binomial(seq.size, d)
    if n_sub:
        pos = rng.integers(0, seq.size, size=n_sub)
        # substitute by a different base: add 1..3 mod 4 in code space
        code = np.searchsorted(_ACGT, seq[pos])
        seq[pos] = _ACGT[(code + rng.integers(1, 4, size=n_sub)) & 3]
    if indel and seq.size > 4 * indel:
        dl = int(rng.integers(0, indel + 1))
        p = int(rng.integers(0, seq.size - dl))
        seq = np.concatenate([seq[:p], seq[p + dl:]])
        ul = int(rng.integers(0, indel + 1))
        p = int(rng.integers(0, seq.size - ul))
        seq = np.concatenate([seq[:p + ul], seq[p:]])          # tandem duplication of seq[p:p+ul]
    if rng.random() < 0.5:
        seq = _COMP[seq[::-1]]
    return seq


def make_genomes(n: int, length: int | tuple[int, int] = 40_000, family: int = 20, seed: int = BASE_SEED,
                 max_div: float = 0.12, indel: int = 500, n_frac: float = 0.0, lower_frac: float = 0.0):
    """Return (names, seqs) with seqs a list of uint8 ASCII arrays.

    ``length`` is an int or a (lo, hi) range sampled log-uniformly per family.  ``n_frac`` of the genomes get a run
    of 100 'N' and ``lower_frac`` get a lower-case block (exercise the symbol-coding rules of both stages).
    """
    rng = np.random.default_rng(seed)
    names, seqs = [], []
    g = 0
    while g < n:
        if isinstance(length, tuple):
            L = int(np.exp(rng.uniform(np.log(length[0]), np.log(length[1]))))
        else:
            L = int(length)
        root = _ACGT[rng.integers(0, 4, size=L)]
        for _ in range(min(family, n - g)):
            s = _member(rng, root, max_div, indel)
            if n_frac and rng.random() < n_frac and s.size > 1000:
                p = int(rng.integers(0, s.size - 100))
                s[p:p + 100] = ord("N")
            if lower_frac and rng.random() < lower_frac and s.size > 1000:
                p = int(rng.integers(0, s.size - 500))
                s[p:p + 500] |= 0x20
            names.append("g%06d" % g)
            seqs.append(s)
            g += 1
    return names, seqs


def family_blocks(n: int, family: int, world: int):
    """Whole families per rank (equal numbers of families): [(first genome, genome count)]; used where every rank
    generates only its own block (c4, c5 on 8 GPUs) -- make_family_block."""
    n_fam = (n + family - 1) // family
    cuts = [n_fam * r // world for r in range(world + 1)]
    return [(min(cuts[r] * family, n), min(cuts[r + 1] * family, n) - min(cuts[r] * family, n)) for r in range(world)]


def make_family_block(first: int, count: int, n: int, length=40_000, family: int = 20, seed: int = BASE_SEED, max_div: float = 0.12,
                      indel: int = 500, n_frac: float = 0.0, lower_frac: float = 0.0, core: tuple | None = None):
    """Genomes [first, first + count) of a set of n (first a multiple of `family`), every FAMILY seeded on its own
    (``default_rng([seed, family index])``), so that any rank can generate any block without generating the rest.
    (A different random stream than make_genomes: sets made by the two functions differ.)
    core = (n_core, every): n_core "core" 25-mers; genome g carries core (g mod n_core) when g mod every < n_core ... i.e.
    each core k-mer is planted in n / every genomes (50 cores in 20 000 of 10^6 genomes each)."""
    assert first % family == 0
    names, seqs = [], []
    cores = None
    if core:
        crng = np.random.default_rng([seed, 0x7fffffff])
        cores = [_ACGT[crng.integers(0, 4, size=25)] for _ in range(core[0])]
    g = first
    while g < min(first + count, n):
        rng = np.random.default_rng([seed, g // family])
        if isinstance(length, tuple):
            L = int(np.exp(rng.uniform(np.log(length[0]), np.log(length[1]))))
        else:
            L = int(length)
        root = _ACGT[rng.integers(0, 4, size=L)]
        for _ in range(min(family, n - g)):
            sq = _member(rng, root, max_div, indel)
            if n_frac and rng.random() < n_frac and sq.size > 1000:
                p = int(rng.integers(0, sq.size - 100))
                sq[p:p + 100] = ord("N")
            if lower_frac and rng.random() < lower_frac and sq.size > 1000:
                p = int(rng.integers(0, sq.size - 500))
                sq[p:p + 500] |= 0x20
            if cores is not None and g % core[1] < core[0] and sq.size > 2000:
                sq[1000:1025] = cores[g % core[1]]
            names.append("g%07d" % g)
            seqs.append(sq)
            g += 1
    return names, seqs


def fasta_bytes(names, seqs, width: int = 80) -> bytes:
    out = bytearray()
    for name, s in zip(names, seqs):
        out += b">" + name.encode() + b"\n"
        n = s.size
        full = (n // width) * width
        if full:
            body = np.empty((n // width, width + 1), dtype=np.uint8)
            body[:, :width] = s[:full].reshape(-1, width)
            body[:, width] = 10
            out += body.tobytes()
        if n > full:
            out += s[full:].tobytes() + b"\n"
    return bytes(out)


def write_fasta(path, names, seqs, width: int = 80) -> None:
    with open(path, "wb") as fh:
        fh.write(fasta_bytes(names, seqs, width))


# BASELINE.json configs (c2..c5) as generator arguments; c1 is example/multifasta.fna (reference fixture).
CONFIGS = {
    "c2": dict(n=1_000, length=40_000, family=20, seed=BASE_SEED + 2),
    "c3": dict(n=10_000, length=40_000, family=20, seed=BASE_SEED + 3),
    "c3_s200": dict(n=10_000, length=40_000, family=200, seed=BASE_SEED + 3),
    # more than 11 585 genomes: N(N-1)/2 > 2^26, the prefilter switches to the hashed pair table by itself
    "n20k": dict(n=20_000, length=10_000, family=20, seed=BASE_SEED + 6),
    # 1.2 x 10^9 k-mers: more than one prefilter pass holds, two passes over k-mer hash shards
    "n30k": dict(n=30_000, length=40_000, family=20, seed=BASE_SEED + 7),
    "c4": dict(n=100_000, length=(5_000, 200_000), family=200, seed=BASE_SEED + 4, n_frac=0.01, lower_frac=0.01),
    "c5": dict(n=1_000_000, length=30_000, family=20, seed=BASE_SEED + 5),
}
# the 8-GPU configurations, generated block by block with make_family_block (c5: 50 core 25-mers, each in 20 000 genomes)
BLOCK_CONFIGS = {
    "c4": dict(n=100_000, length=(5_000, 200_000), family=200, seed=BASE_SEED + 4, n_frac=0.01, lower_frac=0.01),
    "c5": dict(n=1_000_000, length=30_000, family=20, seed=BASE_SEED + 5, core=(50, 50)),
}
