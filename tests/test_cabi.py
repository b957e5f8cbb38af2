"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/vclust_b200.h
declares; host-only entry points (FASTA ingest, filter read/write, number formatting) behave like the reference's."""
import ctypes as C
import re

import numpy as np
import pytest

from oracle import oracle
from vclust_b200 import _lib, api, build


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    header = (_lib.HERE.parent / "include" / "vclust_b200.h").read_text()
    declared = set(re.findall(r"\b(vb_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name


def test_version_and_no_device_is_an_error(lib):
    assert "vclust-b200" in api.version()
    if api.device_count() == 0:
        with pytest.raises(api.VbError) as e:
            api.Context(0)
        assert "no CPU fallback" in str(e.value)


def test_fasta_ingest_matches_reference_rules(lib, golden, tmp_path):
    p = golden / "example" / "multifasta.fna.gz"
    g = api.Genomes.load([p], True, api.FASTA_LZANI)
    names, codes = oracle.load_genomes_lzani([p], True)
    assert g.names() == names
    assert [g.length(i) for i in range(len(g))] == [c.size for c in codes]
    g2 = api.Genomes.load([p], True, api.FASTA_KMERDB)
    recs = oracle.read_records_kmerdb(p)
    assert g2.names() == [n for n, _ in recs]
    assert [g2.length(i) for i in range(len(g2))] == [len(s) for _, s in recs]
    # corner cases: CRLF, blank lines, header with description, unterminated last line, directory mode
    fa = tmp_path / "x.fa"
    fa.write_bytes(b">a desc here\r\nACGT\r\n\r\nNNAC\n>b\nGG\nTT")
    g = api.Genomes.load([fa], True, api.FASTA_LZANI)
    assert g.names() == ["a", "b"] and [g.length(i) for i in range(2)] == [8, 2]      # "TT" dropped (lz-ani quirk)
    g = api.Genomes.load([fa], True, api.FASTA_KMERDB)
    assert g.names() == ["a", "b"] and [g.length(i) for i in range(2)] == [8, 4]
    g = api.Genomes.load([fa, fa], False, api.FASTA_LZANI, sep_len=40)
    assert g.names() == ["x.fa", "x.fa"] and g.length(0) == 8 + 40 + 4
    g = api.Genomes.load([fa], False, api.FASTA_KMERDB)
    assert g.length(0) == 8 + 1 + 4


def _spec_kmerdb(data: bytes):
    """kmer-db record rules, stated byte by byte (genome_input_file.h:287-338)."""
    recs = []
    pos = data.find(b">")
    while pos >= 0:
        eol = data.find(b"\n", pos)
        if eol < 0:
            eol = len(data)
        hend = eol - 1 if eol > pos + 1 and data[eol - 1:eol] == b"\r" else eol
        name = data[pos + 1:hend]
        nxt = data.find(b">", eol) if eol < len(data) else -1
        bend = nxt if nxt >= 0 else len(data)
        body = bytes(c for c in data[min(eol + 1, bend):bend] if c not in (10, 13))
        recs.append((name.split(b" ")[0] if b" " in name else name, body))
        pos = nxt
    return recs


def _spec_lzani(data: bytes, dir_mode: bool):
    """lz-ani record rules, stated line by line (seq_reservoir.cpp:90-210, file_wrapper.h:917-950)."""
    recs, cur, have = [], [b"", b""], False

    def push():
        if (have and cur[0]) or (dir_mode and (have or cur[1])):
            recs.append((cur[0], cur[1]))

    lines = data.split(b"\n")
    tail = lines.pop()
    if dir_mode and tail:
        lines.append(tail)
    for ln in lines:
        if ln.endswith(b"\r"):
            ln = ln[:-1]
        if not ln:
            continue
        if ln[:1] == b">":
            push()
            cur, have = [ln[1:], b""], True
        else:
            cur[1] += ln
    push()
    return [((n.split(b" ")[0] if b" " in n else n), s) for n, s in recs]


@pytest.mark.parametrize("par", [None, (0, 2), (0, 8), (64, 3)])
def test_fasta_ingest_fuzz_against_record_rules(lib, tmp_path, monkeypatch, par):
    """Random byte soup over the alphabet that matters (>, space, CR, LF, bases): the single-pass loader must give
    exactly the records the two tools' rules give, in all four (flavor, mode) combinations -- also when a multi-FASTA
    is cut at line-start '>' bytes and parsed by several threads (forced here on tiny files: par = (from bytes, threads))."""
    if par:
        monkeypatch.setenv("VB_FASTA_PAR_MIN", str(par[0]))
        monkeypatch.setenv("VB_FASTA_THREADS", str(par[1]))
    rng = np.random.default_rng(123)
    alphabet = np.frombuffer(b">> \r\n\n\n\nACGTNacgtUXY", dtype=np.uint8)
    for trial in range(60 if not par else 150):
        data = alphabet[rng.integers(0, alphabet.size, size=int(rng.integers(0, 400)))].tobytes()
        if trial % 3 == 0:
            data = b">s1 d\n" + data
        fa = tmp_path / ("f%d.fa" % trial)
        fa.write_bytes(data)
        want = _spec_kmerdb(data)
        g = api.Genomes.load([fa], True, api.FASTA_KMERDB)
        assert [(g.name(i).encode(), g.sequence(i)) for i in range(len(g))] == want, (trial, data)
        g = api.Genomes.load([fa, fa], True, api.FASTA_KMERDB)             # second file: appended behind the first one's records
        assert [(g.name(i).encode(), g.sequence(i)) for i in range(len(g))] == want + want, (trial, data)
        g = api.Genomes.load([fa], False, api.FASTA_KMERDB)
        assert len(g) == 1 and g.sequence(0) == b"N".join(s for _, s in want), (trial, data)
        want = _spec_lzani(data, False)
        g = api.Genomes.load([fa], True, api.FASTA_LZANI)
        assert [(g.name(i).encode(), g.sequence(i)) for i in range(len(g))] == want, (trial, data)
        g = api.Genomes.load([fa, fa], True, api.FASTA_LZANI)
        assert [(g.name(i).encode(), g.sequence(i)) for i in range(len(g))] == want + want, (trial, data)
        joined = b""
        for _, s in _spec_lzani(data, True):
            if joined:
                joined += b"N" * 7
            joined += s
        g = api.Genomes.load([fa, fa], False, api.FASTA_LZANI, sep_len=7)
        assert len(g) == 2 and g.sequence(0) == joined and g.sequence(1) == joined, (trial, data)


def test_fasta_ingest_big_file_read_by_several_threads(lib, tmp_path):
    """Files of 64 MB and more are read by several threads on disjoint ranges (host_fasta.cpp read_whole): the records
    must be the ones a single read gives, and a second file appended to the same store must start where the first ended."""
    from vclust_b200 import synth
    names, seqs = synth.make_genomes(n=1700, length=40_000, seed=5, family=10)
    big, small = tmp_path / "big.fna", tmp_path / "small.fna"
    synth.write_fasta(big, names, seqs)
    assert big.stat().st_size >= 64 << 20
    small.write_bytes(b">tail one\nACGT\nAC\n")
    for flavor in (api.FASTA_KMERDB, api.FASTA_LZANI):
        g = api.Genomes.load([big, small], True, flavor)
        assert len(g) == len(names) + 1 and g.names()[:-1] == names and g.name(len(names)) == "tail"
        assert all(g.sequence(i) == bytes(seqs[i]) for i in range(0, len(names), 13)) and g.sequence(len(names)) == b"ACGTAC"
        g.close()


def test_filter_roundtrip_and_format(lib, golden, tmp_path):
    p = golden / "example" / "multifasta.fna.gz"
    g = api.Genomes.load([p], True, api.FASTA_LZANI)
    pairs = api.read_filter(golden / "example" / "fltr.txt", 0.0, g)
    assert pairs.n_pairs == 13
    assert list(zip(pairs.rows.tolist(), pairs.cols.tolist()))[:3] == [(1, 0), (2, 0), (2, 1)]
    assert api.read_filter(golden / "example" / "fltr.txt", 0.99, g).n_pairs == 8
    # names must agree (lz_matcher.cpp:43-75)
    g_bad = api.Genomes.from_memory(["x", "y", "z"], [b"ACGT"] * 3)
    with pytest.raises(api.VbError):
        api.read_filter(golden / "example" / "fltr.txt", 0.0, g_bad)


def test_number_formats_match_oracle(lib):
    lib.vb_last_error  # noqa: B018  (library loaded)
    import ctypes
    so = ctypes.CDLL(str(_lib.LIB_PATH))
    f6 = so._Z13vb_fmt_fixed6dPc
    fr = so._Z11vb_fmt_realdiPc
    f6.argtypes = [ctypes.c_double, ctypes.c_char_p]
    fr.argtypes = [ctypes.c_double, ctypes.c_int, ctypes.c_char_p]
    rng = np.random.default_rng(7)
    vals = list(rng.random(2000)) + list(rng.random(500) * 1e-4) + [0.0, 1.0, 0.5, 100.0, 0.9999996, 1e-5, 1e-4, 0.7,
                                                                      89.28928929, 1234567.0, 0.016848, 0.01684796]
    buf = ctypes.create_string_buffer(64)
    for v in vals:
        n = f6(float(v), buf)
        assert buf.raw[:n].decode() == oracle.fixed6(float(v))
        for prec in (4, 6):
            n = fr(float(v), prec, buf)
            assert buf.raw[:n].decode() == oracle.real_str(float(v), prec), v


def test_write_ani_threads_give_identical_bytes(lib, tmp_path, monkeypatch):
    """Large results are formatted by several host threads on disjoint row ranges: same bytes as one thread, with
    output filters and every column set."""
    from vclust_b200 import synth
    names, seqs = synth.make_genomes(n=600, length=150, family=12, seed=3)
    g = api.Genomes.from_memory(names, seqs)
    rows, cols = [], []
    for i in range(600):
        for j in range((i // 12) * 12, i):
            rows.append(i); cols.append(j)
    rows, cols = np.array(rows, dtype=np.uint32), np.array(cols, dtype=np.uint32)
    ref, qry = np.concatenate([rows, cols, rows[:50]]), np.concatenate([cols, rows, cols[:50]])     # + duplicated pairs
    st = np.random.default_rng(5).integers(0, 150, size=(ref.size, 3)).astype(np.int32)
    res = api.align_result_from_pairs(g, ref, qry, st)
    for fmt, flt in (("complete", None), ("lite", {"ani": 0.4, "qcov": 0.3})):
        outs = []
        for thr in ("1", "7"):
            monkeypatch.setenv("VB_WRITE_THREADS", thr)
            p = tmp_path / ("ani_%s_%s.tsv" % (fmt, thr))
            api.write_ani(g, res, p, None, api.ALIGN_OUTFMT[fmt], flt)
            outs.append(p.read_bytes())
        assert outs[0] == outs[1] and outs[0].count(b"\n") > 1000


def test_read_filter_token_rules(lib, tmp_path):
    """lz-ani splits a row at ',' and every token at ':' with a split() that drops an empty last part (L/utils.cpp:15-36,
    L/filter.cpp:60-81): only tokens with exactly two parts count."""
    names = ["ga", "gb", "gc", "gd"]
    g = api.Genomes.from_memory(names, [b"ACGT"] * 4)
    p = tmp_path / "f.txt"
    p.write_text("kmer-length: 25 fraction: 1 ,ga,gb,gc,gd,\n"
                 "ga,\n"
                 "gb,1:0.900000,,junk,2:,1:0.5:7,\r\n"
                 "\n"                                   # lines of up to 2 bytes are skipped, row id not advanced
                 "gc,2:0.800000,1:1e-1,4:0.75\n"         # last token without a trailing comma still counts
                 "gd,3:0.650000,\n")
    pr = api.read_filter(p, 0.0, g)
    assert list(zip(pr.rows.tolist(), pr.cols.tolist(), pr.ani.tolist())) == [(1, 0, 0.9), (2, 1, 0.8), (2, 0, 0.1), (2, 3, 0.75), (3, 2, 0.65)]
    pr = api.read_filter(p, 0.7, g)
    assert list(zip(pr.rows.tolist(), pr.cols.tolist())) == [(1, 0), (2, 1), (2, 3)]
    p.write_text("kmer-length: 25 fraction: 1 ,ga,gb,gc,gd,\ngb,:0.9,\n")
    with pytest.raises(api.VbError):                     # id -1 is out of range
        api.read_filter(p, 0.0, g)
