"""Ingest: FASTA pair -> symbols by the rule of scripts/prepare-alignments.py:99-111, bit exact; binary
container round trip; the reference's text format.  Host only."""
import os

import numpy as np
import pytest

from conftest import example_symbols


def python_rule(a, b):
    clean = set("ACGT")
    return np.array([2 if (x.upper() not in clean or y.upper() not in clean) else (0 if x.upper() == y.upper() else 1)
                     for x, y in zip(a, b)], dtype=np.uint8)


def test_pair_rule_bit_exact_on_every_byte_pair():
    import imcoalhmm_b200 as m
    alphabet = "ACGTacgtNn-.*RYKMSWBDHVX? "
    a = "".join(x for x in alphabet for _ in alphabet)
    b = alphabet * len(alphabet)
    f = m.Forwarder.from_pair(a, b)
    assert f.NSYM == 3 and np.array_equal(f._seq.symbols(), python_rule(a, b))
    with pytest.raises(ValueError):
        m.Forwarder.from_pair("ACGT", "ACG")
    assert len(m.Forwarder.from_pair("", "")) == 0


def test_fasta_reader_and_named_records(tmp_path):
    import imcoalhmm_b200 as m
    rng = np.random.default_rng(0)
    seqs = {name: "".join(rng.choice(list("ACGTacgtN-"), size=1234, p=[.2, .2, .2, .2, .04, .04, .04, .04, .02, .02]))
            for name in ("hg18", "pantro2", "gorilla")}
    p = tmp_path / "aln.fa"
    with open(p, "w") as f:
        for name, s in seqs.items():
            f.write(">%s some description\r\n" % name)
            for i in range(0, len(s), 60):
                f.write(s[i:i + 60] + "\n")
            f.write("\n")
    f12 = m.Forwarder.from_fasta(str(p), names=("hg18", "pantro2"))
    assert np.array_equal(f12._seq.symbols(), python_rule(seqs["hg18"], seqs["pantro2"]))
    f31 = m.Forwarder.from_fasta(str(p), names=("gorilla", "hg18"))
    assert np.array_equal(f31._seq.symbols(), python_rule(seqs["gorilla"], seqs["hg18"]))
    with pytest.raises(ValueError):
        m.Forwarder.from_fasta(str(p))                       # three records: names are required
    with pytest.raises(ValueError):
        m.Forwarder.from_fasta(str(p), names=("hg18", "orang"))
    with pytest.raises(IOError):
        m.Forwarder.from_fasta(str(tmp_path / "missing.fa"))
    two = tmp_path / "two.fa"
    two.write_text(">a\nACGTN\n>b\nACCTA")
    assert m.Forwarder.from_fasta(str(two))._seq.symbols().tolist() == [0, 0, 1, 0, 2]


def test_binary_container_and_text_round_trip(tmp_path):
    import imcoalhmm_b200 as m
    sym = example_symbols()
    f = m.Forwarder.from_symbols(sym, 3)
    f.save(str(tmp_path / "x.imcseq"))
    assert os.path.getsize(tmp_path / "x.imcseq") == 24 + (len(sym) + 3) // 4
    g = m.Forwarder.load(str(tmp_path / "x.imcseq"))
    assert g.NSYM == 3 and np.array_equal(g._seq.symbols(), sym)
    f.write_text(str(tmp_path / "x.txt"))                    # the reference's own format ...
    h = m.Forwarder(str(tmp_path / "x.txt"), 3)              # ... read back by the reference-style constructor
    assert np.array_equal(h._seq.symbols(), sym)
    nine = m.Forwarder.from_symbols(np.arange(1000) % 9, 9)  # alphabets > 4 symbols are stored one byte each
    nine.save(str(tmp_path / "n.imcseq"))
    assert np.array_equal(m.Forwarder.load(str(tmp_path / "n.imcseq"))._seq.symbols(), np.arange(1000) % 9)
    (tmp_path / "bad.imcseq").write_bytes(b"IMCSEQ1\0" + b"\x03\0\0\0\x02\0\0\0" + (100).to_bytes(8, "little") + b"\0" * 3)
    with pytest.raises(IOError):
        m.Forwarder.load(str(tmp_path / "bad.imcseq"))        # truncated


@pytest.mark.skipif(not os.path.exists("/root/reference/examples/example_data.fa"), reason="reference tree not present")
def test_reference_example_alignment_matches_committed_fixture():
    import imcoalhmm_b200 as m
    f = m.Forwarder.from_fasta("/root/reference/examples/example_data.fa", names=("hg18", "pantro2"))
    assert np.array_equal(f._seq.symbols(), example_symbols())


def test_triplet_and_quartet_columns_bit_exact(tmp_path):
    """prepare-alignments.py:113-190: i1 + 4 i2 + 16 i3 (+ 32 i4, the script's own weight) or 64 / 128 for unclean columns."""
    import imcoalhmm_b200 as m
    rng = np.random.default_rng(1)
    nuc = {"A": 0, "C": 1, "G": 2, "T": 3}
    seqs = ["".join(rng.choice(list("ACGTacgtN-"), size=3000, p=[.22, .22, .22, .22, .02, .02, .02, .02, .02, .02])) for _ in range(4)]

    def rule(cols, weights, missing):
        out = []
        for bases in zip(*cols):
            up = [b.upper() for b in bases]
            out.append(sum(w * nuc[b] for w, b in zip(weights, up)) if all(b in nuc for b in up) else missing)
        return np.array(out, dtype=np.uint8)
    f3 = m.Forwarder.from_sequences(*seqs[:3])
    assert f3.NSYM == 65 and np.array_equal(f3._seq.symbols(), rule(seqs[:3], (1, 4, 16), 64))
    f4 = m.Forwarder.from_sequences(*seqs)
    assert f4.NSYM == 160 and np.array_equal(f4._seq.symbols(), rule(seqs, (1, 4, 16, 32), 128))
    f2 = m.Forwarder.from_sequences(seqs[0], seqs[1])
    assert f2.NSYM == 3 and np.array_equal(f2._seq.symbols(), python_rule(seqs[0], seqs[1]))
    p = tmp_path / "t.fa"
    p.write_text("".join(">s%d\n%s\n" % (i, s) for i, s in enumerate(seqs)))
    g3 = m.Forwarder.from_fasta(str(p), names=("s2", "s0", "s3"))
    assert np.array_equal(g3._seq.symbols(), rule([seqs[2], seqs[0], seqs[3]], (1, 4, 16), 64))
    with pytest.raises(ValueError):
        m.Forwarder.from_sequences("ACGT", "ACG", "ACGT")


def _write_phylip(path, seqs, layout, width=50):
    names = list(seqs)
    L = len(seqs[names[0]])
    with open(path, "w") as f:
        f.write(" %d %d\n" % (len(names), L))
        if layout == "sequential":                        # strict 10-column names, data over several lines per taxon
            for n in names:
                s = seqs[n]
                f.write(n.ljust(10) + s[:width] + "\n")
                for i in range(width, L, width):
                    f.write(" ".join(s[j:j + 10] for j in range(i, min(i + width, L), 10)) + "\n")
        else:
            for b, i in enumerate(range(0, L, width)):
                for n in names:
                    chunk = " ".join(seqs[n][j:j + 10] for j in range(i, min(i + width, L), 10))
                    if b == 0:
                        f.write((n.ljust(10) if layout == "strict" else n + "  ") + chunk + "\r\n")
                    else:
                        f.write((" " * 10 if layout == "strict" else "") + chunk + "\n")
                f.write("\n")


def test_phylip_readers_match_fasta(tmp_path):
    """prepare-alignments.py takes any BioPython format; the PHYLIP flavours give the same symbols as the FASTA reader."""
    import imcoalhmm_b200 as m
    rng = np.random.default_rng(1)
    seqs = {name: "".join(rng.choice(list("ACGTacgtN-"), size=1237, p=[.2, .2, .2, .2, .04, .04, .04, .04, .02, .02]))
            for name in ("hg18", "pantro2", "gorilla", "orang")}
    want2 = python_rule(seqs["hg18"], seqs["gorilla"])
    for layout, fmt in (("strict", "phylip"), ("relaxed", "phylip-relaxed"), ("sequential", "phylip-sequential")):
        p = tmp_path / (layout + ".phy")
        _write_phylip(p, seqs, layout)
        f = m.Forwarder.from_alignment(str(p), fmt, names=("hg18", "gorilla"))
        assert f.NSYM == 3 and np.array_equal(f._seq.symbols(), want2), layout
        q = m.Forwarder.from_alignment(str(p), fmt, names=("hg18", "pantro2", "gorilla", "orang"))
        assert q.NSYM == 160 and len(q) == 1237
        with pytest.raises(ValueError):
            m.Forwarder.from_alignment(str(p), fmt)                     # four records: names are required
        with pytest.raises(ValueError):
            m.Forwarder.from_alignment(str(p), fmt, names=("hg18", "bonobo"))
    fa = tmp_path / "four.fa"
    with open(fa, "w") as f:
        for n, s in seqs.items():
            f.write(">%s\n%s\n" % (n, s))
    q_fa = m.Forwarder.from_alignment(str(fa), "fasta", names=("hg18", "pantro2", "gorilla", "orang"))
    assert np.array_equal(q_fa._seq.symbols(), q._seq.symbols())
    assert np.array_equal(m.Forwarder.from_alignment(str(fa), names=("hg18", "gorilla"))._seq.symbols(), want2)
    two = tmp_path / "two.phy"
    two.write_text("2 5\nalpha     ACGTN\nbeta      ACCTA\n")
    assert m.Forwarder.from_alignment(str(two), "phylip")._seq.symbols().tolist() == [0, 0, 1, 0, 2]
    for bad in ("2 5\nalpha     ACGT\nbeta      ACCTA\n",           # a taxon shorter than the header says
                "2 5\nalpha     ACGTN\n",                           # a taxon missing
                "x y\nalpha     ACGTN\nbeta      ACCTA\n",          # no header
                "2 5\nalpha     AC.TN\nbeta      ACCTA\n"):         # match characters
        two.write_text(bad)
        with pytest.raises(ValueError):
            m.Forwarder.from_alignment(str(two), "phylip")
    with pytest.raises(IOError):
        m.Forwarder.from_alignment(str(tmp_path / "missing.phy"), "phylip")
    with pytest.raises(m.IMCError):
        m.Forwarder.from_alignment(str(fa), "nexus")


def test_hostile_headers_are_errors_not_crashes(tmp_path):
    """A PHYLIP header that announces more than the file can hold must come back as an error code (a C++
    exception crossing the C ABI would terminate the caller)."""
    import imcoalhmm_b200 as m
    p = tmp_path / "huge.phy"
    p.write_text("999999999999 10\nA         ACGTACGTAC\nB         ACGTACGTAC\n")
    with pytest.raises(ValueError):
        m.Forwarder.from_alignment(str(p), "phylip")
    q = tmp_path / "sites.phy"
    q.write_text("2 999999999999\nA         ACGTACGTAC\nB         ACGTACGTAC\n")
    with pytest.raises(ValueError):
        m.Forwarder.from_alignment(str(q), "phylip-sequential")
