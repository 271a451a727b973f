"""Secondary path on the GPU (reference src/peptides.rs): codon translation kernel and the open-addressing
hash probe behind `filter` / `build_reference`, against the reference goldens and the oracle."""
import os
import random
import struct
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

CODONS = {}
for aa, cs in [("I", "ATT ATC ATA"), ("L", "CTT CTC CTA CTG TTA TTG"), ("V", "GTT GTC GTA GTG"), ("F", "TTT TTC"), ("M", "ATG"), ("C", "TGT TGC"),
               ("A", "GCT GCC GCA GCG"), ("G", "GGT GGC GGA GGG"), ("P", "CCT CCC CCA CCG"), ("T", "ACT ACC ACA ACG"),
               ("S", "TCT TCC TCA TCG AGT AGC"), ("Y", "TAT TAC"), ("W", "TGG"), ("Q", "CAA CAG"), ("N", "AAT AAC"), ("H", "CAT CAC"),
               ("E", "GAA GAG"), ("D", "GAT GAC"), ("K", "AAA AAG"), ("R", "CGT CGC CGA CGG AGA AGG"), ("X", "TAA TAG TGA")]:
    for c in cs.split():
        CODONS[c] = aa  # reference src/peptides.rs:87-109


def to_protein(s, frame):
    """plain-Python restatement of peptides.rs:128-146 used as the checker for the translation kernel"""
    r = s.upper()
    if frame < 0:
        r = r[::-1].translate(str.maketrans("ACGT", "TGCA"))
    return "".join(CODONS[r[i:i + 3]] for i in range(0, len(r) - 2, 3))


@pytest.mark.parametrize("case,suffix", [("test_filter", "filtered"), ("test_filter_long", "filtered_long"), ("test_filter_fs", "filtered_fs")])
def test_cuda_filter_matches_golden(product, case, suffix, tmp_path):
    d = os.path.join(GOLDEN, case)
    with open(tmp_path / ("tumor.%s.fa" % suffix), "wb") as fo:
        r = subprocess.run([product[1], "filter", "--reference", os.path.join(d, "reference.binary"), "-l", "9", "--tsv", os.path.join(d, "info.tsv"),
                            "--tsv-output", str(tmp_path / ("info.%s.tsv" % suffix)), "--normal-output", str(tmp_path / ("normal.%s.fa" % suffix)),
                            "-s", str(tmp_path / "removed.tsv"), "-p", str(tmp_path / "removed.fa")], stdout=fo, stderr=subprocess.PIPE)
    assert r.returncode == 0, r.stderr.decode()
    for name in ("tumor.%s.fa", "normal.%s.fa", "info.%s.tsv"):
        name = name % suffix
        assert open(tmp_path / name, "rb").read() == open(os.path.join(d, "expected_output", name), "rb").read(), name


def _set_items(path):
    b = open(path, "rb").read()
    n, o, out = struct.unpack_from("<Q", b, 0)[0], 8, set()
    for _ in range(n):
        ln = struct.unpack_from("<Q", b, o)[0]
        out.add(b[o + 8:o + 8 + ln])
        o += 8 + ln
    assert o == len(b)
    return out


def test_cuda_build_reference_matches_golden(product, tmp_path):
    d = os.path.join(GOLDEN, "test_build")
    with open(tmp_path / "ref.fasta", "wb") as fo:
        r = subprocess.run([product[1], "build_reference", "--reference", os.path.join(d, "reference.fa"), "-l4", "--output", str(tmp_path / "ref.bin")],
                           stdout=fo, stderr=subprocess.PIPE)
    assert r.returncode == 0, r.stderr.decode()
    assert open(tmp_path / "ref.fasta", "rb").read() == open(os.path.join(d, "expected_output", "reference_peptides.fasta"), "rb").read()
    assert _set_items(tmp_path / "ref.bin") == _set_items(os.path.join(d, "expected_output", "reference.binary"))


def test_translation_kernel_matches_restatement(product):
    import microphaser_b200 as m
    rng = random.Random(5)
    seqs, frames = [], []
    for _ in range(5000):
        n = rng.choice([27, 27, 27, 21, 22, 30, 31, 2, 3, 5, 60])
        s = "".join(rng.choice("ACGTacgt") for _ in range(n))
        seqs.append(s.encode())
        frames.append(rng.choice([1, -1]))
    seqs.append(b"ACGNNACGT")  # unknown codon: flagged, the reference would panic
    frames.append(1)
    ctx = m.Context(0)
    aa, bad = ctx.translate(seqs, frames)
    ctx.close()
    for s, f, a, b in zip(seqs[:-1], frames[:-1], aa[:-1], bad[:-1]):
        assert b == 0 and a.decode() == to_protein(s.decode(), f), (s, f, a)
    assert bad[-1] == 1


def test_filter_with_real_hits_matches_oracle(product, oracle_bin, tmp_path):
    """The reference's own filter fixtures never hit the set; build a normal peptidome with
    build_reference from sequences that share windows with the tumor rows so that the removed-* outputs are exercised."""
    rng = random.Random(11)
    stop = {"TAA", "TAG", "TGA"}
    def codon():
        while True:
            c = "".join(rng.choice("ACGT") for _ in range(3))
            if c not in stop:
                return c
    hdr = ("id transcript gene_id gene_name chrom offset frame freq depth nvar nsomatic nvariant_sites nsomvariant_sites strand variant_sites "
           "somatic_positions somatic_aa_change germline_positions germline_aa_change normal_sequence mutant_sequence").split()
    rows, normal_fa = [], []
    off = 1000
    for t in range(60):
        wt = "".join(codon() for _ in range(9))
        mt = list(wt)
        p = rng.randrange(27)
        mt[p] = rng.choice([b for b in "acgt" if b.upper() != wt[p]])
        mt = "".join(mt)
        if any(mt.upper()[i:i + 3] in stop for i in range(0, 27, 3)):
            continue
        strand = "Forward"
        rid = "%015x" % rng.getrandbits(60) + "F"
        rows.append([rid, "ENST%05d" % (t // 3), "ENSG1", "G1", "chr1", str(off), "0", repr(rng.choice([0.25, 0.5, 0.3333333333333333])),
                     str(rng.randint(10, 60)), "1", "1", "1", "1", strand, str(off + p), str(off + p), "", "", "", wt, mt])
        off += 3
        if t % 2 == 0:
            normal_fa.append((rid, mt.upper()))  # half of the tumor peptides are also in the healthy peptidome -> removed
        else:
            normal_fa.append((rid, wt))
    with open(tmp_path / "info.tsv", "w") as f:
        f.write("\t".join(hdr) + "\n")
        for r in rows:
            f.write("\t".join(r) + "\n")
    with open(tmp_path / "normal.fa", "w") as f:
        for rid, s in normal_fa:
            f.write(">%s\n%s\n" % (rid, s))
    outs = {}
    for name, binary in (("oracle", oracle_bin), ("cuda", product[1])):
        d = tmp_path / name
        d.mkdir()
        with open(d / "pep.fa", "wb") as fo:
            r = subprocess.run([binary, "build_reference", "-r", str(tmp_path / "normal.fa"), "-o", str(d / "peptides.bin"), "-l", "9"], stdout=fo, stderr=subprocess.PIPE)
        assert r.returncode == 0, r.stderr.decode()
        with open(d / "tumor.fa", "wb") as fo:
            r = subprocess.run([binary, "filter", "-r", str(d / "peptides.bin"), "-t", str(tmp_path / "info.tsv"), "-o", str(d / "info.filtered.tsv"),
                                "-s", str(d / "info.removed.tsv"), "-p", str(d / "peptides.removed.fa"), "-n", str(d / "normal.filtered.fa"), "-l", "9"],
                               stdout=fo, stderr=subprocess.PIPE)
        assert r.returncode == 0, r.stderr.decode()
        outs[name] = {n: open(d / n, "rb").read() for n in ("pep.fa", "tumor.fa", "info.filtered.tsv", "info.removed.tsv", "peptides.removed.fa", "normal.filtered.fa")}
        outs[name]["set"] = _set_items(d / "peptides.bin")
    assert outs["oracle"] == outs["cuda"]
    assert len(outs["cuda"]["peptides.removed.fa"]) > 0 and len(outs["cuda"]["tumor.fa"]) > 0


def test_hash_probe_at_scale(product):
    """BASELINE.json config 5b shape: 1 M probes (half drawn from the set) against ~8 M distinct 9-mers."""
    import microphaser_b200 as m
    rng = np.random.default_rng(7)
    letters = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8)
    n_set, n_q = 8_000_000, 1_000_000
    members = letters[rng.integers(0, 20, size=(n_set, 9))]
    others = letters[rng.integers(0, 20, size=(n_q // 2, 9))]
    picks = members[rng.integers(0, n_set, size=n_q // 2)]
    ctx = m.Context(0)
    ctx.set_load(members, 9)
    hit_in = ctx.set_probe(picks, 9)
    hit_out = ctx.set_probe(others, 9)
    ctx.close()
    assert hit_in.all()
    # a random 9-mer is in an 8 M subset of 20^9 with probability 1.6e-5: verify the few hits exactly
    member_keys = set(map(bytes, members[:0]))  # built lazily only if needed
    if hit_out.any():
        packed = (members.astype(np.uint64) @ (np.uint64(32) ** np.arange(8, -1, -1, dtype=np.uint64)))
        packed_o = (others.astype(np.uint64) @ (np.uint64(32) ** np.arange(8, -1, -1, dtype=np.uint64)))
        truth = np.isin(packed_o, packed)
        assert (truth == hit_out.astype(bool)).all()
    assert hit_out.sum() < 100


def test_long_peptides_probe_and_build_match(product, oracle_bin, tmp_path):
    """Peptide lengths above 12 (MHC-II runs use 13-25; `-l` has no upper bound in src/cli.yaml): the device set compares
    bytes instead of 5-bit packed keys. Probe against a Python set, build_reference + filter at -l 15 against the oracle."""
    import microphaser_b200 as m
    rng = np.random.default_rng(5)
    letters = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWXY", dtype=np.uint8)
    for k in (13, 15, 25):
        members = letters[rng.integers(0, len(letters), size=(200_000, k))]
        members[1000:2000] = members[:1000]  # duplicates in the input
        others = letters[rng.integers(0, len(letters), size=(50_000, k))]
        near = members[rng.integers(0, len(members), size=50_000)].copy()
        near[:, k - 1] = letters[rng.integers(0, len(letters), size=50_000)]  # differ (mostly) in the last letter only
        truth = set(map(bytes, members))
        ctx = m.Context(0)
        ctx.set_load(members, k)
        for q in (members[::7], others, near):
            got = ctx.set_probe(np.ascontiguousarray(q), k)
            want = np.array([bytes(x) in truth for x in q], dtype=bool)
            assert (got.astype(bool) == want).all(), k
        ctx.close()
    # files: a healthy proteome with repeats, windows of 15 amino acids
    cod = [c for c, a in CODONS.items() if a != "X"]
    prng = random.Random(3)
    with open(tmp_path / "normal.fa", "w") as f:
        for i in range(40):
            s = "".join(prng.choice(cod) for _ in range(prng.randint(15, 40)))
            if i % 5 == 4:
                s = s[:45] + s[:45]
            f.write(">tx%d\n%s\n" % (i, s))
    outs = {}
    for name, binary in (("oracle", oracle_bin), ("cuda", product[1])):
        d = tmp_path / name
        d.mkdir()
        with open(d / "pep.fa", "wb") as fo:
            r = subprocess.run([binary, "build_reference", "-r", str(tmp_path / "normal.fa"), "-o", str(d / "peptides.bin"), "-l", "15"], stdout=fo, stderr=subprocess.PIPE)
        assert r.returncode == 0, r.stderr.decode()
        outs[name] = {"pep.fa": open(d / "pep.fa", "rb").read(), "set": _set_items(d / "peptides.bin")}
    assert outs["oracle"]["set"] == outs["cuda"]["set"] and len(outs["cuda"]["set"]) > 100
    assert outs["oracle"]["pep.fa"] == outs["cuda"]["pep.fa"]
