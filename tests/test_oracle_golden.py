"""The CPU oracle against the reference's own golden outputs (reference tests/lib.rs).

Each case mirrors one `#[test]` of the reference: same inputs, same command line, outputs
compared byte-for-byte with the reference's checked-in `expected_output` files (copied to
tests/golden/<case>/expected by tests/golden/make_fixtures.py).
"""
import os
import subprocess

import pytest

from conftest import GOLDEN, materialize_reference

SOMATIC = ["forward_somatic", "empty", "reverse_somatic", "splice_forward_somatic", "splice_reverse_somatic"]


def run_cli(binary, case, tmp_path, gtf="annotation.gtf", extra_env=None):
    d = os.path.join(GOLDEN, case)
    meta = dict(l.rstrip("\n").split("\t") for l in open(os.path.join(d, "case.txt")))
    fa = materialize_reference(d, str(tmp_path))
    cmd = [binary, meta["subcommand"], os.path.join(d, "reads.bam"), "--ref", fa, "--variants",
           os.path.join(d, "variants.vcf"), "--tsv", str(tmp_path / "out.tsv")]
    if meta["subcommand"] == "somatic":
        cmd += ["--normal-output", str(tmp_path / "out.normal.fa")]
    env = dict(os.environ)
    env.update(extra_env or {})
    with open(os.path.join(d, gtf)) as gin, open(tmp_path / "out.fa", "wb") as fout:
        res = subprocess.run(cmd, stdin=gin, stdout=fout, stderr=subprocess.PIPE, env=env, timeout=600)
    return res


def assert_outputs(case, tmp_path):
    exp = os.path.join(GOLDEN, case, "expected")
    for name in sorted(os.listdir(exp)):
        want = open(os.path.join(exp, name), "rb").read()
        got = open(tmp_path / name, "rb").read()
        assert got == want, "%s/%s differs from the reference's expected output" % (case, name)


@pytest.mark.parametrize("case", SOMATIC)
def test_oracle_somatic_matches_reference_golden(oracle_bin, case, tmp_path):
    res = run_cli(oracle_bin, case, tmp_path)
    assert res.returncode == 0, res.stderr.decode()
    assert_outputs(case, tmp_path)


@pytest.mark.parametrize("case", ["forward_normal", "splice_forward_normal"])
def test_oracle_normal_matches_reference_golden(oracle_bin, case, tmp_path):
    """reference tests/lib.rs:237-249, 273-285 — only the FASTA is diffed by the reference's tests; its SHA-1
    ids hash sequence + transcript + offset, so they also pin the reference bases filled in from the fixture."""
    res = run_cli(oracle_bin, case, tmp_path)
    assert res.returncode == 0, res.stderr.decode()
    assert_outputs(case, tmp_path)


def test_oracle_unsorted_gtf_is_fatal(oracle_bin, tmp_path):
    """reference tests/lib.rs:344-382 — unsorted GTF must exit non-zero, sorted must exit zero."""
    res = run_cli(oracle_bin, "unsorted_gtf", tmp_path, gtf="unsorted.gtf")
    assert res.returncode != 0
    res = run_cli(oracle_bin, "unsorted_gtf", tmp_path, gtf="sorted.gtf")
    assert res.returncode == 0, res.stderr.decode()


def test_oracle_build_reference_matches_golden(oracle_bin, tmp_path):
    """reference tests/lib.rs:132-143 (only the peptide FASTA is diffed; the HashSet file order is arbitrary)."""
    d = os.path.join(GOLDEN, "test_build")
    with open(tmp_path / "ref.fasta", "wb") as fo:
        r = subprocess.run([oracle_bin, "build_reference", "--reference", os.path.join(d, "reference.fa"), "-l4", "--output",
                            str(tmp_path / "ref.bin")], stdout=fo, stderr=subprocess.PIPE)
    assert r.returncode == 0, r.stderr.decode()
    assert open(tmp_path / "ref.fasta", "rb").read() == open(os.path.join(d, "expected_output", "reference_peptides.fasta"), "rb").read()
    # bincode layout of HashSet<Vec<u8>>: u64 count, then u64 length + bytes per item — same items as the reference's file
    import struct

    def items(path):
        b = open(path, "rb").read()
        n, o, out = struct.unpack_from("<Q", b, 0)[0], 8, set()
        for _ in range(n):
            ln = struct.unpack_from("<Q", b, o)[0]
            out.add(b[o + 8:o + 8 + ln])
            o += 8 + ln
        assert o == len(b)
        return out
    assert items(tmp_path / "ref.bin") == items(os.path.join(d, "expected_output", "reference.binary"))


@pytest.mark.parametrize("case,suffix", [("test_filter", "filtered"), ("test_filter_long", "filtered_long"), ("test_filter_fs", "filtered_fs")])
def test_oracle_filter_matches_golden(oracle_bin, case, suffix, tmp_path):
    """reference tests/lib.rs:145-209"""
    d = os.path.join(GOLDEN, case)
    with open(tmp_path / ("tumor.%s.fa" % suffix), "wb") as fo:
        r = subprocess.run([oracle_bin, "filter", "--reference", os.path.join(d, "reference.binary"), "-l", "9", "--tsv", os.path.join(d, "info.tsv"),
                            "--tsv-output", str(tmp_path / ("info.%s.tsv" % suffix)), "--normal-output", str(tmp_path / ("normal.%s.fa" % suffix)),
                            "-s", str(tmp_path / "removed.tsv"), "-p", str(tmp_path / "removed.fa")], stdout=fo, stderr=subprocess.PIPE)
    assert r.returncode == 0, r.stderr.decode()
    for name in ("tumor.%s.fa", "normal.%s.fa", "info.%s.tsv"):
        name = name % suffix
        assert open(tmp_path / name, "rb").read() == open(os.path.join(d, "expected_output", name), "rb").read(), name
