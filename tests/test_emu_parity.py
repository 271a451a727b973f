"""Host-side logic (parsers, packer, closed-form window logic, residue, writers) against the
reference goldens and the oracle, with the device side stood in by the test-only CPU emulator.
The CUDA path itself is covered by tests/test_gpu_parity.py (-m gpu)."""
import os
import sys

import pytest

from conftest import GOLDEN, ROOT, make_bcf, materialize_reference, read_outputs, run_cli

sys.path.insert(0, ROOT)
from microphaser_b200 import synth  # noqa: E402

SOMATIC = ["forward_somatic", "empty", "reverse_somatic", "splice_forward_somatic", "splice_reverse_somatic"]
NORMAL = ["forward_normal", "splice_forward_normal"]


@pytest.mark.parametrize("case", SOMATIC)
def test_emulated_path_matches_reference_golden(emu_bin, case, tmp_path):
    d = os.path.join(GOLDEN, case)
    fa = materialize_reference(d, str(tmp_path))
    res = run_cli(emu_bin, d, str(tmp_path), ref=fa)
    assert res.returncode == 0, res.stderr.decode()
    for name in sorted(os.listdir(os.path.join(d, "expected"))):
        assert open(tmp_path / name, "rb").read() == open(os.path.join(d, "expected", name), "rb").read(), name


@pytest.mark.parametrize("case", ["reverse_somatic", "splice_forward_somatic"])
def test_binary_bcf_variants_match_reference_golden(emu_bin, case, tmp_path):
    """The reference reads its variants with bcf::Reader::from_path (src/main.rs:75): BCF2 input must give the golden bytes too
    (typed INFO values, multi-ALT sites, the ANN string and the SOMATIC flag all come out of the binary records)."""
    d = os.path.join(GOLDEN, case)
    fa = materialize_reference(d, str(tmp_path))
    res = run_cli(emu_bin, d, str(tmp_path), ref=fa, variants=make_bcf(d, str(tmp_path)))
    assert res.returncode == 0, res.stderr.decode()
    for name in sorted(os.listdir(os.path.join(d, "expected"))):
        assert open(tmp_path / name, "rb").read() == open(os.path.join(d, "expected", name), "rb").read(), name


def test_emulated_unsorted_gtf_is_fatal(emu_bin, tmp_path):
    d = os.path.join(GOLDEN, "unsorted_gtf")
    fa = materialize_reference(d, str(tmp_path))
    assert run_cli(emu_bin, d, str(tmp_path), gtf="unsorted.gtf", ref=fa).returncode != 0
    assert run_cli(emu_bin, d, str(tmp_path), gtf="sorted.gtf", ref=fa).returncode == 0


PROFILES = {
    "snv": dict(),
    "indel": dict(indel_frac=0.3, multiallelic_frac=0.15, start_loss_frac=0.5),
    "geom": dict(first_frame=True, short_exon_frac=0.3, exon_len=(27, 120), multiallelic_frac=0.1, indel_frac=0.1),
    "dense": dict(germline_per_kb=15.0, somatic_per_kb=15.0, indel_frac=0.2, multiallelic_frac=0.1, lowq_frac=0.08),
    "fs": dict(indel_frac=0.3, frameshift_ok=True, somatic_per_kb=4.0),
    "multi": dict(transcripts_per_gene=3, indel_frac=0.1),
    "carry": dict(intron_len=(15, 80), indel_frac=0.1, multiallelic_frac=0.05),  # introns shorter than a read: observations survive into the next exon
    "anti": dict(transcripts_per_gene=3, antisense_frac=0.6, multiallelic_frac=0.3, indel_frac=0.15),  # multi-allelic sites in genes with transcripts on both strands
    "dups": dict(dup_mate_frac=0.15, dup_extra_frac=0.5, lowq_frac=0.06, indel_frac=0.1),  # three and more reads sharing (start, qname): `contains` on the reverse strand
}


@pytest.mark.parametrize("profile,seed", [(p, s) for p in PROFILES for s in (11, 12, 13)])
def test_emulated_path_matches_oracle_on_synthetic(emu_bin, oracle_bin, profile, seed, tmp_path):
    kw = dict(PROFILES[profile])
    kw.update(seed=seed * 7919 + len(profile), n_genes=3, coverage=25.0)
    d = str(tmp_path / "in")
    synth.generate(d, synth.Params(**kw))
    o, p = tmp_path / "o", tmp_path / "p"
    o.mkdir()
    p.mkdir()
    ro = run_cli(oracle_bin, d, str(o))
    rp = run_cli(emu_bin, d, str(p))
    if rp.returncode == 3:
        pytest.skip("input needs the serial replay path: " + rp.stderr.decode().strip()[-120:])
    assert (ro.returncode == 0) == (rp.returncode == 0), (ro.stderr.decode()[-300:], rp.stderr.decode()[-300:])
    if ro.returncode == 0:
        assert read_outputs(str(o)) == read_outputs(str(p))


@pytest.mark.parametrize("case", NORMAL)
def test_emulated_normal_mode_matches_reference_golden(emu_bin, oracle_bin, case, tmp_path):
    """`normal` sub-command: FASTA against the reference's expected file, TSV (never diffed upstream) against the oracle."""
    d = os.path.join(GOLDEN, case)
    fa = materialize_reference(d, str(tmp_path))
    res = run_cli(emu_bin, d, str(tmp_path), ref=fa, subcommand="normal")
    assert res.returncode == 0, res.stderr.decode()
    assert open(tmp_path / "out.fa", "rb").read() == open(os.path.join(d, "expected", "out.fa"), "rb").read()
    o = tmp_path / "o"
    o.mkdir()
    assert run_cli(oracle_bin, d, str(o), ref=fa, subcommand="normal").returncode == 0
    assert open(tmp_path / "out.tsv", "rb").read() == open(o / "out.tsv", "rb").read()


@pytest.mark.parametrize("profile,seed", [(p, s) for p in PROFILES for s in (31, 32)])
def test_emulated_normal_mode_matches_oracle_on_synthetic(emu_bin, oracle_bin, profile, seed, tmp_path):
    """Both strands, indels, frameshifts, short exons: the reverse-strand `normal` path has no upstream fixture."""
    kw = dict(PROFILES[profile])
    kw.update(seed=seed * 6007 + len(profile), n_genes=3, coverage=25.0)
    d = str(tmp_path / "in")
    synth.generate(d, synth.Params(**kw))
    o, p = tmp_path / "o", tmp_path / "p"
    o.mkdir()
    p.mkdir()
    ro = run_cli(oracle_bin, d, str(o), subcommand="normal")
    rp = run_cli(emu_bin, d, str(p), subcommand="normal")
    if rp.returncode == 3:
        pytest.skip("input needs the serial replay path: " + rp.stderr.decode().strip()[-120:])
    assert (ro.returncode == 0) == (rp.returncode == 0), (ro.stderr.decode()[-300:], rp.stderr.decode()[-300:])
    if ro.returncode == 0:
        assert read_outputs(str(o), "normal") == read_outputs(str(p), "normal")


@pytest.mark.parametrize("case", ["reverse_somatic", "splice_forward_somatic"])
def test_threaded_alignment_reader_gives_the_same_records(emu_bin, case, tmp_path, monkeypatch):
    """MPH_IO_THREADS > 1: BGZF blocks are inflated and BAM records decoded by several threads (the library's file drivers
    do that by default); the output must not change."""
    monkeypatch.setenv("MPH_IO_THREADS", "4")
    d = os.path.join(GOLDEN, case)
    fa = materialize_reference(d, str(tmp_path))
    res = run_cli(emu_bin, d, str(tmp_path), ref=fa)
    assert res.returncode == 0, res.stderr.decode()
    for name in sorted(os.listdir(os.path.join(d, "expected"))):
        assert open(tmp_path / name, "rb").read() == open(os.path.join(d, "expected", name), "rb").read(), name


def test_threaded_alignment_reader_on_a_large_file(emu_bin, oracle_bin, tmp_path, monkeypatch):
    """Enough records for several decode threads per batch and records that straddle batch boundaries."""
    d = str(tmp_path / "in")
    synth.generate(d, synth.Params(seed=99, n_genes=12, coverage=120.0, exons=(6, 8)))
    o, p = tmp_path / "o", tmp_path / "p"
    o.mkdir()
    p.mkdir()
    assert run_cli(oracle_bin, d, str(o)).returncode == 0
    monkeypatch.setenv("MPH_IO_THREADS", "4")
    rp = run_cli(emu_bin, d, str(p))
    assert rp.returncode == 0, rp.stderr.decode()
    assert read_outputs(str(o)) == read_outputs(str(p))


@pytest.mark.parametrize("case", ["reverse_somatic", "forward_somatic"])
def test_wide_read_record_format_gives_the_same_records(emu_bin, case, tmp_path, monkeypatch):
    """MPH_PACK_WIDE=1 packs every read with 4-bit bases and a quality bitmask (the format of reads that contain other
    letters than A C G T); the default is 2-bit bases and a short list of low-quality positions."""
    monkeypatch.setenv("MPH_PACK_WIDE", "1")
    d = os.path.join(GOLDEN, case)
    fa = materialize_reference(d, str(tmp_path))
    res = run_cli(emu_bin, d, str(tmp_path), ref=fa)
    assert res.returncode == 0, res.stderr.decode()
    for name in sorted(os.listdir(os.path.join(d, "expected"))):
        assert open(tmp_path / name, "rb").read() == open(os.path.join(d, "expected", name), "rb").read(), name
