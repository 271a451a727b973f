"""Parity of the CUDA path (through the C ABI / the CLI built on it) with the reference goldens and
with the oracle. Bit-exact: all outputs are integer / byte work plus one IEEE-754 f64 division per
record computed from identical integer counts."""
import os
import sys

import pytest

from conftest import GOLDEN, ROOT, make_bcf, materialize_reference, read_outputs, run_cli

sys.path.insert(0, ROOT)
from microphaser_b200 import synth  # noqa: E402
from test_emu_parity import NORMAL, PROFILES, SOMATIC  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", SOMATIC)
def test_cuda_cli_matches_reference_golden(product, case, tmp_path):
    d = os.path.join(GOLDEN, case)
    fa = materialize_reference(d, str(tmp_path))
    res = run_cli(product[1], d, str(tmp_path), ref=fa)
    assert res.returncode == 0, res.stderr.decode()
    for name in sorted(os.listdir(os.path.join(d, "expected"))):
        assert open(tmp_path / name, "rb").read() == open(os.path.join(d, "expected", name), "rb").read(), name


def test_cuda_cli_reads_binary_bcf(product, tmp_path):
    """Variants as BCF2 (what the reference's README examples pass) through the CUDA CLI: same golden bytes."""
    d = os.path.join(GOLDEN, "reverse_somatic")
    fa = materialize_reference(d, str(tmp_path))
    res = run_cli(product[1], d, str(tmp_path), ref=fa, variants=make_bcf(d, str(tmp_path)))
    assert res.returncode == 0, res.stderr.decode()
    for name in sorted(os.listdir(os.path.join(d, "expected"))):
        assert open(tmp_path / name, "rb").read() == open(os.path.join(d, "expected", name), "rb").read(), name


def test_cuda_unsorted_gtf_is_fatal(product, tmp_path):
    d = os.path.join(GOLDEN, "unsorted_gtf")
    fa = materialize_reference(d, str(tmp_path))
    assert run_cli(product[1], d, str(tmp_path), gtf="unsorted.gtf", ref=fa).returncode != 0
    assert run_cli(product[1], d, str(tmp_path), gtf="sorted.gtf", ref=fa).returncode == 0


def test_c_abi_run_somatic_matches_golden(product, tmp_path):
    """Same call a host would make through FFI: mph_ctx_create + mph_run_somatic."""
    import microphaser_b200 as m
    d = os.path.join(GOLDEN, "reverse_somatic")
    fa = materialize_reference(d, str(tmp_path))
    ctx = m.Context(0)
    ctx.run_somatic(os.path.join(d, "reads.bam"), fa, os.path.join(d, "variants.vcf"), os.path.join(d, "annotation.gtf"),
                    str(tmp_path / "out.fa"), str(tmp_path / "out.tsv"), str(tmp_path / "out.normal.fa"))
    t = ctx.timing()
    ctx.close()
    for name in ("out.fa", "out.tsv", "out.normal.fa"):
        assert open(tmp_path / name, "rb").read() == open(os.path.join(d, "expected", name), "rb").read(), name
    assert t["kernel_launches"] >= 6 and t["windows"] > 0 and t["n_records"] == 27


@pytest.mark.parametrize("profile,seed", [(p, s) for p in PROFILES for s in (21, 22, 23, 24)])
def test_cuda_matches_oracle_on_synthetic(product, oracle_bin, profile, seed, tmp_path):
    kw = dict(PROFILES[profile])
    kw.update(seed=seed * 104729 + len(profile), n_genes=4, coverage=30.0)
    d = str(tmp_path / "in")
    synth.generate(d, synth.Params(**kw))
    o, p = tmp_path / "o", tmp_path / "p"
    o.mkdir()
    p.mkdir()
    ro = run_cli(oracle_bin, d, str(o))
    rp = run_cli(product[1], d, str(p))
    if rp.returncode == 3:
        pytest.skip("input needs the serial replay path: " + rp.stderr.decode().strip()[-120:])
    assert (ro.returncode == 0) == (rp.returncode == 0), (ro.stderr.decode()[-300:], rp.stderr.decode()[-300:])
    if ro.returncode == 0:
        assert read_outputs(str(o)) == read_outputs(str(p))


def test_cuda_matches_oracle_chr22_shape(product, oracle_bin, tmp_path):
    """BASELINE.json config 2 geometry (8 exons, 150 bp reads, 30x, SNVs) on 40 transcripts, incl. window statistics."""
    import json
    import subprocess
    kw = dict(seed=0x4D500002, n_genes=40, exons=(8, 8), exon_len=(90, 250), intron_len=(300, 2000), read_len=150, coverage=30.0,
              germline_per_kb=1.0, somatic_per_kb=1.0, lowercase_frac=0.0)
    d = str(tmp_path / "in")
    synth.generate(d, synth.Params(**kw))
    o, p = tmp_path / "o", tmp_path / "p"
    o.mkdir()
    p.mkdir()
    env = dict(os.environ, MPH_ORACLE_STATS=str(tmp_path / "stats.json"))
    with open(os.path.join(d, "annotation.gtf")) as gin, open(o / "out.fa", "wb") as fo:
        ro = subprocess.run([oracle_bin, "somatic", os.path.join(d, "reads.bam"), "-r", os.path.join(d, "ref.fa"), "-b",
                             os.path.join(d, "variants.vcf"), "-t", str(o / "out.tsv"), "-n", str(o / "out.normal.fa")],
                            stdin=gin, stdout=fo, stderr=subprocess.PIPE, env=env)
    assert ro.returncode == 0, ro.stderr.decode()
    import microphaser_b200 as m
    ctx = m.Context(0)
    try:
        ctx.run_somatic(os.path.join(d, "reads.bam"), os.path.join(d, "ref.fa"), os.path.join(d, "variants.vcf"),
                        os.path.join(d, "annotation.gtf"), str(p / "out.fa"), str(p / "out.tsv"), str(p / "out.normal.fa"))
    except m.MphError as e:
        if e.code == m.MPH_ERR_UNSUPPORTED:
            pytest.skip(str(e))
        raise
    t = ctx.timing()
    ctx.close()
    assert read_outputs(str(o)) == read_outputs(str(p))
    st = json.load(open(tmp_path / "stats.json"))
    assert t["windows"] == st["windows"], "main-ORF window count differs from the oracle's print_haplotypes calls"
    assert t["read_windows"] == st["read_windows"], "sum of depth differs from the oracle"


@pytest.mark.parametrize("case", ["reverse_somatic", "forward_somatic"])
def test_wide_histogram_kernel_matches_golden(product, case, tmp_path, monkeypatch):
    """MPH_FORCE_WIDE=1 routes every window with extra haplotype keys through the warp-per-window overflow kernel."""
    monkeypatch.setenv("MPH_FORCE_WIDE", "1")
    d = os.path.join(GOLDEN, case)
    fa = materialize_reference(d, str(tmp_path))
    res = run_cli(product[1], d, str(tmp_path), ref=fa)
    assert res.returncode == 0, res.stderr.decode()
    for name in sorted(os.listdir(os.path.join(d, "expected"))):
        assert open(tmp_path / name, "rb").read() == open(os.path.join(d, "expected", name), "rb").read(), name


def test_multi_device_sharding_matches_oracle(product, oracle_bin, tmp_path):
    """mph_run_somatic_multi: genes split into contiguous ranges over several contexts (two GPUs when the box has
    them, else two contexts on GPU 0); ordered concatenation must equal the oracle's single-process output."""
    import torch
    import microphaser_b200 as m
    d = str(tmp_path / "in")
    synth.generate(d, synth.Params(seed=4242, n_genes=7, coverage=30.0, indel_frac=0.1, multiallelic_frac=0.1))
    o, p = tmp_path / "o", tmp_path / "p"
    o.mkdir()
    p.mkdir()
    ro = run_cli(oracle_bin, d, str(o))
    assert ro.returncode == 0, ro.stderr.decode()
    n_dev = torch.cuda.device_count()
    devs = (list(range(min(n_dev, 4))) + [0]) if n_dev >= 2 else [0, 0, 0]  # distinct devices when the box has them
    ctxs = [m.Context(x) for x in devs]
    try:
        m.run_somatic_multi(ctxs, os.path.join(d, "reads.bam"), os.path.join(d, "ref.fa"), os.path.join(d, "variants.vcf"),
                            os.path.join(d, "annotation.gtf"), str(p / "out.fa"), str(p / "out.tsv"), str(p / "out.normal.fa"))
    except m.MphError as e:
        if e.code == m.MPH_ERR_UNSUPPORTED:
            pytest.skip(str(e))
        raise
    finally:
        for c in ctxs:
            c.close()
    assert read_outputs(str(o)) == read_outputs(str(p))
    assert len(read_outputs(str(o))["out.tsv"]) > 0


def test_multi_device_normal_mode_matches_oracle(product, oracle_bin, tmp_path):
    """mph_run_normal_multi: the `normal` sub-command sharded by gene range over several contexts (distinct devices when the
    box has them); ordered concatenation must equal the oracle's single-process output."""
    import torch
    import microphaser_b200 as m
    d = str(tmp_path / "in")
    synth.generate(d, synth.Params(seed=4243, n_genes=9, coverage=20.0, indel_frac=0.1, multiallelic_frac=0.1))
    o, p = tmp_path / "o", tmp_path / "p"
    o.mkdir()
    p.mkdir()
    ro = run_cli(oracle_bin, d, str(o), subcommand="normal")
    assert ro.returncode == 0, ro.stderr.decode()
    n_dev = torch.cuda.device_count()
    devs = (list(range(min(n_dev, 4))) + [0]) if n_dev >= 2 else [0, 0, 0]
    ctxs = [m.Context(x) for x in devs]
    try:
        m.run_normal_multi(ctxs, os.path.join(d, "reads.bam"), os.path.join(d, "ref.fa"), os.path.join(d, "variants.vcf"),
                           os.path.join(d, "annotation.gtf"), str(p / "out.fa"), str(p / "out.tsv"))
    except m.MphError as e:
        if e.code == m.MPH_ERR_UNSUPPORTED:
            pytest.skip(str(e))
        raise
    finally:
        for c in ctxs:
            c.close()
    assert read_outputs(str(o), "normal") == read_outputs(str(p), "normal")
    assert len(read_outputs(str(o), "normal")["out.tsv"]) > 0


# ---------------------------------------------------------------- `normal` mode (src/normal_microphasing.rs)
@pytest.mark.parametrize("case", NORMAL)
def test_cuda_normal_mode_matches_reference_golden(product, oracle_bin, case, tmp_path):
    d = os.path.join(GOLDEN, case)
    fa = materialize_reference(d, str(tmp_path))
    res = run_cli(product[1], d, str(tmp_path), ref=fa, subcommand="normal")
    assert res.returncode == 0, res.stderr.decode()
    assert open(tmp_path / "out.fa", "rb").read() == open(os.path.join(d, "expected", "out.fa"), "rb").read()
    o = tmp_path / "o"
    o.mkdir()
    assert run_cli(oracle_bin, d, str(o), ref=fa, subcommand="normal").returncode == 0
    assert open(tmp_path / "out.tsv", "rb").read() == open(o / "out.tsv", "rb").read()


@pytest.mark.parametrize("wide", [False, True])
@pytest.mark.parametrize("profile,seed", [(p, s) for p in PROFILES for s in (41, 42, 43)])
def test_cuda_normal_mode_matches_oracle_on_synthetic(product, oracle_bin, profile, seed, wide, tmp_path, monkeypatch):
    if wide:
        if seed != 41:
            pytest.skip("the overflow kernel is exercised on one seed per profile")
        monkeypatch.setenv("MPH_FORCE_WIDE", "1")
    kw = dict(PROFILES[profile])
    kw.update(seed=seed * 15485863 + len(profile), n_genes=4, coverage=30.0)
    d = str(tmp_path / "in")
    synth.generate(d, synth.Params(**kw))
    o, p = tmp_path / "o", tmp_path / "p"
    o.mkdir()
    p.mkdir()
    ro = run_cli(oracle_bin, d, str(o), subcommand="normal")
    rp = run_cli(product[1], d, str(p), subcommand="normal")
    if rp.returncode == 3:
        pytest.skip("input needs the serial replay path: " + rp.stderr.decode().strip()[-120:])
    assert (ro.returncode == 0) == (rp.returncode == 0), (ro.stderr.decode()[-300:], rp.stderr.decode()[-300:])
    if ro.returncode == 0:
        assert read_outputs(str(o), "normal") == read_outputs(str(p), "normal")


def test_c_abi_run_normal_matches_oracle_with_statistics(product, oracle_bin, tmp_path):
    """mph_run_normal through the binding; window and observation counts against the oracle's own counters."""
    import json
    import subprocess
    import microphaser_b200 as m
    d = str(tmp_path / "in")
    synth.generate(d, synth.Params(seed=977, n_genes=12, coverage=30.0, indel_frac=0.1, multiallelic_frac=0.1))
    o, p = tmp_path / "o", tmp_path / "p"
    o.mkdir()
    p.mkdir()
    env = dict(os.environ, MPH_ORACLE_STATS=str(tmp_path / "stats.json"))
    with open(os.path.join(d, "annotation.gtf")) as gin, open(o / "out.fa", "wb") as fo:
        ro = subprocess.run([oracle_bin, "normal", os.path.join(d, "reads.bam"), "-r", os.path.join(d, "ref.fa"), "-b",
                             os.path.join(d, "variants.vcf"), "-t", str(o / "out.tsv")], stdin=gin, stdout=fo, stderr=subprocess.PIPE, env=env)
    assert ro.returncode == 0, ro.stderr.decode()
    ctx = m.Context(0)
    try:
        ctx.run_normal(os.path.join(d, "reads.bam"), os.path.join(d, "ref.fa"), os.path.join(d, "variants.vcf"),
                       os.path.join(d, "annotation.gtf"), str(p / "out.fa"), str(p / "out.tsv"))
    except m.MphError as e:
        if e.code == m.MPH_ERR_UNSUPPORTED:
            pytest.skip(str(e))
        raise
    t = ctx.timing()
    ctx.close()
    assert read_outputs(str(o), "normal") == read_outputs(str(p), "normal")
    if os.path.exists(tmp_path / "stats.json"):
        st = json.load(open(tmp_path / "stats.json"))
        assert t["windows"] == st["windows"] and t["read_windows"] == st["read_windows"]


# ---------------------------------------------------------------- the packer entry points of the C ABI
@pytest.mark.parametrize("case,sub", [("reverse_somatic", "somatic"), ("splice_forward_somatic", "somatic"), ("forward_normal", "normal")])
def test_c_abi_packer_entry_points_match_golden(abi_host, case, sub, tmp_path):
    """A host that flattens every gene into `mph_gene_in` arrays and calls mph_packer_add_gene / mph_phase_batch /
    mph_result_write (what the Rust binding of INTEGRATION.md does) must produce the reference's bytes."""
    import subprocess
    d = os.path.join(GOLDEN, case)
    fa = materialize_reference(d, str(tmp_path))
    cmd = [abi_host, sub, os.path.join(d, "reads.bam"), fa, os.path.join(d, "variants.vcf"), os.path.join(d, "annotation.gtf"),
           str(tmp_path / "out.fa"), str(tmp_path / "out.tsv")]
    if sub == "somatic":
        cmd.append(str(tmp_path / "out.normal.fa"))
    r = subprocess.run(cmd, stderr=subprocess.PIPE, timeout=600)
    assert r.returncode == 0, r.stderr.decode()
    names = sorted(os.listdir(os.path.join(d, "expected")))
    for name in names:
        assert open(tmp_path / name, "rb").read() == open(os.path.join(d, "expected", name), "rb").read(), name


def test_c_abi_packer_entry_points_match_oracle_on_synthetic(abi_host, oracle_bin, tmp_path):
    import subprocess
    d = str(tmp_path / "in")
    synth.generate(d, synth.Params(seed=31337, n_genes=6, coverage=30.0, indel_frac=0.2, multiallelic_frac=0.1, intron_len=(20, 400)))
    o, p = tmp_path / "o", tmp_path / "p"
    o.mkdir()
    p.mkdir()
    assert run_cli(oracle_bin, d, str(o)).returncode == 0
    r = subprocess.run([abi_host, "somatic", os.path.join(d, "reads.bam"), os.path.join(d, "ref.fa"), os.path.join(d, "variants.vcf"),
                        os.path.join(d, "annotation.gtf"), str(p / "out.fa"), str(p / "out.tsv"), str(p / "out.normal.fa")], stderr=subprocess.PIPE, timeout=600)
    assert r.returncode == 0, r.stderr.decode()
    assert read_outputs(str(o)) == read_outputs(str(p))


@pytest.mark.parametrize("sub", ["somatic", "normal"])
def test_file_driver_shards_do_not_change_output(product, oracle_bin, sub, tmp_path, monkeypatch):
    """The file drivers cut the genes into shards that are packed by parallel host threads, phased one after the other and
    written in order as they complete; three shards with a threaded alignment reader must give the oracle's bytes."""
    d = str(tmp_path / "in")
    synth.generate(d, synth.Params(seed=2718, n_genes=9, coverage=30.0, indel_frac=0.15, multiallelic_frac=0.1, intron_len=(30, 600)))
    o, p = tmp_path / "o", tmp_path / "p"
    o.mkdir()
    p.mkdir()
    assert run_cli(oracle_bin, d, str(o), subcommand=sub).returncode == 0
    monkeypatch.setenv("MPH_PACK_THREADS", "3")
    monkeypatch.setenv("MPH_IO_THREADS", "4")
    rp = run_cli(product[1], d, str(p), subcommand=sub)
    assert rp.returncode == 0, rp.stderr.decode()
    assert read_outputs(str(o), sub) == read_outputs(str(p), sub)
