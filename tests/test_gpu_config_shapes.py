"""BASELINE.json's synthetic configs at the sizes they name, CUDA path vs the oracle, byte for byte:
mph_synth_batch -> mph_phase_batch -> mph_result_write on one side, the oracle on the files
mph_synth_write_files writes for the same parameters on the other. This is the batch bench.py times
(same generator, same packer entry), so the benched workload itself is pinned to the oracle.

  C2  chr22 exome, 450 transcripts, 30x                                  (full size)
  C3  whole exome at 100x: a 2 000-transcript slice of the 20 000         (the oracle needs ~1 min for it)
  C4  hypermutated: 10 somatic / kb, 10 % insertions + 10 % deletions     (full size, 450 transcripts)
  C5a `normal` healthy peptidome, 450 transcripts                         (every window is a record)
"""
import json
import os

import pytest

from conftest import CONFIG_SHAPES, read_outputs, run_oracle_on_files

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", sorted(CONFIG_SHAPES))
def test_config_shape_through_c_abi_matches_oracle(product, oracle_bin, shape, tmp_path):
    import microphaser_b200 as m
    kw = dict(CONFIG_SHAPES[shape])
    mode = kw["mode"]
    files, ora, out = tmp_path / "files", tmp_path / "oracle", tmp_path / "out"
    for d in (files, ora, out):
        d.mkdir()
    file_kw = {k: v for k, v in kw.items() if k != "mode"}
    m.synth_write_files(str(files), **file_kw)
    env = dict(os.environ, MPH_ORACLE_STATS=str(tmp_path / "stats.json"))
    ro = run_oracle_on_files(oracle_bin, str(files), str(ora), mode, env=env)
    assert ro.returncode == 0, ro.stderr.decode()
    ctx = m.Context(0)
    batch = m.Batch.synthetic(pin=True, **kw)
    res = ctx.phase_batch(batch)
    t = ctx.timing()
    res.write(str(out / "out.fa"), str(out / "out.tsv"), str(out / "out.normal.fa"))
    n_records = len(res)
    res.close()
    # the resident entry points give the same records
    ctx.upload(batch)
    ctx.phase_resident()
    res2 = ctx.collect()
    out2 = tmp_path / "out2"
    out2.mkdir()
    res2.write(str(out2 / "out.fa"), str(out2 / "out.tsv"), str(out2 / "out.normal.fa"))
    res2.close()
    # the file driver on the same files: several shards (MPH_PACK_THREADS) phased one after the other on this context, whose
    # device buffers still hold the previous calls' windows (a stale flag of a replayed transcript's window was once read by
    # the record kernels that run beside the replay)
    out3 = tmp_path / "out3"
    out3.mkdir()
    os.environ["MPH_PACK_THREADS"] = "5"
    os.environ["MPH_GPU_INFLATE"] = "1"  # and with the BGZF blocks inflated on the device (kernels/inflate_kernels.cu; off by default)
    try:
        paths = [str(files / n) for n in ("reads.bam", "ref.fa", "variants.vcf", "annotation.gtf")]
        if mode == 1 or mode == "normal":
            ctx.run_normal(*paths, str(out3 / "out.fa"), str(out3 / "out.tsv"))
        else:
            ctx.run_somatic(*paths, str(out3 / "out.fa"), str(out3 / "out.tsv"), str(out3 / "out.normal.fa"))
    finally:
        del os.environ["MPH_PACK_THREADS"]
        del os.environ["MPH_GPU_INFLATE"]
    ctx.close()
    want = read_outputs(str(ora), mode)
    assert read_outputs(str(out), mode) == want
    assert read_outputs(str(out2), mode) == want
    assert read_outputs(str(out3), mode) == want
    assert n_records + 1 == want["out.tsv"].count(b"\n") and n_records > 1000
    st = json.load(open(tmp_path / "stats.json"))
    assert t["windows"] == st["windows"], "main-ORF window count differs from the oracle's print_haplotypes calls"
    assert t["read_windows"] == st["read_windows"], "sum of depth differs from the oracle"
