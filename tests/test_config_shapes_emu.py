"""The native synthetic workload (what bench.py packs with mph_synth_batch) against the oracle, without a GPU:
`synth_check` packs the batch with the product's packer, runs the emulated device side and the product's host
residue / writers, and writes the same workload as files; the oracle runs on the files. Byte-exact on all
streams. Pins (a) that the packed batch and the files describe the same genes / reads / variants, whichever
reads the generator skips, and (b) the four synthetic shapes of BASELINE.json at reduced transcript counts
(the -m gpu tests repeat them at full size through the C ABI)."""
import json
import os
import subprocess

import pytest

from conftest import CONFIG_SHAPES, read_outputs, run_oracle_on_files

CPU_SIZES = {"C2_chr22": 450, "C3_exome_slice": 250, "C4_hypermutated": 150, "C5a_normal": 150}


@pytest.mark.parametrize("shape", sorted(CONFIG_SHAPES))
def test_native_synthetic_batch_matches_oracle_on_its_files(synth_check_bin, oracle_bin, shape, tmp_path):
    kw = CONFIG_SHAPES[shape]
    mode = kw["mode"]
    out, files, ora = tmp_path / "out", tmp_path / "files", tmp_path / "oracle"
    for d in (out, files, ora):
        d.mkdir()
    r = subprocess.run([synth_check_bin, str(out), str(files), "1" if mode == "normal" else "0", str(kw["seed"]), str(CPU_SIZES[shape]),
                        str(kw["coverage"]), str(kw["germline_per_kb"]), str(kw["somatic_per_kb"]), str(kw["ins_var_frac"]), str(kw["del_var_frac"])],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=900)
    assert r.returncode == 0, r.stderr.decode()
    st = json.loads(r.stdout)
    ro = run_oracle_on_files(oracle_bin, str(files), str(ora), mode)
    assert ro.returncode == 0, ro.stderr.decode()
    assert read_outputs(str(out), mode) == read_outputs(str(ora), mode)
    assert st["records"] > 100 and st["records"] + 1 == open(ora / "out.tsv", "rb").read().count(b"\n")
    if shape == "C4_hypermutated":
        vcf = open(files / "variants.vcf").read().splitlines()
        body = [l.split("\t") for l in vcf if not l.startswith("#")]
        assert any(len(f[3]) > 1 for f in body) and any(len(f[4]) > 1 for f in body), "the hypermutated shape must contain deletions and insertions"
