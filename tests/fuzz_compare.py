#!/usr/bin/env python3
"""Differential fuzzing: product host code + device logic (CPU emulator or CUDA CLI) vs the oracle
on random synthetic inputs.  usage: fuzz_compare.py <binary> <n_seeds> [first_seed] [profile] [somatic|normal]"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from microphaser_b200 import synth  # noqa: E402

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


def run(binary, d, out, sub="somatic"):
    cmd = [binary, sub, os.path.join(d, "reads.bam"), "-r", os.path.join(d, "ref.fa"), "-b", os.path.join(d, "variants.vcf"), "-t", out + ".tsv"]
    if sub == "somatic":
        cmd += ["-n", out + ".normal.fa"]
    with open(os.path.join(d, "annotation.gtf")) as gin, open(out + ".fa", "wb") as fo:
        r = subprocess.run(cmd, stdin=gin, stdout=fo, stderr=subprocess.PIPE, timeout=600)
    return r.returncode, r.stderr.decode()[-400:]


def main():
    binary = sys.argv[1]
    n = int(sys.argv[2])
    first = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    profs = sys.argv[4].split(",") if len(sys.argv) > 4 and sys.argv[4] != "all" else list(PROFILES)
    sub = sys.argv[5] if len(sys.argv) > 5 else "somatic"
    oracle = os.path.join(ROOT, "oracle", "_build", "mph_oracle")
    stats = dict(ok=0, both_fail=0, unsupported=0, mismatch=0, records=0)
    for seed in range(first, first + n):
        prof = profs[seed % len(profs)]
        kw = dict(PROFILES[prof])
        kw.update(seed=seed, n_genes=3, coverage=25.0)
        with tempfile.TemporaryDirectory() as d:
            synth.generate(d, synth.Params(**kw))
            rc_o, err_o = run(oracle, d, os.path.join(d, "o"), sub)
            rc_p, err_p = run(binary, d, os.path.join(d, "p"), sub)
            if rc_o != 0 and rc_p != 0:
                stats["both_fail"] += 1
                continue
            if rc_p == 3:
                stats["unsupported"] += 1
                print("seed %d [%s]: unsupported: %s" % (seed, prof, err_p.strip().split("\n")[-1]))
                continue
            same = rc_o == rc_p
            for ext in ((".fa", ".tsv", ".normal.fa") if sub == "somatic" else (".fa", ".tsv")):
                a = open(os.path.join(d, "o") + ext, "rb").read()
                b = open(os.path.join(d, "p") + ext, "rb").read()
                same = same and a == b
            if same:
                stats["ok"] += 1
                stats["records"] += open(os.path.join(d, "o.tsv")).read().count("\n")
            else:
                stats["mismatch"] += 1
                keep = "/tmp/fuzz_fail_%d" % seed
                subprocess.run(["rm", "-rf", keep])
                subprocess.run(["cp", "-r", d, keep])
                print("seed %d [%s]: MISMATCH rc oracle=%d product=%d (kept %s) %s | %s" % (seed, prof, rc_o, rc_p, keep, err_o.strip()[-150:], err_p.strip()[-150:]))
    print(stats)
    return 1 if stats["mismatch"] else 0


if __name__ == "__main__":
    sys.exit(main())
