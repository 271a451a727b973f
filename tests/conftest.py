"""Shared pytest plumbing: markers, paths, build helpers, golden-reference materialisation."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
ORACLE_BIN = os.path.join(ROOT, "oracle", "_build", "mph_oracle")
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def materialize_reference(case_dir, out_dir, name="ref.fa"):
    """Write a sparse chrN.fa (+ .fai) at the original hg38 coordinates from ref_patch.tsv.

    The .fai is the reference's own index (tests/resources/chrN.fa.fai), so offsets are the
    real ones; the file is sparse (a few KB on disk).  Gene spans of the case's GTF(s) are
    pre-filled with 'N', known bases come from ref_patch.tsv.
    """
    fai = open(os.path.join(case_dir, "ref.fa.fai")).read().split("\n")[0].split("\t")
    chrom, length, offset, lb, lw = fai[0], int(fai[1]), int(fai[2]), int(fai[3]), int(fai[4])
    path = os.path.join(out_dir, name)

    def foff(p):
        return offset + (p // lb) * lw + p % lb

    with open(path, "wb") as f:
        f.truncate(foff(length - 1) + 2)

        def put(start, seq):
            # write bases start.. honouring the line geometry (newline after every lb bases)
            p = start
            i = 0
            while i < len(seq):
                take = min(lb - p % lb, len(seq) - i)
                f.seek(foff(p))
                f.write(seq[i:i + take])
                if (p + take) % lb == 0:
                    f.write(b"\n" * (lw - lb))
                p += take
                i += take

        for fn in sorted(os.listdir(case_dir)):
            if not fn.endswith(".gtf"):
                continue
            for line in open(os.path.join(case_dir, fn)):
                t = line.split("\t")
                if len(t) >= 9 and t[2] == "gene" and t[0] == chrom:
                    s, e = int(t[3]) - 1, min(int(t[4]) + 100, length)
                    put(s, b"N" * (e - s))
        for line in open(os.path.join(case_dir, "ref_patch.tsv")):
            if line.startswith("#"):
                continue
            c, s, seq = line.rstrip("\n").split("\t")
            assert c == chrom
            put(int(s), seq.encode())
    with open(path + ".fai", "w") as f:
        f.write("\t".join(fai) + "\n")
    return path


def build_oracle():
    """Compile the CPU oracle (test infrastructure) if it is missing or stale."""
    srcs = [os.path.join(ROOT, "oracle", f) for f in os.listdir(os.path.join(ROOT, "oracle")) if f.endswith((".cpp", ".hpp"))]
    srcs += [os.path.join(ROOT, "microphaser_b200", "csrc", "io", f) for f in ("hts_io.hpp", "fmt_util.hpp")]
    if os.path.exists(ORACLE_BIN) and all(os.path.getmtime(s) <= os.path.getmtime(ORACLE_BIN) for s in srcs):
        return ORACLE_BIN
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    return ORACLE_BIN


@pytest.fixture(scope="session")
def oracle_bin():
    return build_oracle()
