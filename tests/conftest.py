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


EMU_BIN = os.path.join(ROOT, "tests", "emu", "_build", "mph_emu")


def build_emu():
    """Compile the test-only CPU emulator CLI (product host code + emulated device side)."""
    srcs = []
    for d in ("tests/emu", "microphaser_b200/csrc/host", "microphaser_b200/csrc/core", "microphaser_b200/csrc/io"):
        srcs += [os.path.join(ROOT, d, f) for f in os.listdir(os.path.join(ROOT, d)) if f.endswith((".cpp", ".hpp", ".h"))]
    if os.path.exists(EMU_BIN) and all(os.path.getmtime(s) <= os.path.getmtime(EMU_BIN) for s in srcs):
        return EMU_BIN
    os.makedirs(os.path.dirname(EMU_BIN), exist_ok=True)
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Wno-missing-field-initializers", "-o", EMU_BIN,
                    os.path.join(ROOT, "tests", "emu", "emu_cli.cpp"), "-lz"], check=True)
    return EMU_BIN


SYNTH_CHECK_BIN = os.path.join(ROOT, "tests", "emu", "_build", "synth_check")


def build_synth_check():
    """tests/emu/synth_check.cpp: the native synthetic workload through the emulated device + product host code."""
    build_emu()
    src = os.path.join(ROOT, "tests", "emu", "synth_check.cpp")
    if not os.path.exists(SYNTH_CHECK_BIN) or os.path.getmtime(SYNTH_CHECK_BIN) < max(os.path.getmtime(src), os.path.getmtime(EMU_BIN)):
        subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Wno-missing-field-initializers", "-o", SYNTH_CHECK_BIN, src, "-lz"], check=True)
    return SYNTH_CHECK_BIN


# The shapes BASELINE.json names (configs[1..4]) as arguments of the native generator: n_transcripts, coverage,
# germline / somatic variants per kb, insertion / deletion share, mode, seed. "Full" is the size the -m gpu tests
# diff against the oracle; the CPU suite runs the same shapes through the emulator on fewer transcripts.
CONFIG_SHAPES = {
    "C2_chr22": dict(n_transcripts=450, coverage=30.0, germline_per_kb=1.0, somatic_per_kb=1.0, ins_var_frac=0.0, del_var_frac=0.0, mode="somatic", seed=0x4D500002),
    "C3_exome_slice": dict(n_transcripts=2000, coverage=100.0, germline_per_kb=1.0, somatic_per_kb=1.0, ins_var_frac=0.0, del_var_frac=0.0, mode="somatic", seed=0x4D500003),
    "C4_hypermutated": dict(n_transcripts=450, coverage=30.0, germline_per_kb=1.0, somatic_per_kb=10.0, ins_var_frac=0.1, del_var_frac=0.1, mode="somatic", seed=0x4D500004),
    "C5a_normal": dict(n_transcripts=450, coverage=30.0, germline_per_kb=1.0, somatic_per_kb=1.0, ins_var_frac=0.0, del_var_frac=0.0, mode="normal", seed=0x4D500005),
}


def run_oracle_on_files(oracle_bin, files_dir, out_dir, mode="somatic", env=None):
    """The oracle CLI on a directory written by mph_synth_write_files; returns CompletedProcess."""
    cmd = [oracle_bin, mode, os.path.join(files_dir, "reads.bam"), "-r", os.path.join(files_dir, "ref.fa"), "-b",
           os.path.join(files_dir, "variants.vcf"), "-t", os.path.join(out_dir, "out.tsv")]
    if mode == "somatic":
        cmd += ["-n", os.path.join(out_dir, "out.normal.fa")]
    with open(os.path.join(files_dir, "annotation.gtf")) as gin, open(os.path.join(out_dir, "out.fa"), "wb") as fo:
        return subprocess.run(cmd, stdin=gin, stdout=fo, stderr=subprocess.PIPE, timeout=1800, env=env)


def build_product():
    from microphaser_b200 import build
    return build.build_all()


def build_abi_host():
    """tests/abi_host/abi_host.cpp: a host that reaches the library through the C ABI's packer entry points only."""
    lib, _ = build_product()
    out = os.path.join(ROOT, "tests", "abi_host", "_build")
    os.makedirs(out, exist_ok=True)
    exe = os.path.join(out, "abi_host")
    src = os.path.join(ROOT, "tests", "abi_host", "abi_host.cpp")
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(src), os.path.getmtime(lib)):
        subprocess.run(["g++", "-std=c++17", "-O2", "-o", exe, src, "-L" + os.path.dirname(lib), "-lmicrophaser_gpu", "-lz",
                        "-Wl,-rpath," + os.path.dirname(lib)], check=True)
    return exe


def make_bcf(case_dir, out_dir):
    """The case's variants as binary BCF2 (tests/golden/bcflite.py), what `bcf::Reader::from_path` also takes."""
    sys.path.insert(0, GOLDEN)
    import bcflite
    path = os.path.join(out_dir, "variants.bcf")
    bcflite.vcf_to_bcf(os.path.join(case_dir, "variants.vcf"), path)
    return path


def run_cli(binary, case_dir, out_dir, gtf="annotation.gtf", ref=None, subcommand="somatic", variants=None):
    """Run `<binary> somatic ...` the way the reference's tests do (tests/lib.rs:23-35); returns CompletedProcess."""
    ref = ref or os.path.join(case_dir, "ref.fa")
    cmd = [binary, subcommand, os.path.join(case_dir, "reads.bam"), "--ref", ref, "--variants", variants or os.path.join(case_dir, "variants.vcf"),
           "--tsv", os.path.join(out_dir, "out.tsv")]
    if subcommand == "somatic":
        cmd += ["--normal-output", os.path.join(out_dir, "out.normal.fa")]
    with open(os.path.join(case_dir, gtf)) as gin, open(os.path.join(out_dir, "out.fa"), "wb") as fout:
        return subprocess.run(cmd, stdin=gin, stdout=fout, stderr=subprocess.PIPE, timeout=900)


def read_outputs(out_dir, subcommand="somatic"):
    names = ("out.fa", "out.tsv", "out.normal.fa") if subcommand == "somatic" else ("out.fa", "out.tsv")
    return {n: open(os.path.join(out_dir, n), "rb").read() for n in names}


@pytest.fixture(scope="session")
def abi_host():
    return build_abi_host()


@pytest.fixture(scope="session")
def oracle_bin():
    return build_oracle()


@pytest.fixture(scope="session")
def emu_bin():
    return build_emu()


@pytest.fixture(scope="session")
def synth_check_bin():
    return build_synth_check()


@pytest.fixture(scope="session")
def product():
    """(library path, CLI path) of the CUDA build; compiled on demand (nvcc cross-compiles without a GPU)."""
    return build_product()
